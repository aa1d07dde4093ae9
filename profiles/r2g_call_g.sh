#!/bin/bash
# tools/r2_call_g.sh -- round 2, seventh GPU call: split coder launches (long streams on a high-priority stream) A/B; inverse BWT single walk timing
mkdir -p gpurun_out
run() { name=$1; shift; envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 500 python bench.py "$@" --no-cpu-baseline --no-e2e --no-extras --steps 3 --warmup 1 > gpurun_out/r2g_$name.json 2> gpurun_out/r2g_$name.err
  python -c "import json;d=json.load(open('gpurun_out/r2g_$name.json'));print('$name: value', round(d['value'],1), 'compress', round(d['compress_MBps'],1), 'decompress', round(d['decompress_MBps'],1), 'ms/step', round(d['ms_per_step']))" || tail -5 gpurun_out/r2g_$name.err
}
{
echo "== 1. parity with split launches"
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "== 2. phased steps: split launches on / off"
run split_on_64 X=1 -- --mode phased --blocks 64
run split_off_64 BSCB200_CODER_SPLIT=0 -- --mode phased --blocks 64
echo "== 2b. pipelined steps: split launches on / off, blocks in flight"
run pipe_split_on_64 X=1 -- --mode pipeline --blocks 64
run pipe_split_off_64 BSCB200_CODER_SPLIT=0 -- --mode pipeline --blocks 64
run pipe_split_on_96 X=1 -- --mode pipeline --blocks 96
run pipe_split_on_48 X=1 -- --mode pipeline --blocks 48
echo "== 3. inverse BWT: single walk, kernel-only sum of one call"
TRC_SIDE=ours TRC_STAGE=bwt_decode timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2g_launches_ours_bwt_decode.csv python tools/time_ref_cuda.py > /dev/null 2>&1
python tools/ncu_launch_sum.py gpurun_out/r2g_launches_ours_bwt_decode.csv 8
timeout 400 python tools/time_ref_cuda.py 3 > gpurun_out/r2g_ref_cuda.json 2> gpurun_out/r2g_ref_cuda.err; tail -9 gpurun_out/r2g_ref_cuda.err
} 2>&1 | tee gpurun_out/r2_call_g.log
