#!/bin/bash
# tools/r2_call_f.sh -- round 2, sixth GPU call: phased vs pipelined steps, coder-slot cap, blocks in flight
mkdir -p gpurun_out
run() { # name, env..., -- bench args
  name=$1; shift
  envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 500 python bench.py "$@" --no-cpu-baseline --no-e2e --no-extras --steps 3 --warmup 1 > gpurun_out/r2f_$name.json 2> gpurun_out/r2f_$name.err
  python -c "import json;d=json.load(open('gpurun_out/r2f_$name.json'));print('$name: value', round(d['value'],1), 'compress', round(d['compress_MBps'],1), 'decompress', round(d['decompress_MBps'],1), 'ms/step', round(d['ms_per_step']))" || tail -5 gpurun_out/r2f_$name.err
}
{
echo "== 1. quick parity"
timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -m gpu -x -q 2>&1 | tail -3
echo "== 2. phased steps"
run phased48 X=1 -- --mode phased --blocks 48
run phased64 X=1 -- --mode phased --blocks 64
run phased96 X=1 -- --mode phased --blocks 96
echo "== 3. pipelined steps, coder CTAs capped"
run pipe64_cap160 BSCB200_CODER_SLOTS=160 -- --mode pipeline --blocks 64
run pipe64_cap208 BSCB200_CODER_SLOTS=208 -- --mode pipeline --blocks 64
run pipe64_cap256 BSCB200_CODER_SLOTS=256 -- --mode pipeline --blocks 64
run pipe64_nocap X=1 -- --mode pipeline --blocks 64
} 2>&1 | tee gpurun_out/r2_call_f.log
