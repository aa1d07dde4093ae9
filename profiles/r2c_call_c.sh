#!/bin/bash
# tools/r2_call_c.sh -- round 2, third GPU call: pruned library (diet coders only, scratch pool, gates lifted): full parity, pipeline bench, reference CUDA kernels
mkdir -p gpurun_out
{
echo "== 1. full parity suite"
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "== 2. pipeline bench: blocks / contexts in flight"
for cfg in "48 48" "64 64" "48 32"; do set -- $cfg
  timeout 400 python bench.py --blocks $1 --workers $2 --no-cpu-baseline --no-e2e --steps 3 --warmup 2 > gpurun_out/r2c_bench_$1_$2.json 2> gpurun_out/r2c_bench_$1_$2.err
  python -c "import json;d=json.load(open('gpurun_out/r2c_bench_$1_$2.json'));print('blocks $1 in flight $2: value', round(d['value'],1), 'compress', round(d['compress_MBps'],1), 'decompress', round(d['decompress_MBps'],1), 'launches', d['gpu_launches'])" || tail -5 gpurun_out/r2c_bench_$1_$2.err
done
echo "== 3. reference CUDA path vs ours: wall clock, then kernel-only sums (ncu launch lists of ONE call each)"
timeout 400 python tools/time_ref_cuda.py 3 > gpurun_out/r2c_ref_cuda.json 2> gpurun_out/r2c_ref_cuda.err; tail -9 gpurun_out/r2c_ref_cuda.err
for side in ref ours; do for stage in bwt_encode bwt_decode st6_encode; do
  TRC_SIDE=$side TRC_STAGE=$stage timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2c_launches_${side}_$stage.csv python tools/time_ref_cuda.py > /dev/null 2>&1
  python tools/ncu_launch_sum.py gpurun_out/r2c_launches_${side}_$stage.csv 6
done; done
echo "== 4. coder kernels on a 64 MiB block: DRAM bytes, instructions, issue rate"
timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active.avg.per_cycle_active,sm__cycles_active.avg,gpu__time_duration.sum --clock-control none -k "regex:q_(de|en)code" --csv --log-file gpurun_out/r2c_coder_metrics_64MiB.csv python tools/one_block.py 64 > /dev/null 2>&1
cat gpurun_out/r2c_coder_metrics_64MiB.csv | grep -v "^==" | cut -d, -f5,13- | head -20
} 2>&1 | tee gpurun_out/r2_call_c.log
