#!/bin/bash
# tools/r2_call_e.sh -- round 2, fifth GPU call: kernel-side completion signal, single-walk inverse BWT, sort stall reasons
mkdir -p gpurun_out
{
echo "== 1. parity (inverse BWT rewritten, completion signal in every coder kernel)"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "== 2. pipeline bench"
for cfg in "48 48" "64 64"; do set -- $cfg
  timeout 400 python bench.py --blocks $1 --workers $2 --no-cpu-baseline --no-e2e --no-extras --steps 3 --warmup 2 > gpurun_out/r2e_bench_$1_$2.json 2> gpurun_out/r2e_bench_$1_$2.err
  python -c "import json;d=json.load(open('gpurun_out/r2e_bench_$1_$2.json'));print('blocks $1 in flight $2: value', round(d['value'],1), 'compress', round(d['compress_MBps'],1), 'decompress', round(d['decompress_MBps'],1), 'launches', d['gpu_launches'])" || tail -5 gpurun_out/r2e_bench_$1_$2.err
done
echo "== 3. standalone kernel table (one block at a time)"
python -c "
import json;d=json.load(open('gpurun_out/r2e_bench_48_48.json'))
for k in d['kernels_standalone'][:14]: print(k)
"
echo "== 4. where do the sort kernels wait?  (BWT of one 64 MiB block; 3 launches of each kernel)"
for t in 1 0; do
BSCB200_SORT_TMA=$t timeout 300 ncu --section SchedulerStats --section WarpStateStats --section SpeedOfLight --section Occupancy --section LaunchStats --section MemoryWorkloadAnalysis --section SourceCounters --import-source on --clock-control none -k "regex:rs_onesweep" -s 8 -c 3 -f -o gpurun_out/r2e_sort_tma$t python tools/one_block.py 64 > gpurun_out/r2e_ncu_sort$t.log 2>&1
done
ls -la gpurun_out/*.ncu-rep | tail -3
} 2>&1 | tee gpurun_out/r2_call_e.log
