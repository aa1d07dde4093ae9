#!/bin/bash
# tools/r2_call_d.sh -- round 2, fourth GPU call: TMA-staged sort + ballot ranking, polling wait (no work queued behind coder kernels), new tests
mkdir -p gpurun_out
{
echo "== 0. quick parity with the TMA sort (bounded: a hang here must not cost the call)"
timeout 180 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "bwt_encode_matches or st_encode_matches or st_tiny" 2>&1 | tail -3
echo "== 0b. the same without TMA (ballot ranking only)"
BSCB200_SORT_TMA=0 timeout 180 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "bwt_encode_matches or st_encode_matches or st_tiny" 2>&1 | tail -3
echo "== 1. full parity suite"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "== 2. sort A/B on one block (kernel table of one 64 MiB block, standalone)"
for t in 1 0; do BSCB200_SORT_TMA=$t timeout 200 python bench.py --blocks 2 --workers 2 --no-cpu-baseline --no-e2e --steps 1 --warmup 1 > gpurun_out/r2d_sort_tma$t.json 2>gpurun_out/r2d_sort_tma$t.err
python -c "
import json;d=json.load(open('gpurun_out/r2d_sort_tma$t.json'))
for k in d['kernels_standalone']:
    if k['kernel'].startswith(('rs_','unbwt','bwt_')): print('TMA=$t', k)
" || tail -3 gpurun_out/r2d_sort_tma$t.err; done
echo "== 3. pipeline bench"
for cfg in "48 48" "64 64" "40 40"; do set -- $cfg
  timeout 400 python bench.py --blocks $1 --workers $2 --no-cpu-baseline --no-e2e --steps 3 --warmup 2 > gpurun_out/r2d_bench_$1_$2.json 2> gpurun_out/r2d_bench_$1_$2.err
  python -c "import json;d=json.load(open('gpurun_out/r2d_bench_$1_$2.json'));print('blocks $1 in flight $2: value', round(d['value'],1), 'compress', round(d['compress_MBps'],1), 'decompress', round(d['decompress_MBps'],1), 'launches', d['gpu_launches'])" || tail -5 gpurun_out/r2d_bench_$1_$2.err
done
} 2>&1 | tee gpurun_out/r2_call_d.log
