#!/bin/bash
# Run at commit 4a20aea; kept as the record of how profiles/r2b_call_b.log was produced.
# tools/r2_call_b.sh -- round 2, second GPU call: inverse ST (new), reference CUDA path timing, co-residency / oversubscription benches.
mkdir -p gpurun_out
{
echo "== 1. inverse ST parity"
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "st_decode or st_blocks or k5" 2>&1 | tail -5
timeout 200 python -m pytest tests/test_golden.py -m gpu -x -q 2>&1 | tail -3
echo "== 2. reference CUDA path vs ours"
timeout 400 python tools/time_ref_cuda.py 3 > gpurun_out/r2_ref_cuda.json 2> gpurun_out/r2_ref_cuda.err; tail -12 gpurun_out/r2_ref_cuda.err
echo "== 3. 36 blocks per GPU, 18 compression contexts, 36 decode contexts: full-layout decoder (1 stream/SM, oversubscribed) vs diet (2/SM)"
for g in 7 6; do
  BSCB200_QDEC=$g timeout 300 python bench.py --blocks 36 --workers 18 --decode-workers 36 --no-cpu-baseline --no-e2e --steps 2 --warmup 1 > gpurun_out/r2_bench36_gen$g.json 2> gpurun_out/r2_bench36_gen$g.err
  python -c "import json;d=json.load(open('gpurun_out/r2_bench36_gen$g.json'));print('QDEC=$g 36 blocks value', d['value'], 'compress', d['compress_MBps'], 'decompress', d['decompress_MBps'])" || tail -3 gpurun_out/r2_bench36_gen$g.err
done
echo "== 4. 18 blocks baseline with the tuned decoder (QDEC=7)"
BSCB200_QDEC=7 timeout 300 python bench.py --no-cpu-baseline --no-e2e --steps 2 --warmup 1 > gpurun_out/r2_bench18_gen7.json 2> gpurun_out/r2_bench18_gen7.err
python -c "import json;d=json.load(open('gpurun_out/r2_bench18_gen7.json'));print('QDEC=7 18 blocks value', d['value'], 'compress', d['compress_MBps'], 'decompress', d['decompress_MBps'])" || tail -3 gpurun_out/r2_bench18_gen7.err
echo "== 5. diet encoder + diet decoder, 32 full contexts (2 per SM both directions)"
BSCB200_QDEC=6 BSCB200_QENC=6 timeout 300 python bench.py --blocks 32 --no-cpu-baseline --no-e2e --steps 2 --warmup 1 > gpurun_out/r2_bench32_diet.json 2> gpurun_out/r2_bench32_diet.err
python -c "import json;d=json.load(open('gpurun_out/r2_bench32_diet.json'));print('diet both, 32 blocks value', d['value'], 'compress', d['compress_MBps'], 'decompress', d['decompress_MBps'])" || tail -3 gpurun_out/r2_bench32_diet.err
} 2>&1 | tee gpurun_out/r2_call_b.log
