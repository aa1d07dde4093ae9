#!/bin/bash
# tools/round2_first_gpu_call.sh -- everything that was prepared WITHOUT a GPU at the end of round 1, in the order it should
# meet one (gpurun --timeout 900 -- 'bash tools/round2_first_gpu_call.sh').  Each step is bounded by its own timeout and
# writes to gpurun_out/; nothing here changes defaults.  Expected: ~6-8 minutes.
mkdir -p gpurun_out
{
echo "== 1. default parity suite"
timeout 200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "== 2. adaptive + fast coders (coder ids 2, 3): parity through the C ABI, then lift the gates in qlfc.cu:coder_gate"
BSCB200_ENABLE_ADAPTIVE=1 BSCB200_ENABLE_FAST=1 timeout 300 python -m pytest tests/test_gpu_other_coders.py tests/test_golden.py -m gpu -q 2>&1 | tail -5
echo "== 2b. file-level front end (bsc1 container, multi-GPU scheduler) on the GPU"
BSCB200_TEST_CLI=1 BSCB200_TEST_LZP=1 BSCB200_ENABLE_LZP=1 timeout 300 python -m pytest tests/test_cli_container.py tests/test_gpu_parity.py -m gpu -q -k "cli or reference_default" 2>&1 | tail -3
echo "== 3. decoder A/B on one 64 MiB block: 4 = default, 7 = tuned code + full layout, 6 = tuned code + diet layout (2 streams/SM)"
timeout 200 python tools/dec_ab.py 64 4 7 6 2>&1 | tail -4
echo "== 3b. encoder variants: 2 = one-multiply-add range recurrence, 6 = diet counter file (two encoders per SM), 7 = both: parity + time"
for v in 2 6 7; do
  BSCB200_QENC=$v timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "coder_compress or block_bytes or k3" 2>&1 | tail -1
  BSCB200_QENC=$v timeout 100 python tools/dec_ab.py 64 4 2>&1 | tail -2
done
echo "== 4. parity of the diet decoder as the default decoder"
BSCB200_QDEC=6 timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
echo "== 5. does co-residency pay?  36 blocks per GPU (18 compression contexts of 4.6 GiB, 36 decode-only contexts of 1.4 GiB): default decoder vs diet decoder"
timeout 300 python bench.py --blocks 36 --workers 18 --decode-workers 36 --no-cpu-baseline --no-e2e --steps 2 > gpurun_out/r2_bench36_gen4.json 2> gpurun_out/r2_bench36_gen4.err; python -c "import json;d=json.load(open('gpurun_out/r2_bench36_gen4.json'));print('gen4 36 blocks', d['value'], d['compress_MBps'], d['decompress_MBps'])"
BSCB200_QDEC=6 timeout 300 python bench.py --blocks 36 --workers 18 --decode-workers 36 --no-cpu-baseline --no-e2e --steps 2 > gpurun_out/r2_bench36_gen6.json 2> gpurun_out/r2_bench36_gen6.err; python -c "import json;d=json.load(open('gpurun_out/r2_bench36_gen6.json'));print('gen6 36 blocks', d['value'], d['compress_MBps'], d['decompress_MBps'])"
} 2>&1 | tee gpurun_out/r2_first_call.log
{
echo "== 6. everything two-per-SM: 32 blocks in flight in both directions (32 contexts x 4.6 GiB), diet decoder + diet encoder with RANGE3"
BSCB200_QDEC=6 BSCB200_QENC=7 timeout 300 python bench.py --blocks 32 --no-cpu-baseline --no-e2e --steps 2 > gpurun_out/r2_bench32_diet.json 2> gpurun_out/r2_bench32_diet.err; python -c "import json;d=json.load(open('gpurun_out/r2_bench32_diet.json'));print('diet both, 32 blocks', d['value'], d['compress_MBps'], d['decompress_MBps'])"
} 2>&1 | tee -a gpurun_out/r2_first_call.log
echo "== 7. lone-warp issue rate (cost model of DESIGN.md 4.5)"; nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/warp_latency tools/warp_latency.cu && gpurun_out/warp_latency | tee -a gpurun_out/r2_first_call.log
echo "== 8. phase breakdown of the tuned decoder (cycles per run)"; for g in 7 6; do BSCB200_QDEC=$g BSCB200_QDEC_PROF=1 timeout 60 python tools/one_block.py 64 2>&1 | grep prof | tee -a gpurun_out/r2_first_call.log; done
