#!/bin/bash
# tools/round2_first_gpu_call.sh -- everything that was prepared WITHOUT a GPU at the end of round 1, in the order it should
# meet one (gpurun --timeout 1500 -- 'bash tools/round2_first_gpu_call.sh'; or run it in two halves: steps 0-4, then 5-9).  Each step is
# bounded by its own timeout and writes to gpurun_out/; nothing here changes defaults.  Expected: ~15 minutes.
mkdir -p gpurun_out
{
echo "== 1. default parity suite"
timeout 200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "== 2. adaptive + fast coders (coder ids 2, 3): parity through the C ABI, then lift the gates in qlfc.cu:coder_gate"
BSCB200_ENABLE_ADAPTIVE=1 BSCB200_ENABLE_FAST=1 timeout 300 python -m pytest tests/test_gpu_other_coders.py tests/test_golden.py -m gpu -q 2>&1 | tail -5
echo "== 2b. file-level front end (bsc1 container, multi-GPU scheduler) on the GPU"
BSCB200_TEST_CLI=1 BSCB200_TEST_LZP=1 BSCB200_ENABLE_LZP=1 timeout 300 python -m pytest tests/test_cli_container.py tests/test_gpu_parity.py -m gpu -q -k "cli or reference_default" 2>&1 | tail -3
echo "== 0. lone-warp microbenchmarks FIRST (seconds): issue rate, branch cost against code footprint (cases 120, 124-126), LDS.U16"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/warp_latency tools/warp_latency.cu && timeout 60 gpurun_out/warp_latency
python tools/fetch_verdict.py gpurun_out/r2_first_call.log 2>&1 | tail -20
echo "== 3. decoder A/B on one 64 MiB block: 4 = default, 7 = tuned code + full layout, 6 = tuned code + diet layout (2 streams/SM), 8 / 9 = 7 / 6 with rolled decision loops (small code footprint)"
timeout 300 python tools/dec_ab.py 64 4 7 8 6 9 2>&1 | tail -6
echo "== 3a. parity of the decoder variants on the small-input suite"
for g in 7 8 9; do BSCB200_QDEC=$g timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "decompress or cross_decoding or k2 or k3" 2>&1 | tail -1; done
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
echo "== 8. phase breakdown of the tuned decoder (cycles per run)"; for g in 7 8 6; do BSCB200_QDEC=$g BSCB200_QDEC_PROF=1 timeout 60 python tools/one_block.py 64 2>&1 | grep prof | tee -a gpurun_out/r2_first_call.log; done
echo "== 9. where do the coder warps wait (encoder + decoder)?  ncu --set full on a SMALL block (4 MiB: the coder kernel runs ~0.1 s, 40 replays fit) for the default and the rolled decoder"
for g in 4 8; do
  BSCB200_QDEC=$g timeout 280 ncu --set full --clock-control none --import-source on -k "regex:q_(de|en)code" -c 2 -f -o gpurun_out/r2_qdec_gen$g python tools/one_block.py 4 > gpurun_out/r2_ncu_gen$g.log 2>&1
  ncu -i gpurun_out/r2_qdec_gen$g.ncu-rep --page details --csv 2>/dev/null | grep -i -E "stall|no instruction|issued|ipc|branch" | cut -c1-200 | head -40 | tee -a gpurun_out/r2_first_call.log
done
