#!/bin/bash
# Run at commit 4a20aea (BSCB200_QDEC / BSCB200_QENC selected kernel generations that have since been removed); kept as the record of how profiles/r2a_call_a.log was produced.
# tools/r2_call_a.sh -- round 2, first GPU call: never-GPU-verified code first (coders 2/3, CLI, LZP, decoder generations), then stall-reason ncu.
mkdir -p gpurun_out
{
echo "== 1. default parity suite"
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "== 2. adaptive + fast coders"
BSCB200_ENABLE_ADAPTIVE=1 BSCB200_ENABLE_FAST=1 timeout 300 python -m pytest tests/test_gpu_other_coders.py tests/test_golden.py -m gpu -q 2>&1 | tail -8
echo "== 2b. CLI + LZP on the GPU"
BSCB200_TEST_CLI=1 BSCB200_TEST_LZP=1 BSCB200_ENABLE_LZP=1 timeout 300 python -m pytest tests/test_cli_container.py tests/test_gpu_parity.py -m gpu -q -k "cli or reference_default or lzp" 2>&1 | tail -8
echo "== 0. lone-warp microbenchmarks"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/warp_latency tools/warp_latency.cu && timeout 60 gpurun_out/warp_latency
echo "== 3. decoder A/B on one 64 MiB block"
timeout 400 python tools/dec_ab.py 64 4 7 8 6 9 2>&1 | tail -12
echo "== 3a. parity of decoder variants"
for g in 7 8 9 6; do echo "QDEC=$g"; BSCB200_QDEC=$g timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "decompress or cross_decoding or k2 or k3" 2>&1 | tail -1; done
echo "== 3b. encoder variants"
for v in 2 6 7; do
  echo "QENC=$v"
  BSCB200_QENC=$v timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "coder_compress or block_bytes or k3" 2>&1 | tail -1
  BSCB200_QENC=$v timeout 100 python tools/dec_ab.py 64 4 2>&1 | tail -2
done
echo "== 8. phase breakdown (cycles per run)"
for g in 4 7 8 6; do BSCB200_QDEC=$g BSCB200_QDEC_PROF=1 timeout 60 python tools/one_block.py 64 2>&1 | grep prof; done
echo "== 9. ncu stall reasons on a 4 MiB block (default decoder + encoder, then rolled decoder)"
for g in 4 8; do
  BSCB200_QDEC=$g timeout 240 ncu --section SchedulerStats --section WarpStateStats --section SourceCounters --section LaunchStats --section Occupancy --section SpeedOfLight --section InstructionStats --clock-control none --import-source on -k "regex:q_(de|en)code" -c 2 -f -o gpurun_out/r2_qdec_gen$g python tools/one_block.py 4 > gpurun_out/r2_ncu_gen$g.log 2>&1
  ncu -i gpurun_out/r2_qdec_gen$g.ncu-rep --page details --csv 2>/dev/null | grep -i -E "stall|no instruction|issued|ipc|branch|eligible|cycles per" | cut -c1-220 | head -60
done
} 2>&1 | tee gpurun_out/r2_call_a.log
