#!/bin/bash
# tools/r2_call_j.sh -- round 2, tenth GPU call: encoders really three / four per SM (96 / 80 registers; at 122 only two fitted whatever the shared memory),
# decoder layout by load, blocks in flight 96 / 128 with 64 blocks per step, blocking stream waits A/B.
mkdir -p gpurun_out
run() { name=$1; shift; envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 500 python bench.py "$@" --blocks 64 --no-cpu-baseline --no-e2e --no-extras --steps 3 --warmup 1 > gpurun_out/r2j_$name.json 2> gpurun_out/r2j_$name.err
  python -c "import json;d=json.load(open('gpurun_out/r2j_$name.json'));print('$name: value', round(d['value'],1), 'compress', round(d['compress_MBps'],1), 'decompress', round(d['decompress_MBps'],1), 'ms/step', round(d['ms_per_step']), 'in flight', d['config']['blocks_in_flight_per_gpu'])" || tail -5 gpurun_out/r2j_$name.err
}
{
echo "== 1. coder / block parity: defaults (decoder layout by load, three encoders per SM), then four encoders per SM"
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -m gpu -x -q -k "coder or block or golden or k2 or k3 or inplace" 2>&1 | tail -3
BSCB200_ENC_PER_SM=4 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -m gpu -x -q -k "coder or block or golden or k2 or k3" 2>&1 | tail -3
echo "== 2. one 64 MiB block alone: encoder at 122 / 96 / 80 registers"
for p in 2 3 4; do echo "-- ENC_PER_SM=$p"; BSCB200_ENC_PER_SM=$p timeout 300 python tools/dec_ab.py 64 2>&1 | tail -2 | head -1; done
echo "== 3. pipeline A/B, 64 blocks per step"
run auto_enc3_w96   X=1 -- --workers 96
run dec5_enc3_w96   BSCB200_DEC_PER_SM=5 -- --workers 96
run dec5_enc4_w96   BSCB200_DEC_PER_SM=5 BSCB200_ENC_PER_SM=4 -- --workers 96
run dec5_enc4_w128  BSCB200_DEC_PER_SM=5 BSCB200_ENC_PER_SM=4 -- --workers 128
run dec5_enc3_w128  BSCB200_DEC_PER_SM=5 -- --workers 128
run dec5_enc3_w96_spin BSCB200_DEC_PER_SM=5 BSCB200_SYNC=spin -- --workers 96
run dec4_enc4_w96   BSCB200_DEC_PER_SM=4 BSCB200_ENC_PER_SM=4 -- --workers 96
} 2>&1 | tee gpurun_out/r2_call_j.log
