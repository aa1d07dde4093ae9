#!/bin/bash
# tools/r2_call_l.sh -- round 2, twelfth GPU call: stream priorities (sorts highest, long coder streams middle, other coder streams lowest) A/B;
# decoder layout by load with hysteresis (no mix of 55 KB and 39 KB CTAs)
mkdir -p gpurun_out
run() { name=$1; shift; envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 500 python bench.py "$@" --blocks 64 --no-cpu-baseline --no-e2e --no-extras --steps 3 --warmup 1 > gpurun_out/r2l_$name.json 2> gpurun_out/r2l_$name.err
  python - <<PY || tail -5 gpurun_out/r2l_$name.err
import json
d=json.load(open('gpurun_out/r2l_$name.json'))
k={x['kernel']:x for x in d['kernels']}
def per(n): return round(k[n]['ms_total']/k[n]['launches'],2) if n in k else None
print('$name: value', round(d['value'],1), 'compress', round(d['compress_MBps'],1), 'decompress', round(d['decompress_MBps'],1), 'ms/step', round(d['ms_per_step']), 'in flight', d['arm']['blocks_in_flight_per_gpu'],
      '| per launch ms: decode', per('q_decode6'), 'encode', per('q_encode5'), 'onesweep', per('rs_onesweep'), 'lr_jump', per('lr_jump'))
PY
}
{
echo "== 1. coder / block parity with priorities on (default)"
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_golden.py tests/test_gpu_other_coders.py -m gpu -x -q -k "coder or block or golden or k2 or k3 or inplace" 2>&1 | tail -3
echo "== 2. pipeline A/B: priorities on / off, 96 and 128 in flight"
run prio_on_w96   X=1 -- --workers 96
run prio_off_w96  BSCB200_PRIO=0 -- --workers 96
run prio_on_w128  X=1 -- --workers 128
run prio_on_w96_dec4 BSCB200_DEC_PER_SM=4 -- --workers 96
run prio_on_w96_nosplit BSCB200_CODER_SPLIT=0 -- --workers 96
run prio_off_w96_again BSCB200_PRIO=0 -- --workers 96
run prio_on_w96_again X=1 -- --workers 96
} 2>&1 | tee gpurun_out/r2_call_l2.log
