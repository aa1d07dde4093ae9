#!/bin/bash
# tools/r2_call_i.sh -- round 2, ninth GPU call: state tables through L1 (3 / 4 / 5 decoder streams per SM, 3 encoders per SM; sort at 4 CTAs per SM = 64 registers): parity, one block
# alone, pipeline A/B; then the full default bench with the winner and the reference arm at N = 1.
mkdir -p gpurun_out
run() { name=$1; shift; envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 500 python bench.py "$@" --no-cpu-baseline --no-e2e --no-extras --steps 3 --warmup 1 > gpurun_out/r2i_$name.json 2> gpurun_out/r2i_$name.err
  python -c "import json;d=json.load(open('gpurun_out/r2i_$name.json'));print('$name: value', round(d['value'],1), 'compress', round(d['compress_MBps'],1), 'decompress', round(d['decompress_MBps'],1), 'ms/step', round(d['ms_per_step']), 'rs_onesweep alone GB/s', (d.get('roofline_hbm_kernel') or {}).get('achieved'))" || tail -5 gpurun_out/r2i_$name.err
}
{
echo "== 1. parity at the new defaults (3 decoder / 3 encoder streams per SM, state tables through L1)"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
echo "== 1b. coder parity with 4 and 5 decoder streams per SM"
BSCB200_DEC_PER_SM=4 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "coder or block or golden or k2 or k3" 2>&1 | tail -3
BSCB200_DEC_PER_SM=5 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "coder or block" 2>&1 | tail -3
echo "== 2. one 64 MiB block alone: coder kernel times by layout"
for p in 2 3 4 5; do echo "-- DEC_PER_SM=$p ENC_TG=$([ $p = 2 ] && echo 0 || echo 1)"; BSCB200_DEC_PER_SM=$p BSCB200_ENC_TG=$([ $p = 2 ] && echo 0 || echo 1) timeout 300 python tools/dec_ab.py 64 2>&1 | tail -2; done
echo "== 3. pipeline A/B"
run base_2_64   BSCB200_DEC_PER_SM=2 BSCB200_ENC_TG=0 BSCB200_SORT_OCC=3 -- --blocks 64
run base_2_64_sortocc4 BSCB200_DEC_PER_SM=2 BSCB200_ENC_TG=0 -- --blocks 64
run dec3_enc2_64 BSCB200_DEC_PER_SM=3 BSCB200_ENC_TG=0 -- --blocks 64
run dec3_enc3_64 BSCB200_DEC_PER_SM=3 BSCB200_ENC_TG=1 -- --blocks 64
run dec3_enc3_96 BSCB200_DEC_PER_SM=3 BSCB200_ENC_TG=1 -- --blocks 96
run dec4_enc3_64 BSCB200_DEC_PER_SM=4 BSCB200_ENC_TG=1 -- --blocks 64
run dec4_enc3_96 BSCB200_DEC_PER_SM=4 BSCB200_ENC_TG=1 -- --blocks 96
run dec5_enc3_96 BSCB200_DEC_PER_SM=5 BSCB200_ENC_TG=1 -- --blocks 96
echo "== 4. full default bench (library defaults), N = 1"
SECONDS=0
timeout 1200 python bench.py --gpus 1 --steps 3 --warmup 3 > gpurun_out/r2i_bench_n1.json 2> gpurun_out/r2i_bench_n1.err; echo "bench wall: $SECONDS s"; tail -3 gpurun_out/r2i_bench_n1.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2i_bench_n1.json'))
print('N=1 value', round(d['value'],1), 'e2e', d['e2e'] and {k:(round(v,1) if isinstance(v,float) else v) for k,v in d['e2e'].items() if k!='pageable'}, 'pageable', d['e2e'] and d['e2e']['pageable'])
print('compress', round(d['compress_MBps'],1), 'decompress', round(d['decompress_MBps'],1), 'launches', d['gpu_launches'], 'clocks', d['clocks'])
print('cpu', d['cpu_baseline'])
print('roofline', d['roofline']); print('issue', d['roofline_issue']); print('hbm', d['roofline_hbm_kernel'])
for k,v in (d.get('other_configs') or {}).items(): print(k, {a:(round(b,1) if isinstance(b,float) else b) for a,b in v.items() if a!='workload'})
for k in d['kernels'][:12]: print(k)
PY
echo "== 5. reference arm, N = 1"
SECONDS=0
timeout 900 python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > gpurun_out/r2i_ref_n1.json 2> gpurun_out/r2i_ref_n1.err; echo "reference arm wall: $SECONDS s"
python -c "import json;d=json.load(open('gpurun_out/r2i_ref_n1.json'));print('reference arm N=1: value', round(d['value'],1), 'c', round(d['compress_MBps'],1), 'd', round(d['decompress_MBps'],1), d['cpu_baseline']['sample'])" || tail -5 gpurun_out/r2i_ref_n1.err
} 2>&1 | tee gpurun_out/r2_call_i.log
