#!/bin/bash
# tools/r2_call_n.sh -- round 2, last GPU call: the whole GPU parity suite at HEAD (stream priorities on, decoder layout with hysteresis), smoke(), then
# the bench exactly as the driver launches it (--gpus 1 --steps 20 --warmup 5)
mkdir -p gpurun_out
{
echo "== 1. pytest -m gpu at HEAD"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "== 2. smoke()"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== 3. python bench.py --gpus 1 --steps 20 --warmup 5"
SECONDS=0
timeout 860 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2n_bench_n1.json 2> gpurun_out/r2n_bench_n1.err; echo "rc $? bench wall: $SECONDS s"; tail -3 gpurun_out/r2n_bench_n1.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2n_bench_n1.json'))
print('N=1 value', round(d['value'],1), 'ms/step', round(d['ms_per_step']), 'e2e', d['e2e'] and {k:(round(v,1) if isinstance(v,float) else v) for k,v in d['e2e'].items() if k!='pageable'}, 'pageable', d['e2e'] and d['e2e']['pageable'])
print('compress', round(d['compress_MBps'],1), 'decompress', round(d['decompress_MBps'],1), 'launches', d['gpu_launches'], 'clocks', d['clocks'])
print('cpu', d['cpu_baseline'])
print('roofline', {k:v for k,v in d['roofline'].items() if k!='note'}); print('hbm', d['roofline_hbm_kernel'])
for k,v in (d.get('other_configs') or {}).items(): print(k, {a:(round(b,1) if isinstance(b,float) else b) for a,b in v.items() if a!='workload'})
for k in d['kernels'][:8]: print(k)
PY
} 2>&1 | tee gpurun_out/r2_call_n.log
