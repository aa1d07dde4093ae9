#!/bin/bash
# tools/r2_call_h.sh -- round 2, eighth GPU call: parity at HEAD, full default bench at N = 1, reference arm, launch list
mkdir -p gpurun_out
{
echo "== 1. parity (list ranking fix, 128-byte slabs, 55 KB sort)"
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "== 2. inverse BWT kernel sum"
TRC_SIDE=ours TRC_STAGE=bwt_decode timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2h_launches_ours_bwt_decode.csv python tools/time_ref_cuda.py > /dev/null 2>&1
python tools/ncu_launch_sum.py gpurun_out/r2h_launches_ours_bwt_decode.csv 8
echo "== 3. default bench, N = 1 (what the driver runs, fewer steps)"
/usr/bin/time -v timeout 1200 python bench.py --gpus 1 --steps 3 --warmup 3 > gpurun_out/r2h_bench_n1.json 2> gpurun_out/r2h_bench_n1.err; grep -E "Elapsed|Maximum resident" gpurun_out/r2h_bench_n1.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2h_bench_n1.json'))
print('N=1 value', round(d['value'],1), 'e2e', d['e2e'] and {k:(round(v,1) if isinstance(v,float) else v) for k,v in d['e2e'].items() if k!='pageable'}, 'pageable', d['e2e'] and d['e2e']['pageable'])
print('compress', round(d['compress_MBps'],1), 'decompress', round(d['decompress_MBps'],1), 'launches', d['gpu_launches'], 'clocks', d['clocks'])
print('cpu', d['cpu_baseline'])
print('roofline', d['roofline']); print('issue', d['roofline_issue']); print('hbm', d['roofline_hbm_kernel'])
for k,v in (d.get('other_configs') or {}).items(): print(k, {a:(round(b,1) if isinstance(b,float) else b) for a,b in v.items() if a!='workload'})
for k in d['kernels'][:8]: print(k)
PY
echo "== 4. reference arm, N = 1"
/usr/bin/time -v timeout 900 python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > gpurun_out/r2h_ref_n1.json 2> gpurun_out/r2h_ref_n1.err; grep -E "Elapsed|Maximum resident" gpurun_out/r2h_ref_n1.err
python -c "import json;d=json.load(open('gpurun_out/r2h_ref_n1.json'));print('reference arm N=1: value', round(d['value'],1), 'c', round(d['compress_MBps'],1), 'd', round(d['decompress_MBps'],1), d['cpu_baseline']['sample'])" || tail -5 gpurun_out/r2h_ref_n1.err
echo "== 5. launch list of a short bench run (shares, not absolutes)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2h_launches_bench.csv python bench.py --blocks 8 --workers 8 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/r2h_bench_under_ncu.log 2>&1
python tools/ncu_launch_sum.py gpurun_out/r2h_launches_bench.csv 14
} 2>&1 | tee gpurun_out/r2_call_h1.log
