#!/bin/bash
# tools/r2_call_m.sh -- round 2, thirteenth GPU call (TWO GPUs): the default bench and the reference arm under torchrun at N = 2, as the driver launches them
mkdir -p gpurun_out
{
echo "== 1. bench.py, N = 2 (torchrun), defaults except --steps 2 --warmup 1"
SECONDS=0
timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 1 > gpurun_out/r2m_bench_n2.json 2> gpurun_out/r2m_bench_n2.err; echo "rc $? wall $SECONDS s"; grep -v "^  File\|^    \|Warning\|warn" gpurun_out/r2m_bench_n2.err | tail -8
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2m_bench_n2.json') if l.startswith('{')][-1])
print('N=2 value', round(d['value'],1), 'n_gpus', d['n_gpus'], 'e2e', d['e2e'] and round(d['e2e']['value'],1), 'pageable', d['e2e'] and round(d['e2e']['pageable']['value'],1), 'compress', round(d['compress_MBps'],1), 'decompress', round(d['decompress_MBps'],1), 'launches', d['gpu_launches'])
print('config', d['config']); print('arm', d['arm']); print('clocks', d['clocks'])
for k,v in (d.get('other_configs') or {}).items(): print(k, {a:(round(b,1) if isinstance(b,float) else b) for a,b in v.items() if a!='workload'})
PY
echo "== 2. reference arm, N = 2 (torchrun; rank 0 alone works)"
SECONDS=0
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r2m_ref_n2.json 2> gpurun_out/r2m_ref_n2.err; echo "rc $? wall $SECONDS s"
python -c "import json;d=json.loads([l for l in open('gpurun_out/r2m_ref_n2.json') if l.startswith('{')][-1]);print('reference arm N=2: value', round(d['value'],1), d['config'], d['arm'], d['cpu_baseline']['sample'])" || tail -5 gpurun_out/r2m_ref_n2.err
} 2>&1 | tee gpurun_out/r2_call_m.log
