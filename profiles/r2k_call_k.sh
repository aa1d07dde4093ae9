#!/bin/bash
# tools/r2_call_k.sh -- round 2, eleventh GPU call: parity at HEAD (multi-hop list ranking, layouts by load, six sort slabs), sort-slab A/B,
# the full default bench and the reference arm at N = 1, launch list of a bench run, ncu of the dominant kernels.
mkdir -p gpurun_out
run() { name=$1; shift; envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 500 python bench.py "$@" --blocks 64 --no-cpu-baseline --no-e2e --no-extras --steps 3 --warmup 1 > gpurun_out/r2k_$name.json 2> gpurun_out/r2k_$name.err
  python -c "import json;d=json.load(open('gpurun_out/r2k_$name.json'));print('$name: value', round(d['value'],1), 'compress', round(d['compress_MBps'],1), 'decompress', round(d['decompress_MBps'],1), 'ms/step', round(d['ms_per_step']), 'in flight', d['config']['blocks_in_flight_per_gpu'], 'launches', d['gpu_launches'])" || tail -5 gpurun_out/r2k_$name.err
}
{
echo "== 1. parity at HEAD"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
echo "== 2. sort slabs per device: 3 / 6 / 9 (96 blocks in flight, layouts by load)"
run slabs3 BSCB200_SORT_SLABS=3 --
run slabs6 BSCB200_SORT_SLABS=6 --
run slabs9 BSCB200_SORT_SLABS=9 --
echo "== 3. full default bench, N = 1"
SECONDS=0
timeout 1200 python bench.py --gpus 1 --steps 3 --warmup 3 > gpurun_out/r2k_bench_n1.json 2> gpurun_out/r2k_bench_n1.err; echo "bench wall: $SECONDS s"; tail -3 gpurun_out/r2k_bench_n1.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2k_bench_n1.json'))
print('N=1 value', round(d['value'],1), 'e2e', d['e2e'] and {k:(round(v,1) if isinstance(v,float) else v) for k,v in d['e2e'].items() if k!='pageable'}, 'pageable', d['e2e'] and d['e2e']['pageable'])
print('compress', round(d['compress_MBps'],1), 'decompress', round(d['decompress_MBps'],1), 'launches', d['gpu_launches'], 'clocks', d['clocks'])
print('cpu', d['cpu_baseline'])
print('roofline', {k:v for k,v in d['roofline'].items() if k!='note'}); print('hbm', d['roofline_hbm_kernel'])
for k,v in (d.get('other_configs') or {}).items(): print(k, {a:(round(b,1) if isinstance(b,float) else b) for a,b in v.items() if a!='workload'})
for k in d['kernels'][:8]: print(k)
PY
echo "== 4. reference arm, N = 1 (16-block sample)"
SECONDS=0
timeout 900 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/r2k_ref_n1.json 2> gpurun_out/r2k_ref_n1.err; echo "reference arm wall: $SECONDS s"
python -c "import json;d=json.load(open('gpurun_out/r2k_ref_n1.json'));print('reference arm N=1: value', round(d['value'],1), 'c', round(d['compress_MBps'],1), 'd', round(d['decompress_MBps'],1), d['cpu_baseline']['sample'])" || tail -5 gpurun_out/r2k_ref_n1.err
echo "== 5. launch list of a short bench run (shares, not absolutes)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2k_launches_bench.csv python bench.py --blocks 8 --workers 8 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/r2k_bench_under_ncu.log 2>&1
python tools/ncu_launch_sum.py gpurun_out/r2k_launches_bench.csv 14
echo "== 6. ncu: the coder kernels at HEAD on a 4 MiB block (decoder LayoutDiet5 and LayoutDiet4, encoder), then one rs_onesweep pass of a 64 MiB block with --set full"
for p in 5 4; do
BSCB200_DEC_PER_SM=$p timeout 300 ncu --section SchedulerStats --section WarpStateStats --section SourceCounters --section LaunchStats --section Occupancy --section SpeedOfLight --section InstructionStats --section MemoryWorkloadAnalysis --clock-control none --import-source on -k "regex:q_(de|en)code" -c 2 -f -o gpurun_out/r2k_coder_dec$p python tools/one_block.py 4 > gpurun_out/r2k_ncu_coder$p.log 2>&1
done
timeout 300 ncu --set full --clock-control none --import-source on -k "regex:rs_onesweep" -s 8 -c 2 -f -o gpurun_out/r2k_sort_full python tools/one_block.py 64 > gpurun_out/r2k_ncu_sort.log 2>&1
timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active.avg.per_cycle_active,sm__cycles_active.avg,gpu__time_duration.sum --clock-control none -k "regex:q_(de|en)code" --csv --log-file gpurun_out/r2k_coder_metrics_64MiB.csv python tools/one_block.py 64 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep | tail -4
} 2>&1 | tee gpurun_out/r2_call_k.log
