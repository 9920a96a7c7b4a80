mkdir -p gpurun_out
timeout 300 python tools/bench_hbm_kernels.py > gpurun_out/r02_hbm_kernels.jsonl 2> gpurun_out/r3g_hbm.err; echo "hbm rc=$?"; cut -c1-120 gpurun_out/r02_hbm_kernels.jsonl
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -4
python bench.py --steps 20 --warmup 5 > gpurun_out/r3g_bench.json 2> gpurun_out/r3g_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r3g_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','windows_ms','gpu_launches','clocks')})
print('roofline', d['roofline'])
for k in d['kernels'][:16]: print('  %-28s %.4f ms/step x%.1f  frac=%s traffic=%s'%(k['name'],k['ms_per_step'],k['launches_per_step'],k.get('frac'),k.get('traffic')))
print('e2e',d.get('e2e')); print('lib',d.get('library_baseline')); print('cpu',d.get('cpu_baseline'))
for a in d.get('also',[]): print('also',{k:a.get(k) for k in ('workload','graphs_per_gpu','value','ms_per_step','eager_fresh','error')}, (a.get('roofline') or {}).get('kernel'), (a.get('roofline') or {}).get('frac'))
PY
