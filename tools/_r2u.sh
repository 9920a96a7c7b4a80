mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-library-baseline > gpurun_out/r2u_bench.json 2> gpurun_out/r2u_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2u_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','windows_ms','gpu_launches')})
for a in d.get('also',[]): print('also',{k:a.get(k) for k in ('workload','graphs_per_gpu','value','ms_per_step','eager_fresh','error')}, (a.get('roofline') or {}).get('kernel'), (a.get('roofline') or {}).get('frac'))
PY
