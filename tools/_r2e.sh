mkdir -p gpurun_out
python bench.py --steps 20 --warmup 5 --no-also --no-cpu-baseline --no-library-baseline > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2e_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','windows_ms','gpu_launches')})
for k in d['kernels']: print('  %-28s %.4f ms/step x%.1f  frac=%s'%(k['name'],k['ms_per_step'],k['launches_per_step'],k.get('frac')))
print('e2e',d.get('e2e'))
PY
