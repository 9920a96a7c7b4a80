mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-library-baseline --no-also > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2k_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','windows_ms','gpu_launches')})
for k in d['kernels'][:8]: print('  %-28s %.4f ms/step x%.1f  frac=%s'%(k['name'],k['ms_per_step'],k['launches_per_step'],k.get('frac')))
PY
