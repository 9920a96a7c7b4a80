mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "gemm" 2>&1 | tail -3
timeout 120 python tools/bench_gemm.py 2>&1 | tail -10 | tee gpurun_out/r3a_gemm.log
timeout 900 python -m pytest tests/test_baseline_shapes_gpu.py tests/test_elbo_gpu.py -x -q -m gpu 2>&1 | tail -3
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-library-baseline > gpurun_out/r3a_bench.json 2> gpurun_out/r3a_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r3a_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','windows_ms','gpu_launches')})
for k in d['kernels'][:12]: print('  %-28s %.4f ms/step x%.1f  frac=%s'%(k['name'],k['ms_per_step'],k['launches_per_step'],k.get('frac')))
for a in d.get('also',[]): print('also',{k:a.get(k) for k in ('workload','graphs_per_gpu','value','ms_per_step','eager_fresh','error')}, (a.get('roofline') or {}).get('kernel'), (a.get('roofline') or {}).get('frac'))
PY
