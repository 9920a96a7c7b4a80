set -x
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/r2b_gpu.log 2>&1; echo "gpu rc=$?"
tail -25 gpurun_out/r2b_gpu.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/r2b_bench.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2b_bench.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','windows_ms','gpu_launches')})
    print('roofline',d['roofline'])
    for k in d['kernels']: print('  %-28s %.4f ms/step x%.1f  frac=%s'%(k['name'],k['ms_per_step'],k['launches_per_step'],k.get('frac')))
    print('e2e',d.get('e2e')); print('lib',d.get('library_baseline')); print('cpu',d.get('cpu_baseline'))
    for a in d.get('also',[]): print('also',{k:a.get(k) for k in ('workload','graphs_per_gpu','value','ms_per_step','eager_fresh','error')}, (a.get('roofline') or {}).get('kernel'), (a.get('roofline') or {}).get('frac'))
except Exception as e: print('parse failed',e)
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2b_smoke_launches.csv python -c "
import os
print('ENV', sorted(k for k in os.environ if any(s in k for s in ('NV','CUDA','NSIGHT','INJECT'))))
import __graft_entry__ as g
g.smoke()
" > gpurun_out/r2b_smoke_ncu.log 2>&1; echo "ncu smoke rc=$?"; tail -4 gpurun_out/r2b_smoke_ncu.log; grep -c gru_cluster gpurun_out/r2b_smoke_launches.csv
