mkdir -p gpurun_out
ARK_GRU_PERSIST_DBG=1 timeout 120 python tools/gru_persist_bench.py 1024 256 10 2>&1 | tee gpurun_out/r2v_persist.log
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-library-baseline --no-also > gpurun_out/r2v_bench.json 2> gpurun_out/r2v_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2v_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','windows_ms','gpu_launches')})
PY
