mkdir -p gpurun_out
for wl in syn-paths wd-movies; do for m in auto layer; do
ARK_GRU_MODE=$m python bench.py --workload $wl --steps 20 --warmup 5 --windows 3 --no-cpu-baseline --no-library-baseline --no-also --no-e2e > gpurun_out/r2x_${wl}_$m.json 2> gpurun_out/r2x.err; echo "$wl mode=$m rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r2x_${wl}_$m.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')}, [(k['name'], round(k['ms_per_step'],4)) for k in d['kernels'][:6]])
PY
done; done
