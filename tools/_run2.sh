timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu 2>&1 | tail -3
for ts in 0 1; do
for w in wd-articles wd-movies; do
  ARK_GEMM_TMA_STORE=$ts timeout 300 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/bt_$w.log 2> gpurun_out/bt_$w.err; echo "tma_store=$ts $w rc=$?"
  tail -1 gpurun_out/bt_$w.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['config']['workload'], round(d['ms_per_step'],4), round(d['value']))"
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_breakdown_${w}_n1.json'))
for k,v in d.items():
    if isinstance(v,dict):
        print("   ", {kk: round(vv['ms_per_step'],3) for kk,vv in v.items() if 'vocab' in kk})
        break
PY
done
done
