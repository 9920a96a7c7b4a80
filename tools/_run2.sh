set -x
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
for w in wd-articles wd-movies; do
  timeout 300 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/b_$w.log 2> gpurun_out/b_$w.err; echo "$w rc=$?"
  tail -1 gpurun_out/b_$w.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['config']['workload'], d['ms_per_step'], d['value'], 'e2e', d['e2e']['value'], d['roofline'])"
  cp gpurun_out/bench_breakdown_${w}_n1.json gpurun_out/bench_breakdown_${w}_n1_cluster.json 2>/dev/null
done
timeout 300 python bench.py --workload wd-articles --model ARK --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-400
