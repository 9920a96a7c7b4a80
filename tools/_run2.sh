timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -3
for w in syn-types; do
  timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/b_default.log 2> gpurun_out/b_default.err; echo "default rc=$?"
  tail -1 gpurun_out/b_default.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['config']['workload'], round(d['ms_per_step'],4), round(d['value']), 'e2e', round(d['e2e']['value']), d['roofline'], d.get('cpu_baseline'))"
done
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-600
