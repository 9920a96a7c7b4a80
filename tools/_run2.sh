timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
for w in wd-articles wd-movies syn-types; do
  timeout 300 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/b_$w.log 2> gpurun_out/b_$w.err; echo "$w rc=$?"
  tail -1 gpurun_out/b_$w.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['config']['workload'], round(d['ms_per_step'],4), round(d['value']), 'e2e', round(d['e2e']['value']), d['roofline']['kernel'], d['roofline'].get('traffic'))"
done
