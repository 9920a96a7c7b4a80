timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
for w in wd-articles; do
  timeout 300 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/b_$w.log 2> gpurun_out/b_$w.err; echo "$w rc=$?"
  tail -1 gpurun_out/b_$w.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['config']['workload'], round(d['ms_per_step'],4), round(d['value']), 'e2e', round(d['e2e']['value']), d['roofline']['kernel'], d['roofline'].get('traffic'))"
done
timeout 300 python bench.py --workload wd-articles --batch 256 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/b_wda256.log 2> gpurun_out/b_wda256.err; echo "b256 rc=$?"
tail -1 gpurun_out/b_wda256.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('b256', d['config']['workload'], round(d['ms_per_step'],4), round(d['value']), 'e2e', round(d['e2e']['value']), d['roofline']['kernel'])"
timeout 200 python tools/gru_cluster_check.py 512 16 3 212 2>&1 | grep -E "gru_|loss diff|grad rel"
