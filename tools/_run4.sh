TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29511 tools/dp_check.py > gpurun_out/dp7.log 2>&1; echo "dp_check rc=$?"; grep -E "DP_CHECK|Error|error|assert" gpurun_out/dp7.log | head -4
for w in syn-types wd-articles; do
  timeout 300 $TR --master-port 29512 bench.py --gpus 2 --workload $w --steps 20 --warmup 5 > gpurun_out/n2i_$w.log 2> gpurun_out/n2i_$w.err; echo "$w rc=$?"
  tail -1 gpurun_out/n2i_$w.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('N=2', d['config']['workload'], d['ms_per_step'], d['value'], 'e2e', d['e2e']['value'], d['roofline']['kernel'])"
done
