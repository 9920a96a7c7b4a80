TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
n=8
for ctas in 16 8; do
for w in syn-types wd-articles; do
  NCCL_MAX_CTAS=$ctas timeout 240 $TR --nproc-per-node $n --master-port 2952$ctas bench.py --gpus $n --workload $w --steps 20 --warmup 5 > gpurun_out/n${n}k${ctas}_$w.log 2> gpurun_out/n${n}k${ctas}_$w.err; echo "N=$n ctas=$ctas $w rc=$?"
  tail -1 gpurun_out/n${n}k${ctas}_$w.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('N=$n ctas=$ctas', d['config']['workload'], round(d['ms_per_step'],4), round(d['value']), 'e2e', round(d['e2e']['value']))"
  cp gpurun_out/bench_breakdown_${w}_n${n}.json gpurun_out/bench_breakdown_${w}_n${n}_k${ctas}.json 2>/dev/null
done
done
NCCL_MAX_CTAS=16 timeout 240 $TR --nproc-per-node 8 --master-port 29533 bench.py --gpus 8 --workload wd-articles --batch 256 --steps 10 --warmup 3 > gpurun_out/n8_wda256.log 2> gpurun_out/n8_wda256.err; echo "N=8 b256 rc=$?"
tail -1 gpurun_out/n8_wda256.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('N=8 b256', d['config']['workload'], round(d['ms_per_step'],4), round(d['value']), 'e2e', round(d['e2e']['value']))"
