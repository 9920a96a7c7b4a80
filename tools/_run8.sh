TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
i=0
for hold in 0 1; do
  i=$((i+1))
  ARK_DP_HOLD_COMM=$hold timeout 200 $TR --master-port 2953$i bench.py --gpus 8 --workload syn-types --steps 20 --warmup 5 --no-e2e > gpurun_out/n8p_$hold.log 2> gpurun_out/n8p_$hold.err; echo "hold=$hold rc=$?"
  tail -1 gpurun_out/n8p_$hold.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('N=8 hold=$hold', d['config']['workload'], round(d['ms_per_step'],4), round(d['value']))"
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_breakdown_syn-types_n8.json'))
for k,v in d.items():
    if isinstance(v,dict):
        print("   ", {kk: round(vv['ms_per_step'],3) for kk,vv in v.items() if 'nccl' in kk or 'gru_persist' in kk or 'adam' in kk})
        break
PY
done
timeout 200 $TR --master-port 29539 bench.py --gpus 8 --workload wd-articles --steps 20 --warmup 5 > gpurun_out/n8p_wda.log 2> gpurun_out/n8p_wda.err; echo "wda rc=$?"
tail -1 gpurun_out/n8p_wda.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('N=8', d['config']['workload'], round(d['ms_per_step'],4), round(d['value']), 'e2e', round(d['e2e']['value']))"
