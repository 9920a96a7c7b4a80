TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
for cfg in "1 1" "1 0" "0 0"; do
  set -- $cfg
  ARK_COMM_PRIORITY=$1 ARK_DP_HOLD_COMM=$2 timeout 300 $TR --master-port 29512 bench.py --gpus 2 --workload syn-types --steps 20 --warmup 5 --no-e2e > gpurun_out/n2p.log 2> gpurun_out/n2p.err; echo "prio=$1 hold=$2 rc=$?"
  tail -1 gpurun_out/n2p.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('N=2', d['config']['workload'], round(d['ms_per_step'],4), round(d['value']))"
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_breakdown_syn-types_n2.json'))
for k,v in d.items():
    if isinstance(v,dict):
        print("   ", {kk: round(vv['ms_per_step'],3) for kk,vv in v.items() if 'nccl' in kk or 'gru_persist' in kk or 'adam' in kk})
        break
PY
done
ARK_COMM_PRIORITY=1 timeout 300 $TR --master-port 29513 bench.py --gpus 2 --workload wd-articles --steps 20 --warmup 5 --no-e2e 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('N=2 prio=1', d['config']['workload'], round(d['ms_per_step'],4), round(d['value']))"
