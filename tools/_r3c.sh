mkdir -p gpurun_out
export NCCL_DEBUG=WARN
N=$1
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29516 tools/dp_check.py 2>&1 | grep "DP_CHECK\|Error" | head -4
for bm in 48; do
ARK_BUCKET_MB=$bm timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus $N --steps 20 --warmup 5 --no-also > gpurun_out/r3c_n${N}_$bm.json 2> gpurun_out/r3c_n$N.err; echo "n$N bucket=$bm rc=$?"
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r3c_n${N}_$bm.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step')}, 'e2e', d['e2e']['value'])
    for k in d['kernels'][:6]: print('  %-28s %.4f ms/step x%.1f'%(k['name'],k['ms_per_step'],k['launches_per_step']))
except Exception as e:
    print('fail', e); print(open('gpurun_out/r3c_n$N.err').read()[-2500:])
PY
done
