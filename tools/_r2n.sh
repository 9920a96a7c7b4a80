mkdir -p gpurun_out
export NCCL_DEBUG=WARN
ARK_CAPTURE_NCCL=1 timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 5 --no-also --no-kernel-profile > gpurun_out/r2n_n2.json 2> gpurun_out/r2n_n2.err; echo "n2 capture_nccl rc=$?"
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2n_n2.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','windows_ms','gpu_launches')}, 'e2e', d['e2e']['value'], d['final_loss'])
except Exception as e:
    print('fail', e); print(open('gpurun_out/r2n_n2.err').read()[-2500:])
PY
ARK_CAPTURE_NCCL=1 timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 tools/dp_check.py 2>&1 | tail -4
