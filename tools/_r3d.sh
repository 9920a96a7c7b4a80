mkdir -p gpurun_out
export NCCL_DEBUG=WARN
for N in 2 4; do
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$N bench.py --gpus $N --steps 20 --warmup 5 --scaling strong --no-also --no-kernel-profile > gpurun_out/r3d_strong_n$N.json 2> gpurun_out/r3d_n$N.err; echo "strong n$N rc=$?"
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r3d_strong_n$N.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','scaling')}, d['config']['graphs_per_gpu'], d['config']['global_batch'], 'e2e', d['e2e']['value'])
except Exception as e:
    print('fail', e); print(open('gpurun_out/r3d_n$N.err').read()[-2000:])
PY
done
