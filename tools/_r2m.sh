mkdir -p gpurun_out
export NCCL_DEBUG=WARN
for bm in 16 64; do
ARK_BUCKET_MB=$bm timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-also > gpurun_out/r2m_n2_$bm.json 2> gpurun_out/r2m_n2_$bm.err; echo "n2 bucket=$bm rc=$?"
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2m_n2_$bm.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','windows_ms','gpu_launches')}, 'e2e', d['e2e']['value'])
    for k in d['kernels'][:14]: print('  %-28s %.4f ms/step x%.1f'%(k['name'],k['ms_per_step'],k['launches_per_step']))
except Exception as e:
    print('fail', e); print(open('gpurun_out/r2m_n2_$bm.err').read()[-1500:])
PY
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/dp_check.py 2>&1 | tail -8
