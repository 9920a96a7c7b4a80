mkdir -p gpurun_out
export NCCL_DEBUG=WARN
N=$1
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r3b_n$N.json 2> gpurun_out/r3b_n$N.err; echo "n$N rc=$?"
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r3b_n$N.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','windows_ms','gpu_launches')}, 'e2e', d['e2e']['value'], d['final_loss'])
    for k in d['kernels'][:8]: print('  %-28s %.4f ms/step x%.1f'%(k['name'],k['ms_per_step'],k['launches_per_step']))
    for a in d.get('also',[]): print('also',{k:a.get(k) for k in ('workload','graphs_per_gpu','n_gpus','value','ms_per_step','error')})
except Exception as e:
    print('fail', e); print(open('gpurun_out/r3b_n$N.err').read()[-2500:])
PY
