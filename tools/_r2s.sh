mkdir -p gpurun_out
export NCCL_DEBUG=WARN
for c in 1 0; do ARK_CAPTURE_NCCL=$c timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$c tools/dp_check.py > gpurun_out/r2s_dpcheck_$c.log 2>&1; echo "dp_check capture=$c rc=$?"; grep "DP_CHECK\|AssertionError\|Mismatched\|Greatest\|Error" gpurun_out/r2s_dpcheck_$c.log | head -6; done
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 --no-also > gpurun_out/r2s_n2.json 2> gpurun_out/r2s_n2.err; echo "n2 rc=$?"
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2s_n2.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','windows_ms','gpu_launches')}, 'e2e', d['e2e']['value'], d['final_loss'])
    for k in d['kernels'][:8]: print('  %-28s %.4f ms/step x%.1f'%(k['name'],k['ms_per_step'],k['launches_per_step']))
except Exception as e:
    print('fail', e); print(open('gpurun_out/r2s_n2.err').read()[-2500:])
PY
