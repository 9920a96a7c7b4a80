export NCCL_DEBUG=WARN
for c in 1 0; do ARK_CAPTURE_NCCL=$c timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$c tools/dp_check.py > gpurun_out/r2q_dpcheck_$c.log 2>&1; echo "dp_check capture=$c rc=$?"; grep "DP_CHECK\|AssertionError\|Mismatched\|Greatest" gpurun_out/r2q_dpcheck_$c.log | head -8; done
