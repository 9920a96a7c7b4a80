export NCCL_DEBUG=WARN
ARK_CAPTURE_NCCL=1 timeout 90 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dp_check.py > gpurun_out/r2t_dpcheck.log 2>&1; echo "dp_check capture=1 rc=$?"; grep "DP_CHECK\|AssertionError\|Error" gpurun_out/r2t_dpcheck.log | head -6; tail -5 gpurun_out/r2t_dpcheck.log
