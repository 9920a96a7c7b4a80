mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "gru_persist" 2>&1 | tail -5
for ks in 1 0; do ARK_GRU_KSPLIT=$ks ARK_GRU_PERSIST_DBG=1 timeout 120 python tools/gru_persist_bench.py 1024 256 10; done 2>&1 | tee gpurun_out/r2d_persist.log
ARK_GRU_KSPLIT=1 timeout 120 python tools/gru_persist_bench.py 1024 256 16 2>&1 | tee -a gpurun_out/r2d_persist.log
timeout 600 python -m pytest tests/test_baseline_shapes_gpu.py tests/test_elbo_gpu.py -x -q -m gpu 2>&1 | tail -3
