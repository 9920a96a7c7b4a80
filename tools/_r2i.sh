mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "gru_persist" 2>&1 | tail -5
for h2 in 1 0; do ARK_GRU_H2=$h2 ARK_GRU_DEBUG=1 ARK_GRU_PERSIST_DBG=1 timeout 120 python tools/gru_persist_bench.py 1024 256 10; done 2>&1 | tee gpurun_out/r2i_persist.log
