mkdir -p gpurun_out
python -m pytest tests/test_elbo_gpu.py tests/test_kernels_gpu.py -x -q -m gpu 2>&1 | tail -3
for cfg in "4 1" "6 1" "4 8" "6 8" "6 4" "6 2"; do set -- $cfg; ARK_GRU_STAGES=$1 ARK_GRU_CLUSTER=$2 ARK_GRU_PERSIST_DBG=1 timeout 120 python tools/gru_persist_bench.py 1024 256 10; done 2>&1 | tee gpurun_out/r2c_persist.log
