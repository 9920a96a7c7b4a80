mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "gather_pool or softmax" 2>&1 | tail -3
timeout 300 python tools/bench_hbm_kernels.py > gpurun_out/r02_hbm_kernels.jsonl 2> gpurun_out/r2z_hbm.err; echo "hbm rc=$?"; cat gpurun_out/r02_hbm_kernels.jsonl | cut -c1-170
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
