mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "softmax" 2>&1 | tail -2
for v in 0 51242 51243; do echo "variant $v"; ARK_CE_VARIANT=$v timeout 60 python tools/bench_ce.py 2>&1 | tail -3; done
B="python bench.py --steps 2 --warmup 3 --windows 1 --no-e2e --no-cpu-baseline --no-library-baseline --no-also --no-graph --no-kernel-profile --workload wd-articles"
$B > gpurun_out/plain4.log 2>&1 && ncu --set full --clock-control none -k "regex:softmax_ce" -s 1 -c 2 -o gpurun_out/r02c_wda_ce $B > gpurun_out/ncu4.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/r02c_wda_ce.ncu-rep --page raw --csv > gpurun_out/r02c_wda_ce_raw.csv 2>/dev/null; rm -f gpurun_out/r02c_wda_ce.ncu-rep
