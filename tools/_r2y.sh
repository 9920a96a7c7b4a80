for v in 0 51242 51243 51224 25644 25646 25628 25626 102441 102422; do echo "variant $v"; ARK_CE_VARIANT=$v timeout 60 python tools/bench_ce.py 2>&1 | tail -3; done | tee gpurun_out/r2y_ce.log
