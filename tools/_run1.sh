echo "=== d=512 B=16 nl=3 hi=212"
ARK_GRU_CLUSTER_DBG=300 timeout 200 python tools/gru_cluster_check.py 512 16 3 212 2>&1 | grep -v "rel " | grep -v "it+[23]" | grep -v "fwd proj\|fwd rec[12]\|bwd proj\|loader"
echo "=== d=128 B=256 nl=3 hi=23"
timeout 200 python tools/gru_cluster_check.py 128 256 3 23 2>&1 | grep -v "rel "
echo "=== tests"
timeout 900 python -m pytest tests/test_elbo_gpu.py -x -q -m gpu -k "gru_cluster or wavefront" 2>&1 | tail -5
