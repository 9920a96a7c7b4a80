for cl in 1 2 4 8; do ARK_GRU_DEBUG=1 ARK_GRU_KSPLIT=0 ARK_GRU_CLUSTER=$cl timeout 120 python tools/gru_persist_bench.py 1024 256 10 2>&1 | grep -v "^  " | tail -4; done
