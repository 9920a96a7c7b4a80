"""Times the reference's library path (torch eager on cuda: cuDNN GRU, cuBLASLt, ATen) on every BASELINE workload.
python tools/library_baseline.py [workload ...]   -> one JSON line per (workload, precision)"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ark_b200.synthetic import model_config, synth_batch  # noqa: E402
from oracle.torch_cpu_port import time_cuda_library_baseline  # noqa: E402

for w in (sys.argv[1:] or ["syn-paths", "syn-types", "syn-tipr", "wd-movies", "wd-articles"]):
    cfg = model_config(w)
    hb = [synth_batch(cfg, cfg["batch_size"], 1234 + i) for i in range(4)]
    for tf32 in (False, True):
        r = time_cuda_library_baseline(cfg, [(t, s) for t, s, _ in hb], [n for _, _, n in hb], tf32=tf32)
        r["workload"] = w
        print(json.dumps(r), flush=True)
        torch.cuda.empty_cache()
