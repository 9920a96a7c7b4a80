#!/usr/bin/env python
"""profiles/ncu_traffic.json from `ncu --set full` raw-page CSVs: mean DRAM bytes (read + write) per launch of the
kernels bench.py can name.   python tools/ncu_traffic.py workload=csv [workload=csv ...]"""
import collections
import csv
import json
import re
import sys

TAGS = [("gru_persist_fwd", "gru_persist_fwd"), ("gru_persist_bwd", "gru_persist_bwd"), ("gru_cluster_fwd", "gru_cluster_fwd"),
        ("gru_cluster_bwd", "gru_cluster_bwd"), ("gru_wave_fwd", "gru_wave_fwd"), ("gru_wave_bwd", "gru_wave_bwd"),
        ("softmax_ce", "softmax_ce"), ("gather_pool_bwd", "gather_pool_bwd"), ("gather_pool_fwd", "gather_pool_fwd"),
        ("tok_scatter_add", "tok_scatter_add"), ("adam_flat", "adam_flat")]
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
out = {}
for arg in sys.argv[1:]:
    workload, path = arg.split("=", 1)
    rd = csv.reader(open(path))
    hdr, units = next(rd), next(rd)
    col = {h: i for i, h in enumerate(hdr)}
    ir, iw, it = col["dram__bytes_read.sum"], col["dram__bytes_write.sum"], col["gpu__time_duration.sum"]
    acc = collections.defaultdict(list)
    for r in rd:
        name = r[col["Kernel Name"]]
        for pat, tag in TAGS:
            if re.search(pat, name):
                b = float(r[ir].replace(",", "")) * UNIT[units[ir]] + float(r[iw].replace(",", "")) * UNIT[units[iw]]
                acc[tag].append((b, float(r[it].replace(",", ""))))
    for tag, v in acc.items():
        if tag == "adam_flat":          # the per-bucket launches differ in size: keep the largest
            v = [max(v)]
        out[f"{workload}|{tag}"] = {"bytes": sum(b for b, _ in v) / len(v), "launches": len(v),
                                    "src": "profiles/r02_ncu_summary.md (" + path.split("/")[-1] + ")"}
json.dump(out, open("profiles/ncu_traffic.json", "w"), indent=1)
print(json.dumps(out, indent=1))
