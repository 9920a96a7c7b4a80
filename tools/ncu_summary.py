#!/usr/bin/env python
"""Turn the ncu artefacts a gpurun call brought back into the tracked evidence under profiles/.

    python tools/ncu_summary.py launches gpurun_out/launches.csv [--steps-marker pack_tokens --first 7 --last 9]
    python tools/ncu_summary.py full gpurun_out/prof.ncu-rep [more.ncu-rep ...]

`launches`: per-kernel launch count / total device time / share over the TIMED steps of a
`ncu --metrics gpu__time_duration.sum` pass of bench.py (steps are delimited by the one pack_tokens launch
each step makes).  `full`: one markdown row per profiled launch from `ncu -i <rep> --page raw --csv`:
duration, DRAM bytes read/written, DRAM throughput %, tensor-pipe %, active warps %, registers.
Runs in the build container (ncu reads reports without a GPU).
"""
import collections
import csv
import io
import re
import subprocess
import sys


def short(name):
    name = re.sub(r"^void ", "", name).replace("ark::", "")
    m = re.match(r"([\w:]+)(<[^(]*>)?", name)
    base, targs = m.group(1), (m.group(2) or "")
    return base, (base + targs)


def launches(path, marker="pack_tokens", first=None, last=None):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    names = [short(r["Kernel Name"])[0] for r in rows]
    marks = [i for i, n in enumerate(names) if n.startswith(marker)]
    # a step starts a fixed number of launches before its marker: use marker-to-marker windows
    first = len(marks) - 4 if first is None else first          # default: the two timed steps (before the 2 profiled ones)
    last = len(marks) - 2 if last is None else last
    lead = marks[0] if marks[0] < marks[1] - marks[0] else 0
    lo, hi = marks[first] - lead, marks[last] - lead
    agg = collections.OrderedDict()
    for r, n in zip(rows[lo:hi], names[lo:hi]):
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        a[1] += v / 1e3 if unit in ("ns", "nsecond") else (v * 1e3 if unit in ("ms", "msecond") else v)
    tot = sum(a[1] for a in agg.values())
    n_steps = last - first
    out = [f"{hi - lo} launches over {n_steps} timed step(s), {tot / n_steps:.1f} us of kernel time per step "
           f"(cold-cache, serialised by ncu: compare SHARES with bench.py's live breakdown, not absolutes)", "",
           "| kernel | launches/step | us/step | share |", "|---|---|---|---|"]
    for n, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| `{n}` | {c / n_steps:g} | {us / n_steps:.1f} | {100 * us / tot:.1f}% |")
    return "\n".join(out)


COLS = collections.OrderedDict([
    ("time", "gpu__time_duration.sum"), ("rd", "dram__bytes_read.sum"), ("wr", "dram__bytes_write.sum"),
    ("dram%", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    ("tensor%", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
    ("warps%", "sm__warps_active.avg.pct_of_peak_sustained_active"), ("regs", "launch__registers_per_thread")])


def full(paths, only=None, limit_per_kernel=3):
    out = ["| kernel | grid | time | dram rd | dram wr | dram % | tensor pipe % | warps active % | regs |",
           "|---|---|---|---|---|---|---|---|---|"]
    for path in paths:
        if path.endswith(".csv"):      # already exported on the GPU box (`ncu -i rep --page raw --csv`): reports > 64 MiB cannot travel
            txt = open(path).read()
        else:
            txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rd = csv.reader(io.StringIO(txt))
        hdr, units = next(rd), next(rd)
        col = {h: i for i, h in enumerate(hdr)}
        seen = collections.Counter()
        for r in rd:
            base, full_name = short(r[col["Kernel Name"]])
            if only and not re.search(only, base):
                continue
            key = (full_name, r[col["Grid Size"]])
            seen[key] += 1
            if seen[key] > limit_per_kernel:
                continue
            cells = []
            for k, metric in COLS.items():
                i = col.get(metric)
                cells.append("n/a" if i is None or r[i] == "" else f"{float(r[i].replace(',', '')):.3f} {units[i]}".strip())
            out.append(f"| `{full_name}` | {r[col['Grid Size']]} | " + " | ".join(cells) + " |")
    return "\n".join(out)


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        print(launches(sys.argv[2]))
    else:
        print(full(sys.argv[2:]))
