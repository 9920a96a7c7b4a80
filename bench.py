#!/usr/bin/env python
"""bench.py — ELBO train triples/sec of the KG-VAE (SAIL) hot path on N B200s.

    python bench.py --gpus 1 --steps 20 --warmup 5                       # this repository's CUDA path
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...                                 # the reference's CPU path (port), host cores

A step = zero_grad + ELBO forward + backward (+ NCCL gradient all-reduce) + Adam over one synthetic
IntelliGraphs-shaped batch per GPU (weak scaling: the YAML batch_size on every GPU).  `value` = real (non-PAD)
triples per second, whole job, device-timed with CUDA events, max over ranks, inputs resident in HBM.
`e2e` = the same through the public call `SAIL.elbo_step(host tensors)`: host-side packing, pinned H2D copies
and a device->host read of (ce, kl) every step inside the timed region.
One JSON line on stdout (rank 0).  Per-op breakdown goes to stderr / gpurun_out/bench_breakdown.json.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "ELBO train triples/sec (SAIL fwd+bwd+Adam, synthetic IntelliGraphs-shaped batches)"
UNIT = "triples/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ark", choices=["ark", "reference"])
    ap.add_argument("--workload", default="syn-types",
                    choices=["syn-paths", "syn-types", "syn-tipr", "wd-movies", "wd-articles"])
    ap.add_argument("--model", default="SAIL", choices=["SAIL", "t-SAIL", "ARK", "t-ARK"],
                    help="SAIL = the KG-VAE ELBO path (headline); t-SAIL = Transformer KG-VAE; ARK = decoder-only GRU")
    ap.add_argument("--batch", type=int, default=0, help="graphs per GPU (default: the YAML batch_size)")
    ap.add_argument("--dense", action="store_true", help="every graph at max_edges (what the reference pays for)")
    ap.add_argument("--backend", default="tc", choices=["tc", "simt"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel eagerly instead of replaying CUDA graphs")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "src": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "src": "fallback"}


class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.sm, self.reasons, self.max_mhz = index, False, [], set(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}
        while not self.stop_flag:
            try:
                self.sm.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                bits = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for b, n in names.items():
                    if bits & b:
                        self.reasons.add(n)
            except Exception:
                pass
            time.sleep(0.05)

    def result(self):
        self.stop_flag = True
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}
        s = sorted(self.sm)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def make_host_batches(cfg, batch, rank, n_batches, dense):
    from ark_b200.synthetic import synth_batch
    return [synth_batch(cfg, batch, 1234 + 1000 * rank + i, dense=dense) for i in range(n_batches)]


def run_reference(args):
    """The reference's CPU path (torch-CPU port, oracle/torch_cpu_port.py) on the host cores, same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from ark_b200.synthetic import model_config
    from oracle.torch_cpu_port import CpuSail, train_steps
    cfg = model_config(args.workload)
    batch = args.batch or cfg["batch_size"]
    hb = make_host_batches(cfg, batch, 0, 2, args.dense)
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    torch.manual_seed(0)
    model = CpuSail(cfg)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    # bounded: each "step" is one full train step; K and W as asked but capped so the run ends within minutes
    t0 = time.perf_counter()
    train_steps(model, opt, [hb[0][:2]], 0.5)
    first = time.perf_counter() - t0
    W = max(0, min(args.warmup, 1) - 1)      # the probe step above already is the warm-up
    K = int(max(1, min(args.steps, 120.0 / max(first, 1e-3))))
    for i in range(W):
        train_steps(model, opt, [hb[i % 2][:2]], 0.5)
    t0 = time.perf_counter()
    tri = 0
    for i in range(K):
        train_steps(model, opt, [hb[i % 2][:2]], 0.5)
        tri += hb[i % 2][2]
    dt = time.perf_counter() - t0
    v = tri / dt
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": K,
        "warmup": 1 + W, "ms_per_step": 1e3 * dt / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"autoreg_{args.workload} SAIL", "batch_per_step": batch, "dense": args.dense,
                   "note": "reference CPU path = torch-CPU port of the reference step (same ATen/MKL calls); "
                           "steps capped so the run ends within minutes"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{K} full train steps of the workload batch, {dt:.1f} s wall"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch multi-GPU runs with torch.distributed.run (one rank per GPU)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    group = None
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=dev)
        group = torch.distributed.group.WORLD

    from ark_b200 import _C
    from ark_b200.layout import pack_tlayout
    from ark_b200.synthetic import DeviceBatch, model_config
    from kgvae.model.models import ARK, SAIL

    cfg = model_config(args.workload, model_type=args.model)
    mt = args.model
    batch = args.batch or cfg["batch_size"]
    torch.manual_seed(0)                       # identical initial weights on every rank
    dec_only = mt in ("ARK", "t-ARK")
    model = (ARK if dec_only else SAIL)(cfg).to(dev)
    eng = model.engine(lr=1e-3, gemm_backend=args.backend, dist_group=group)
    n_params = sum(p.numel() for p in model.parameters())

    NB = 4
    host = make_host_batches(cfg, batch, rank, NB, args.dense)
    dbs = [DeviceBatch(t, s, n, dev, 1234 + 1000 * rank + i) for i, (t, s, n) in enumerate(host)]
    if mt in ("t-SAIL", "t-ARK"):               # graph-major ragged rows instead of time-major packed rows
        for b_, (t, s, _) in zip(dbs, host):
            b_.layout = pack_tlayout(t, s, cfg.get("pad_rid")).to(dev)
    eps = [b.eps(cfg["d_latent"], dev) if not dec_only else None for b in dbs]
    # global normalisers: the sampler knows every rank's lengths, so no per-step collective
    ntok = torch.tensor([b.layout.n_tok for b in dbs], device=dev, dtype=torch.float64)
    ntri = torch.tensor([b.n_triples for b in dbs], device=dev, dtype=torch.float64)
    if world > 1:
        torch.distributed.all_reduce(ntok)
        torch.distributed.all_reduce(ntri)
    ntok_g, ntri_g = ntok.tolist(), ntri.tolist()
    bg = batch * world
    beta = 0.5

    use_graph = not args.no_graph      # N>1: graph segments cut at the gradient buckets, NCCL eager in between

    def step(i, graph=use_graph):
        j = i % NB
        fn = eng.train_step_graphed if graph else eng.train_step
        return fn(dbs[j].triples if not dec_only else None, dbs[j].seq, dbs[j].layout, eps[j], beta if not dec_only else 0.0,
                  n_tok_global=ntok_g[j], batch_global=bg if not dec_only else None)

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    for i in range(max(args.warmup, 3) + (NB if use_graph else 0)):   # graph mode: first visit of a layout captures
        step(i)
    barrier()
    lib = _C.lib()
    sampler = ClockSampler(local)
    sampler.start()
    lib.reset_launch_count()
    replayed0 = eng.launches_replayed
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(args.steps):
        step(i)
    e1.record()
    barrier()
    launches = lib.launch_count() + (eng.launches_replayed - replayed0)   # eager launches + kernels inside replayed graphs
    clocks = sampler.result()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
    total_ms = ms.item()
    triples_done = sum(ntri_g[i % NB] for i in range(args.steps))
    value = triples_done / (total_ms / 1e3)
    final = eng.read_stats(beta)

    # ---------------- e2e: public API with HOST buffers, H2D + D2H inside the timed region
    e2e = None
    if not args.no_e2e:
        pinned = [b.host for b in dbs]
        lay0 = dbs[0].layout
        meta_bytes = (batch + 2 * lay0.L) * 4 if hasattr(lay0, "L") else \
            (lay0.idx.nbytes + lay0.tok.nbytes * 3 + lay0.enc.cu.nbytes * 2 + lay0.enc.sq_off.nbytes * 2 + lay0.enc.graph.nbytes
             + lay0.dec.graph.nbytes)
        h2d = pinned[0][0].numel() * 8 + pinned[0][1].numel() * 8 + meta_bytes
        def api_step(j):
            if dec_only:
                return model.ce_step(pinned[j][1], n_tok_global=ntok_g[j])
            return model.elbo_step(pinned[j][0], pinned[j][1], beta, n_tok_global=ntok_g[j], batch_global=bg, graph=use_graph)

        for i in range(2):
            api_step(i % NB).tolist()
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        f0.record()
        for i in range(args.steps):
            j = i % NB
            out = api_step(j)
            out.tolist()                                    # device->host read of the step's (ce, kl)
        f1.record()
        barrier()
        wall_ms = (time.perf_counter() - t0) * 1e3
        ems = torch.tensor([max(f0.elapsed_time(f1), wall_ms)], device=dev, dtype=torch.float64)
        if world > 1:
            torch.distributed.all_reduce(ems, op=torch.distributed.ReduceOp.MAX)
        e2e = {"value": triples_done / (ems.item() / 1e3), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": 8, "ms_per_step": ems.item() / args.steps,
               "api": "kgvae.model.models.ARK.ce_step(seq_cpu)" if dec_only else
                      "kgvae.model.models.SAIL.elbo_step(triples_cpu, seq_cpu, beta)"}

    # ---------------- roofline pass: CUDA events around every op of the same steps (rank 0 reports)
    eng.prof = []
    for i in range(min(args.steps, 8)):
        step(i, graph=False)          # events cannot be recorded inside a replayed graph: eager launches
    agg = eng.profile_summary()
    n_prof = min(args.steps, 8)
    eng.prof = None
    pk = peaks()
    tot_ms = sum(a["ms"] for a in agg.values())
    fam = {}
    for tag, a in agg.items():
        k = tag.split(":")[0]
        f_ = fam.setdefault(k, {"ms": 0.0, "calls": 0, "flops": 0.0, "bytes": 0.0})
        for kk in f_:
            f_[kk] += a[kk]
    top = max(((k, v) for k, v in fam.items() if not k.startswith("nccl_")), key=lambda kv: kv[1]["ms"])   # our kernels only
    name, a = top
    if a["flops"] > 0:
        ach = a["flops"] / (a["ms"] * 1e-3) / 1e12
        roof = {"kernel": name, "bound": "tensor", "achieved": ach, "peak": pk["bf16_tflops_sustained"],
                "unit": "TFLOP/s", "frac": ach / pk["bf16_tflops_sustained"], "traffic": None,
                "peak_src": pk["src"] + " (sustained: timed inside a long step)"}
    else:
        ach = a["bytes"] / (a["ms"] * 1e-3) / 1e9
        roof = {"kernel": name, "bound": "hbm", "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s",
                "frac": ach / pk["hbm_gbs"], "traffic": None, "peak_src": pk["src"]}
    # DRAM bytes per launch of that kernel from the committed `ncu --set full` captures (profiles/), where one exists
    # for this workload: dram__bytes_read.sum + dram__bytes_write.sum
    ncu_traffic = {("syn-types", "gru_persist_bwd"): (44.08e6 + 3.47e6, "profiles/r01_ncu_summary.md"),
                   ("syn-types", "gru_persist_fwd"): (39.37e6 + 2.27e6, "profiles/r01_ncu_summary.md"),
                   ("wd-articles", "gru_cluster_bwd"): (99.65e6 + 90.60e6, "profiles/r01c_ncu_summary.md"),
                   ("wd-articles", "gru_cluster_fwd"): (54.30e6 + 225.63e6, "profiles/r01c_ncu_summary.md")}
    if mt == "SAIL" and not args.dense and args.batch == 0 and (args.workload, name) in ncu_traffic:
        roof["traffic"], roof["traffic_src"] = ncu_traffic[(args.workload, name)]
    roof["share_of_step"] = a["ms"] / max(tot_ms, 1e-9)
    roof["avg_launch_ms"] = a["ms"] / max(a["calls"], 1)
    breakdown = {t: {"ms_per_step": v["ms"] / n_prof, "calls_per_step": v["calls"] / n_prof,
                     "tflops": (v["flops"] / (v["ms"] * 1e-3) / 1e12) if v["flops"] and v["ms"] else None,
                     "gbs": (v["bytes"] / (v["ms"] * 1e-3) / 1e9) if v["bytes"] and v["ms"] else None}
                 for t, v in sorted(agg.items(), key=lambda kv: -kv[1]["ms"])}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and mt == "SAIL":
        from oracle.torch_cpu_port import time_cpu_baseline   # the checker/baseline, never the product
        cpu = time_cpu_baseline(cfg, [(t, s) for t, s, _ in host], [n for _, _, n in host], beta=beta,
                                budget_s=20.0, max_steps=8)

    if rank == 0:
        print(json.dumps({"breakdown_ms_per_step": breakdown, "sum_ms": tot_ms / n_prof}, indent=1), file=sys.stderr)
        try:
            os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
            with open(os.path.join(ROOT, "gpurun_out", f"bench_breakdown_{args.workload}_n{world}.json"), "w") as f:
                json.dump({"breakdown": breakdown, "families": fam, "profiled_steps": n_prof}, f, indent=1)
        except OSError:
            pass
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"autoreg_{args.workload} {mt}", "graphs_per_gpu": batch, "global_batch": bg,
                       "d_model": cfg["d_model"], "d_latent": cfg["d_latent"], "n_layers": cfg["n_layers"],
                       "vocab_size": cfg["vocab_size"], "seq_len": cfg["seq_len"], "params": n_params,
                       "triples_per_step": triples_done / args.steps, "tokens_per_step_rank0": dbs[0].layout.n_tok,
                       "dense": bool(args.dense), "parallelism": f"dp{world}",
                       "l2": "per-step working set (params+grads+Adam state+activations) exceeds the 126 MB L2; "
                             "no explicit flush", "gemm_backend": args.backend, "cuda_graph": bool(use_graph),
                       "precision": "bf16 GEMM operands, fp32 accumulate/master/state"},
            "clocks": clocks, "gpu_launches": int(launches),
            "roofline": roof, "final_loss": {"loss": final[0], "ce": final[1], "kl": final[2]},
        }
        if e2e is not None:
            line["e2e"] = e2e
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line))
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
