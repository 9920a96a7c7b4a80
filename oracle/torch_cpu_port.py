"""CPU baseline port of the reference's SAIL train step in plain PyTorch — TEST / BASELINE INFRASTRUCTURE ONLY.

The reference's arithmetic for this path IS a sequence of PyTorch CPU ops (nn.Embedding, nn.Linear, nn.GELU,
nn.GRU, F.cross_entropy, torch.optim.Adam).  /root/reference does not exist on the GPU box,
so ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs time THIS restatement (kind = "port") with all
host threads: it issues the same library calls on the same shapes, hence the same MKL/oneDNN kernels.
It is pinned to the reference by tests/test_cpu_port.py (golden outputs of the unmodified reference).
Never imported by the product path.

Citations: encoder kgvae/model/models.py:46-64, decoder :136-142, loss kgvae/experiments/ablation_study.py:59-71,
optimiser :571, step body :43,75-80.
"""
from __future__ import annotations

import time

import torch
import torch.nn.functional as F
from torch import nn


class CpuSail(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        d, dz, nl = cfg["d_model"], cfg["d_latent"], cfg["n_layers"]
        self.pad_rid = cfg.get("pad_rid")
        self.enc = nn.Module()
        self.enc.e_emb = nn.Embedding(cfg["n_entities"], d, padding_idx=cfg.get("pad_eid"))
        self.enc.r_emb = nn.Embedding(cfg["n_relations"], d, padding_idx=self.pad_rid)
        stack = []
        for _ in range(nl):
            stack += [nn.Linear(3 * d, 3 * d), nn.GELU()]
        self.enc.mlp = nn.Sequential(*stack)
        self.enc.mu, self.enc.logv = nn.Linear(3 * d, dz), nn.Linear(3 * d, dz)
        self.dec = nn.Module()
        self.dec.tok_emb = nn.Embedding(cfg["vocab_size"], d)
        self.dec.z_proj = nn.Linear(dz, d)
        p = cfg.get("dec_dropout", 0.1)
        self.dec.gru = nn.GRU(d, d, nl, batch_first=True, dropout=p if nl > 1 else 0.0)
        self.dec.out = nn.Linear(d, cfg["vocab_size"])
        if cfg.get("tie_weights", True):
            self.dec.out.weight = self.dec.tok_emb.weight

    def encode(self, triples, eps=None):
        e, r = self.enc.e_emb, self.enc.r_emb
        x = torch.cat([e(triples[..., 0]), r(triples[..., 1]), e(triples[..., 2])], dim=-1)
        if self.pad_rid is None:
            pooled = x.mean(1)
        else:
            keep = triples[..., 1].ne(self.pad_rid)
            pooled = (x * keep[..., None]).sum(1) / keep.sum(1, keepdim=True).clamp(min=1)
        hid = self.enc.mlp(pooled)
        mu, logv = self.enc.mu(hid), self.enc.logv(hid).clamp(-10, 10)
        noise = torch.randn_like(mu) if eps is None else eps
        return mu + noise * (0.5 * logv).exp(), mu, logv

    def decode(self, z, tokens):
        h0 = torch.tanh(self.dec.z_proj(z))[None].repeat(self.dec.gru.num_layers, 1, 1)
        states, _ = self.dec.gru(self.dec.tok_emb(tokens), h0)
        return self.dec.out(states)

    def elbo(self, triples, seq, beta, eps=None):
        z, mu, logv = self.encode(triples, eps)
        logits = self.decode(z, seq[:, :-1])
        ce = F.cross_entropy(logits.flatten(0, 1), seq[:, 1:].flatten(), ignore_index=0)
        kl = -0.5 * (1 + logv - mu.square() - logv.exp()).mean()
        return ce + beta * kl, ce, kl


def train_steps(model, opt, batches, beta, eps_list=None):
    """The reference's loop body, once per batch; returns per-step (loss, ce, kl) read back like its .item() calls."""
    model.train()
    rec = []
    for i, (triples, seq) in enumerate(batches):
        opt.zero_grad()
        loss, ce, kl = model.elbo(triples, seq, beta, None if eps_list is None else eps_list[i])
        loss.backward()
        opt.step()
        rec.append((loss.item(), ce.item(), kl.item()))
    return rec


def time_cpu_baseline(cfg, batches, n_triples, beta=0.5, lr=1e-3, budget_s=20.0, max_steps=10, threads=None):
    """Wall-clock triples/s of the port on the host cores over a bounded sample of `batches`."""
    import os
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    torch.manual_seed(0)
    model = CpuSail(cfg)
    opt = torch.optim.Adam(model.parameters(), lr=lr)
    t0 = time.perf_counter()
    train_steps(model, opt, batches[:1], beta)                       # warm-up (allocations, MKL init)
    warm = time.perf_counter() - t0
    steps = int(max(2, min(max_steps, budget_s / max(warm, 1e-3))))
    t0 = time.perf_counter()
    done_triples = 0
    for i in range(steps):
        j = i % len(batches)
        train_steps(model, opt, batches[j:j + 1], beta)
        done_triples += n_triples[j]
    dt = time.perf_counter() - t0
    return {"value": done_triples / dt, "unit": "triples/s", "cores": threads, "kind": "port",
            "sample": f"{steps} full train steps (fwd+bwd+Adam, fp32, train mode) of the same workload batches after 1 "
                      f"warm-up, {dt:.1f} s wall", "s_per_step": dt / steps, "steps": steps}


def time_cuda_library_baseline(cfg, batches, n_triples, beta=0.5, lr=1e-3, steps=20, warmup=5, tf32=False, device="cuda"):
    """The SAME restated reference modules on device='cuda' (what the reference does when a GPU is present,
    kgvae/experiments/ablation_study.py:392): torch eager -> cuDNN RNN, cuBLASLt, ATen embedding / softmax kernels,
    torch.optim.Adam.  This is the library-path bar of SURVEY.md §2.3 / §8(d), CUDA-event timed over whole train steps
    (zero_grad, forward, CE + beta*KL, backward, Adam, three .item() read-backs per step exactly like
    ablation_study.py:59-80), inputs already resident on the device."""
    old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = bool(tf32)
    torch.backends.cudnn.allow_tf32 = bool(tf32)
    try:
        torch.manual_seed(0)
        model = CpuSail(cfg).to(device)
        opt = torch.optim.Adam(model.parameters(), lr=lr)
        dev_batches = [(t.to(device), s.to(device)) for t, s in batches]
        for i in range(warmup):
            train_steps(model, opt, dev_batches[i % len(dev_batches):i % len(dev_batches) + 1], beta)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        done = 0
        e0.record()
        for i in range(steps):
            j = i % len(dev_batches)
            train_steps(model, opt, dev_batches[j:j + 1], beta)
            done += n_triples[j]
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old
    return {"value": done / (ms * 1e-3), "unit": "triples/s", "ms_per_step": ms / steps, "steps": steps,
            "precision": "tf32" if tf32 else "fp32", "kind": "port on cuda (torch eager: cuDNN GRU, cuBLASLt, ATen)"}


@torch.no_grad()
def posterior_bits_port(model: CpuSail, triples, seq, eps):
    """(ar_bits, kl_bits) of ONE graph: the reference's posterior_bits record (kgvae/model/models.py:218-260) with
    bits_per_sequence (:202-213) restated as a single teacher-forced pass — the GRU decoder is causal, so the logits
    of position t under the full prefix equal the last-position logits of the reference's prefix [:t] loop."""
    import math
    model.eval()
    z, mu, logv = model.encode(triples[None], eps[None])
    n = int((seq[1:] != 0).long().cumprod(0).sum())
    ar = 0.0
    if n:
        logp = F.log_softmax(model.decode(z, seq[None, :n]), dim=-1)[0]
        ar = float(-logp[torch.arange(n), seq[1:n + 1]].sum() / math.log(2))
    kl = float(-0.5 * (1 + logv - mu.square() - logv.exp()).sum() / math.log(2))
    return ar, kl
