"""Flat fp32 parameter / gradient / Adam-state storage with a bf16 shadow, shared with an nn.Module.

The module keeps the reference's parameter names and shapes — every ``nn.Parameter.data``
becomes a VIEW into one flat fp32 buffer, ordered by the time its gradient becomes final in the
hand-written backward pass, so that (a) Adam is one launch over one buffer, (b) data-parallel gradient
buckets are contiguous slices that can be all-reduced while the rest of backward is still running, and
(c) the bf16 operand copies of the weights are one flat shadow refreshed by the Adam kernel.
"""
from __future__ import annotations

import torch

ALIGN = 64  # elements: 256 B for fp32, 128 B for the bf16 shadow (TMA needs 16 B)


def sail_param_order(model):
    """Gradient-readiness order of SAIL's parameters in ark_b200.elbo's backward pass.

    Returns a list of groups; each group is a list of (name, parameter) stored back to back WITHOUT padding
    (enc.mu / enc.logv are fused into one [2*dz, 3d] GEMM operand, models.py:43-44,61-62).
    """
    named = dict(model.named_parameters())  # tied dec.out.weight is deduplicated by torch
    nl = model.dec.gru.num_layers
    groups = [[("dec.out.bias", named["dec.out.bias"])]]
    if "dec.out.weight" in named:  # tie_weights: false
        groups.append([("dec.out.weight", named["dec.out.weight"])])
    for k in range(nl - 1, -1, -1):
        for nm in (f"weight_ih_l{k}", f"weight_hh_l{k}", f"bias_ih_l{k}", f"bias_hh_l{k}"):
            groups.append([(f"dec.gru.{nm}", named[f"dec.gru.{nm}"])])
    groups.append([("dec.tok_emb.weight", named["dec.tok_emb.weight"])])
    groups.append([("dec.z_proj.weight", named["dec.z_proj.weight"])])
    groups.append([("dec.z_proj.bias", named["dec.z_proj.bias"])])
    groups.append([("enc.mu.weight", named["enc.mu.weight"]), ("enc.logv.weight", named["enc.logv.weight"])])
    groups.append([("enc.mu.bias", named["enc.mu.bias"]), ("enc.logv.bias", named["enc.logv.bias"])])
    n_mlp = sum(1 for n in named if n.startswith("enc.mlp.") and n.endswith(".weight"))
    for k in range(n_mlp - 1, -1, -1):
        groups.append([(f"enc.mlp.{2 * k}.weight", named[f"enc.mlp.{2 * k}.weight"])])
        groups.append([(f"enc.mlp.{2 * k}.bias", named[f"enc.mlp.{2 * k}.bias"])])
    groups.append([("enc.r_emb.weight", named["enc.r_emb.weight"])])
    groups.append([("enc.e_emb.weight", named["enc.e_emb.weight"])])
    seen = {n for g in groups for n, _ in g}
    missing = set(named) - seen
    if missing:
        raise RuntimeError(f"parameters without a slot in the flat layout: {sorted(missing)}")
    return groups


def ark_param_order(model):
    """Gradient-readiness order for the decoder-only ARK model (reference models.py:323-405)."""
    named = dict(model.named_parameters())
    nl = model.dec.gru.num_layers
    groups = [[("dec.out.bias", named["dec.out.bias"])]]
    if "dec.out.weight" in named:
        groups.append([("dec.out.weight", named["dec.out.weight"])])
    for k in range(nl - 1, -1, -1):
        for nm in (f"weight_ih_l{k}", f"weight_hh_l{k}", f"bias_ih_l{k}", f"bias_hh_l{k}"):
            groups.append([(f"dec.gru.{nm}", named[f"dec.gru.{nm}"])])
    groups.append([("dec.tok_emb.weight", named["dec.tok_emb.weight"])])
    groups.append([("dec.pos_emb.weight", named["dec.pos_emb.weight"])])
    missing = set(named) - {n for g in groups for n, _ in g}
    if missing:
        raise RuntimeError(f"parameters without a slot in the flat layout: {sorted(missing)}")
    return groups


def merge_span(pending, s, e, gap=64):
    """Append the slot range [s, e) of the flat buffer to `pending` (a list of (start, end)); a range that begins
    directly behind the last one (within the alignment gap) extends it, anything else — including ranges that arrive
    out of layout order — starts a new entry."""
    if pending and s - gap <= pending[-1][1] <= s:
        pending[-1] = (pending[-1][0], e)
    else:
        pending.append((s, e))
    return pending


class FlatParams:
    def __init__(self, groups, device, alloc=None):
        """alloc(numel) -> (param f32, grad f32, shadow bf16) zero-filled flat buffers: data parallelism places them in
        symmetric multicast memory (ark_b200.symm.SymmFlat)."""
        self.slots = {}  # name -> (offset, numel, shape)
        off = 0
        for g in groups:
            off = (off + ALIGN - 1) // ALIGN * ALIGN
            for name, p in g:
                self.slots[name] = (off, p.numel(), tuple(p.shape))
                off += p.numel()
        self.numel = (off + ALIGN - 1) // ALIGN * ALIGN
        self.device = device
        if alloc is not None:
            self.param, self.grad, shadow = alloc(self.numel)
        else:
            self.param = torch.zeros(self.numel, device=device, dtype=torch.float32)
            self.grad = torch.zeros(self.numel, device=device, dtype=torch.float32)
        self.exp_avg = torch.zeros(self.numel, device=device, dtype=torch.float32)
        self.exp_avg_sq = torch.zeros(self.numel, device=device, dtype=torch.float32)
        self.shadow = shadow if alloc is not None else torch.zeros(self.numel, device=device, dtype=torch.bfloat16)
        self.order = [n for g in groups for n, _ in g]
        with torch.no_grad():
            for g in groups:
                for name, p in g:
                    view = self.view(self.param, name)
                    view.copy_(p.data.to(device))
                    p.data = view                     # the module now aliases the flat buffer
                    p.grad = self.view(self.grad, name)

    def view(self, buf, name):
        off, n, shape = self.slots[name]
        return buf[off:off + n].view(shape)

    def fused(self, buf, first, last, shape):
        """View spanning consecutive slots first..last (stored back to back) as one matrix."""
        o0 = self.slots[first][0]
        o1 = self.slots[last][0] + self.slots[last][1]
        return buf[o0:o1].view(shape)

    def p(self, name):
        return self.view(self.param, name)

    def g(self, name):
        return self.view(self.grad, name)

    def s(self, name):
        return self.view(self.shadow, name)

    def span(self, first, last):
        """(start, end) element range covering slots first..last in flat order (for gradient buckets)."""
        return self.slots[first][0], self.slots[last][0] + self.slots[last][1]
