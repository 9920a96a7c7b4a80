"""CPU restatement of the Transformer KG-VAE ('t-SAIL') ELBO step — TEST INFRASTRUCTURE ONLY.

The reference builds this model from torch's stock modules (nn.TransformerEncoder / nn.TransformerDecoder,
/root/reference/kgvae/model/models.py:66-114) and differentiates it with autograd.  This file restates the same
arithmetic as explicit tensor formulas (matmul, softmax, mean/variance) over a plain dict of the reference's
``state_dict`` tensors — no nn.Transformer*, no nn.MultiheadAttention, no F.layer_norm — so that it checks the
algorithm rather than re-running the library module; gradients come from autograd over these formulas in
float64.  It is pinned to outputs of the unmodified reference (tests/golden/tsail_*.npz, written by
oracle/make_golden.py::tsail_case) by tests/test_oracle_golden.py.  Only tests/ may import it.

Dropout is not restated (the parity fixtures run the reference with every dropout probability set to 0).
"""
from __future__ import annotations

import math

import torch


def _ln(x, w, b, eps=1e-5):
    """nn.LayerNorm over the last dim (biased variance), models.py:73,104 via TransformerEncoder/DecoderLayer."""
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * w + b


def _mha(p, pre, q_in, kv_in, n_heads, key_mask=None, causal=False):
    """nn.MultiheadAttention (batch_first): packed in_proj, scaled dot product, softmax, out_proj.
    key_mask [B, S] True = attend; causal adds the triu(-inf) mask of models.py:112."""
    W, b = p[pre + "in_proj_weight"], p[pre + "in_proj_bias"]
    E = W.shape[1]
    q = q_in @ W[:E].T + b[:E]
    k = kv_in @ W[E:2 * E].T + b[E:2 * E]
    v = kv_in @ W[2 * E:].T + b[2 * E:]
    B, T, _ = q.shape
    S = k.shape[1]
    hd = E // n_heads
    q = q.view(B, T, n_heads, hd).transpose(1, 2)
    k = k.view(B, S, n_heads, hd).transpose(1, 2)
    v = v.view(B, S, n_heads, hd).transpose(1, 2)
    s = q @ k.transpose(-1, -2) / math.sqrt(hd)
    if key_mask is not None:
        s = s.masked_fill(~key_mask[:, None, None, :], float("-inf"))
    if causal:
        s = s.masked_fill(torch.triu(torch.ones(T, S, dtype=torch.bool), 1), float("-inf"))
    o = (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(B, T, E)
    return o @ p[pre + "out_proj.weight"].T + p[pre + "out_proj.bias"]


def _ffn(p, pre, x):
    h = torch.relu(x @ p[pre + "linear1.weight"].T + p[pre + "linear1.bias"])
    return h @ p[pre + "linear2.weight"].T + p[pre + "linear2.bias"]


def encoder(p, cfg, triples, eps):
    """AutoRegEncoder.forward — models.py:78-95 (post-LN layers; masked mean-pool AFTER the stack; no clamp)."""
    x = torch.cat([p["enc.e_emb.weight"][triples[:, :, 0]], p["enc.r_emb.weight"][triples[:, :, 1]],
                   p["enc.e_emb.weight"][triples[:, :, 2]]], -1)
    pad_rid = cfg.get("pad_rid")
    mask = (triples[:, :, 1] != pad_rid) if pad_rid is not None else torch.ones(triples.shape[:2], dtype=torch.bool)
    for l in range(cfg["n_layers"]):
        pre = f"enc.txf.layers.{l}."
        x = _ln(x + _mha(p, pre + "self_attn.", x, x, cfg["n_heads"], key_mask=mask), p[pre + "norm1.weight"], p[pre + "norm1.bias"])
        x = _ln(x + _ffn(p, pre, x), p[pre + "norm2.weight"], p[pre + "norm2.bias"])
    m = mask.unsqueeze(-1).to(x.dtype)
    pooled = (x * m).sum(1) / m.sum(1).clamp(min=1)
    mu = pooled @ p["enc.mu.weight"].T + p["enc.mu.bias"]
    logv = pooled @ p["enc.logv.weight"].T + p["enc.logv.bias"]
    return mu + eps * torch.exp(0.5 * logv), mu, logv


def decoder(p, cfg, z, tgt):
    """AutoRegDecoder.forward — models.py:108-114 (memory = z_proj(z) repeated L times, causal self-attention)."""
    B, L = tgt.shape
    x = p["dec.tok_emb.weight"][tgt] + p["dec.pos_emb.weight"][torch.arange(L)][None]
    mem = (z @ p["dec.z_proj.weight"].T + p["dec.z_proj.bias"]).unsqueeze(1).repeat(1, L, 1)
    for l in range(cfg["n_layers"]):
        pre = f"dec.txf.layers.{l}."
        x = _ln(x + _mha(p, pre + "self_attn.", x, x, cfg["n_heads"], causal=True), p[pre + "norm1.weight"], p[pre + "norm1.bias"])
        x = _ln(x + _mha(p, pre + "multihead_attn.", x, mem, cfg["n_heads"]), p[pre + "norm2.weight"], p[pre + "norm2.bias"])
        x = _ln(x + _ffn(p, pre, x), p[pre + "norm3.weight"], p[pre + "norm3.bias"])
    return x @ p["dec.out.weight"].T + p["dec.out.bias"]


def elbo_step(params, cfg, triples, seq, eps, beta, n_tok_global=None, batch_global=None, dtype=torch.float64):
    """loss = CE(ignore PAD) + beta * KL_mean (ablation_study.py:59-71, models.py:199-200) and its gradients.
    params: dict name -> numpy/torch array (reference state_dict names).  Returns (losses, grads, extras)."""
    p = {k: torch.as_tensor(v).to(dtype).clone().requires_grad_(True) for k, v in params.items()}
    triples, seq = torch.as_tensor(triples), torch.as_tensor(seq)
    eps = torch.as_tensor(eps).to(dtype)
    z, mu, logv = encoder(p, cfg, triples, eps)
    logits = decoder(p, cfg, z, seq[:, :-1])
    tgt = seq[:, 1:]
    valid = tgt != 0
    lse = torch.logsumexp(logits, -1)
    nll = lse - logits.gather(-1, tgt.unsqueeze(-1)).squeeze(-1)
    n_tok = float(valid.sum()) if n_tok_global is None else float(n_tok_global)
    B = triples.shape[0]
    b_g = B if batch_global is None else int(batch_global)
    ce = (nll * valid).sum() / n_tok
    kl = -0.5 * (1 + logv - mu ** 2 - logv.exp()).sum() / (b_g * mu.shape[1])
    loss = ce + beta * kl
    loss.backward()
    grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)).detach().numpy() for k, v in p.items()}
    return ({"loss": loss.item(), "ce": ce.item(), "kl": kl.item(), "n_tok": n_tok}, grads,
            {"mu": mu.detach().numpy(), "logv": logv.detach().numpy(), "logits": logits.detach().numpy()})


def tark_step(params, cfg, seq, n_tok_global=None, dtype=torch.float64):
    """Decoder-only Transformer (DecoderOnlyTransformer.forward, models.py:360-365) with the CE-only step of
    train.py:42-58: tok_emb + pos_emb, post-LN encoder layers under a causal mask, (tied) vocabulary projection."""
    p = {k: torch.as_tensor(v).to(dtype).clone().requires_grad_(True) for k, v in params.items() if k != "dec.out.weight"}
    tied = cfg.get("tie_weights", True)
    if not tied:
        p["dec.out.weight"] = torch.as_tensor(params["dec.out.weight"]).to(dtype).clone().requires_grad_(True)
    seq = torch.as_tensor(seq)
    tgt_in, tgt = seq[:, :-1], seq[:, 1:]
    B, L = tgt_in.shape
    x = p["dec.tok_emb.weight"][tgt_in] + p["dec.pos_emb.weight"][torch.arange(L)][None]
    for l in range(cfg["n_layers"]):
        pre = f"dec.txf.layers.{l}."
        x = _ln(x + _mha(p, pre + "self_attn.", x, x, cfg["n_heads"], causal=True), p[pre + "norm1.weight"], p[pre + "norm1.bias"])
        x = _ln(x + _ffn(p, pre, x), p[pre + "norm2.weight"], p[pre + "norm2.bias"])
    w_out = p["dec.tok_emb.weight"] if tied else p["dec.out.weight"]
    logits = x @ w_out.T + p["dec.out.bias"]
    valid = tgt != 0
    nll = torch.logsumexp(logits, -1) - logits.gather(-1, tgt.unsqueeze(-1)).squeeze(-1)
    n_tok = float(valid.sum()) if n_tok_global is None else float(n_tok_global)
    ce = (nll * valid).sum() / n_tok
    ce.backward()
    grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)).detach().numpy() for k, v in p.items()}
    if tied:
        grads["dec.out.weight"] = grads["dec.tok_emb.weight"]
    return {"loss": ce.item(), "ce": ce.item(), "kl": 0.0, "n_tok": n_tok}, grads, {"logits": logits.detach().numpy()}
