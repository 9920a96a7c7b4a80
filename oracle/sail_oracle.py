"""CPU oracle for the KG-VAE (SAIL) ELBO training step — TEST INFRASTRUCTURE ONLY.

A numpy restatement (float64 by default) of the algorithm the reference executes through
PyTorch, with the backward pass written out by hand so that it is independent of both
torch.autograd and of the CUDA kernels it checks.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` leg may import this module;
the product path (``ark_b200/``, ``kgvae/``) never does.

Parity status: the reference ships no tests / golden vectors for this path
 so there is nothing of the reference's own to pin against.  This oracle is
instead pinned against OUTPUTS OF THE REFERENCE ITSELF, executed in the build container by
``oracle/make_golden.py`` and committed under ``tests/golden/`` (see tests/test_oracle_golden.py).

Every function cites the reference file:line (relative to /root/reference) it restates.
Parameter names are the reference's ``state_dict`` keys.
"""
from __future__ import annotations

import math

import numpy as np

try:  # scipy is present in the image; keep a pure-numpy fallback for erf anyway
    from scipy.special import erf as _erf
except Exception:  # pragma: no cover
    _erf = np.vectorize(math.erf)

PAD, BOS, EOS = 0, 1, 2


# --------------------------------------------------------------------------------------
# Integer side: vocabulary layout and triple <-> token-sequence indexing (bit-exact)
# --------------------------------------------------------------------------------------

def vocab_layout(n_entities, n_relations, max_edges, use_padding):
    """kgvae/experiments/ablation_study.py:436-454 — vocabulary sizing and token bases."""
    nE, nR = int(n_entities), int(n_relations)
    if use_padding:
        pad_eid, pad_rid = nE, nR
        nE += 1
        nR += 1
    else:
        pad_eid = pad_rid = None
    ent_base = 3
    rel_base = ent_base + nE
    return {
        "n_entities": nE, "n_relations": nR, "pad_eid": pad_eid, "pad_rid": pad_rid,
        "special_tokens": {"PAD": PAD, "BOS": BOS, "EOS": EOS},
        "ENT_BASE": ent_base, "REL_BASE": rel_base, "vocab_size": rel_base + nR,
        "seq_len": 1 + max_edges * 3 + 1, "max_edges": int(max_edges),
    }


def triples_to_seq(triples, layout):
    """kgvae/model/utils.py:102-108 — [BOS, (E+h, R+r, E+t)*n, EOS, PAD...] of length seq_len."""
    seq = [BOS]
    for h, r, t in triples:
        seq += [layout["ENT_BASE"] + h, layout["REL_BASE"] + r, layout["ENT_BASE"] + t]
    seq.append(EOS)
    seq += [PAD] * (layout["seq_len"] - len(seq))
    return np.asarray(seq, dtype=np.int64)


def seq_to_triples(seq, layout):
    """kgvae/model/utils.py:70-78 — inverse map; stops at EOS or when < 3 tokens remain."""
    seq = [int(x) for x in seq]
    out, i = [], 1
    while i + 2 < len(seq) and seq[i] != EOS:
        h, r, t = seq[i:i + 3]
        out.append((h - layout["ENT_BASE"], r - layout["REL_BASE"], t - layout["ENT_BASE"]))
        i += 3
    return out


def build_batch(graphs, layout):
    """kgvae/model/utils.py:131-146 (GraphSeqDataset.__getitem__, permute off) + default collate.

    Returns (triples[B,T,3] int64, seq[B,seq_len] int64).  With padding every graph is padded
    to ``max_edges`` with (pad_eid, pad_rid, pad_eid); without padding all graphs must have
    the same number of triples (that is what torch's default collate requires too).
    """
    tri, seqs = [], []
    for g in graphs:
        g = [tuple(int(v) for v in t) for t in g]
        if layout["pad_rid"] is not None:
            pad = (layout["pad_eid"], layout["pad_rid"], layout["pad_eid"])
            tri.append(g + [pad] * (layout["max_edges"] - len(g)))
        else:
            tri.append(g)
        seqs.append(triples_to_seq(g, layout))
    return np.asarray(tri, dtype=np.int64), np.stack(seqs)


# --------------------------------------------------------------------------------------
# Floating-point side
# --------------------------------------------------------------------------------------

def gelu_erf(x):
    """nn.GELU() default = exact erf form (kgvae/model/models.py:37)."""
    return 0.5 * x * (1.0 + _erf(x / math.sqrt(2.0)))


def gelu_erf_grad(x):
    return 0.5 * (1.0 + _erf(x / math.sqrt(2.0))) + x * np.exp(-0.5 * x * x) / math.sqrt(2.0 * math.pi)


def _sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def n_mlp_layers(params):
    k = 0
    while f"enc.mlp.{2 * k}.weight" in params:
        k += 1
    return k


def n_gru_layers(params):
    k = 0
    while f"dec.gru.weight_ih_l{k}" in params:
        k += 1
    return k


def encoder_forward(params, triples, eps, pad_rid):
    """AutoRegEncoderMLP.forward — kgvae/model/models.py:46-64 (ε passed in, not drawn)."""
    E, R = params["enc.e_emb.weight"], params["enc.r_emb.weight"]
    h, r, t = triples[:, :, 0], triples[:, :, 1], triples[:, :, 2]
    x = np.concatenate([E[h], R[r], E[t]], axis=-1)                      # :47-50
    if pad_rid is not None:
        mask = (r != pad_rid)                                             # :52
        cnt = np.maximum(mask.sum(1, keepdims=True), 1).astype(x.dtype)   # :54
        g = (x * mask[..., None]).sum(1) / cnt                            # :53,55
    else:
        mask = np.ones(r.shape, dtype=bool)
        cnt = np.full((x.shape[0], 1), x.shape[1], dtype=x.dtype)
        g = x.mean(1)                                                     # :58
    acts, pre = [g], []
    a = g
    for k in range(n_mlp_layers(params)):                                 # :32-41,60
        p = a @ params[f"enc.mlp.{2 * k}.weight"].T + params[f"enc.mlp.{2 * k}.bias"]
        a = gelu_erf(p)
        pre.append(p)
        acts.append(a)
    mu = a @ params["enc.mu.weight"].T + params["enc.mu.bias"]            # :61
    logv_raw = a @ params["enc.logv.weight"].T + params["enc.logv.bias"]
    logv = np.clip(logv_raw, -10.0, 10.0)                                 # :62
    sigma = np.exp(0.5 * logv)
    z = mu + eps * sigma                                                  # :63
    return {"mask": mask, "cnt": cnt, "g": g, "acts": acts, "pre": pre, "mu": mu,
            "logv_raw": logv_raw, "logv": logv, "sigma": sigma, "z": z}


def gru_decoder_forward(params, z, seq_in, drop_masks=None, tied=True, decoder_only=False):
    """AutoRegDecoderGRU.forward — kgvae/model/models.py:136-142.

    GRU equations are torch.nn.GRU's (gate order r,z,n).
    ``drop_masks[k]`` (optional, shape [B,L,d], already scaled by 1/(1-p)) is the inter-layer
    dropout multiplier applied to the OUTPUT of layer k for k < n_layers-1 (train mode only).
    """
    Wt = params["dec.tok_emb.weight"]
    B, L = seq_in.shape
    d = Wt.shape[1]
    nl = n_gru_layers(params)
    if decoder_only:   # DecoderOnlyGRU.forward — models.py:340-343: tok_emb + pos_emb, nn.GRU default h0 = 0
        x = Wt[seq_in] + params["dec.pos_emb.weight"][np.arange(L)][None]
        h0 = np.zeros((B, d), dtype=Wt.dtype)
    else:
        x = Wt[seq_in]                                                    # :138
        h0 = np.tanh(z @ params["dec.z_proj.weight"].T + params["dec.z_proj.bias"])  # :139
    layers = []
    u = x
    for k in range(nl):                                                   # :141
        Wih, Whh = params[f"dec.gru.weight_ih_l{k}"], params[f"dec.gru.weight_hh_l{k}"]
        bih, bhh = params[f"dec.gru.bias_ih_l{k}"], params[f"dec.gru.bias_hh_l{k}"]
        gi = u @ Wih.T + bih                                              # [B,L,3d]
        y = np.empty((B, L, d), dtype=x.dtype)
        rr = np.empty_like(y); zz = np.empty_like(y); nn_ = np.empty_like(y); ghn = np.empty_like(y)
        h = h0
        for i in range(L):
            gh = h @ Whh.T + bhh
            r = _sigmoid(gi[:, i, :d] + gh[:, :d])
            zg = _sigmoid(gi[:, i, d:2 * d] + gh[:, d:2 * d])
            n = np.tanh(gi[:, i, 2 * d:] + r * gh[:, 2 * d:])
            h = (1.0 - zg) * n + zg * h
            y[:, i], rr[:, i], zz[:, i], nn_[:, i], ghn[:, i] = h, r, zg, n, gh[:, 2 * d:]
        layers.append({"u": u, "y": y, "r": rr, "z": zz, "n": nn_, "ghn": ghn})
        u = y
        if drop_masks is not None and k < nl - 1:
            u = y * drop_masks[k]
    Wout = Wt if tied else params["dec.out.weight"]
    logits = u @ Wout.T + params["dec.out.bias"]                          # :142
    return {"x": x, "h0": h0, "layers": layers, "y_top": u, "logits": logits}


def cross_entropy_ignore_pad(logits, tgt, n_tok=None):
    """F.cross_entropy(..., ignore_index=PAD) — kgvae/experiments/ablation_study.py:65-69."""
    V = logits.shape[-1]
    lg = logits.reshape(-1, V)
    tg = tgt.reshape(-1)
    m = lg.max(1, keepdims=True)
    lse = m[:, 0] + np.log(np.exp(lg - m).sum(1))
    valid = tg != PAD
    n = float(valid.sum()) if n_tok is None else float(n_tok)
    nll = lse - lg[np.arange(lg.shape[0]), tg]
    ce = (nll * valid).sum() / n
    return ce, lse, valid, n


def kl_mean(mu, logv):
    """SAIL.kl_mean — kgvae/model/models.py:199-200 (mean over B*dz, not sum over dz)."""
    return -0.5 * np.mean(1.0 + logv - mu ** 2 - np.exp(logv))


def elbo_forward(params, cfg, triples, seq, eps, beta, drop_masks=None):
    """SAIL.forward + loss — models.py:317-320, ablation_study.py:59-71."""
    tied = cfg.get("tie_weights", True)
    enc = encoder_forward(params, triples, eps, cfg.get("pad_rid"))
    dec = gru_decoder_forward(params, enc["z"], seq[:, :-1], drop_masks, tied)
    ce, lse, valid, n_tok = cross_entropy_ignore_pad(dec["logits"], seq[:, 1:])
    kl = kl_mean(enc["mu"], enc["logv"])
    return {"enc": enc, "dec": dec, "ce": ce, "kl": kl, "loss": ce + beta * kl,
            "lse": lse, "valid": valid, "n_tok": n_tok}


def _decoder_backward(params, fw, seq_in, tgt, n_tok, tied, drop_masks, grads):
    """CE + vocabulary projection + GRU stack + token-embedding backward, shared by the SAIL and ARK steps.
    Fills ``grads`` in place; returns (gradient w.r.t. the layer-0 input rows [B,L,d], summed dh0 [B,d])."""
    dec = fw["dec"]
    B, L = seq_in.shape
    Wt = params["dec.tok_emb.weight"]
    V, d = Wt.shape
    nl = n_gru_layers(params)
    dt = Wt.dtype
    # ---- CE backward: (softmax - onehot)/N_tok on non-PAD rows ----
    lg = dec["logits"].reshape(-1, V)
    p = np.exp(lg - fw["lse"][:, None])
    p[np.arange(B * L), tgt.reshape(-1)] -= 1.0
    dlogits = p * (fw["valid"][:, None] / n_tok)
    ytop = dec["y_top"].reshape(-1, d)
    Wout = Wt if tied else params["dec.out.weight"]
    gW_out = dlogits.T @ ytop
    grads["dec.out.bias"] = dlogits.sum(0)
    dy = (dlogits @ Wout).reshape(B, L, d)
    if tied:
        grads["dec.tok_emb.weight"] += gW_out
    else:
        grads["dec.out.weight"] = gW_out

    # ---- GRU backward, top layer down, reverse time ----
    dh0_total = np.zeros((B, d), dtype=dt)
    for k in range(nl - 1, -1, -1):
        lay = dec["layers"][k]
        if drop_masks is not None and k < nl - 1:
            dy = dy * drop_masks[k]
        Wih, Whh = params[f"dec.gru.weight_ih_l{k}"], params[f"dec.gru.weight_hh_l{k}"]
        dgi = np.empty((B, L, 3 * d), dtype=dt)
        dgh = np.empty((B, L, 3 * d), dtype=dt)
        hprev_all = np.concatenate([dec["h0"][:, None, :], lay["y"][:, :-1]], axis=1)
        dh = np.zeros((B, d), dtype=dt)
        for i in range(L - 1, -1, -1):
            r, zg, n, ghn, hp = lay["r"][:, i], lay["z"][:, i], lay["n"][:, i], lay["ghn"][:, i], hprev_all[:, i]
            dht = dy[:, i] + dh
            dn = dht * (1.0 - zg)
            dzg = dht * (hp - n)
            dan = dn * (1.0 - n * n)
            dar = dan * ghn * r * (1.0 - r)
            daz = dzg * zg * (1.0 - zg)
            dgi[:, i] = np.concatenate([dar, daz, dan], axis=1)
            dgh[:, i] = np.concatenate([dar, daz, dan * r], axis=1)
            dh = dht * zg + dgh[:, i] @ Whh
        dh0_total += dh
        dgi2, dgh2 = dgi.reshape(-1, 3 * d), dgh.reshape(-1, 3 * d)
        grads[f"dec.gru.weight_ih_l{k}"] = dgi2.T @ lay["u"].reshape(-1, d)
        grads[f"dec.gru.weight_hh_l{k}"] = dgh2.T @ hprev_all.reshape(-1, d)
        grads[f"dec.gru.bias_ih_l{k}"] = dgi2.sum(0)
        grads[f"dec.gru.bias_hh_l{k}"] = dgh2.sum(0)
        dy = (dgi2 @ Wih).reshape(B, L, d)            # gradient w.r.t. this layer's input
    # token-embedding gather backward (scatter-add; tok_emb has no padding_idx)
    np.add.at(grads["dec.tok_emb.weight"], seq_in.reshape(-1), dy.reshape(-1, d))

    return dy, dh0_total


def elbo_step(params, cfg, triples, seq, eps, beta, drop_masks=None,
              n_tok_global=None, batch_global=None):
    """Forward + hand-written backward of ``loss = CE + beta*KL`` (ablation_study.py:59-75).

    ``n_tok_global`` / ``batch_global`` replace the local CE / KL normalisers; the data-parallel
    path uses them so that SUMMED rank gradients equal the single-process gradient on the
    concatenated batch.  Returns (losses dict, grads dict keyed like params).
    """
    tied = cfg.get("tie_weights", True)
    pad_rid, pad_eid = cfg.get("pad_rid"), cfg.get("pad_eid")
    fw = elbo_forward(params, cfg, triples, seq, eps, beta, drop_masks)
    enc, dec = fw["enc"], fw["dec"]
    seq_in, tgt = seq[:, :-1], seq[:, 1:]
    B, L = seq_in.shape
    Wt = params["dec.tok_emb.weight"]
    V, d = Wt.shape
    dz_dim = enc["mu"].shape[1]
    nl = n_gru_layers(params)
    dt = Wt.dtype
    grads = {k: np.zeros_like(v) for k, v in params.items()}

    n_tok = fw["n_tok"] if n_tok_global is None else float(n_tok_global)
    b_glob = B if batch_global is None else int(batch_global)
    if n_tok_global is not None or batch_global is not None:
        ce_local = fw["ce"] * fw["n_tok"] / n_tok
        kl_local = fw["kl"] * B / b_glob
        fw = dict(fw, ce=ce_local, kl=kl_local, loss=ce_local + beta * kl_local)

    dy, dh0_total = _decoder_backward(params, fw, seq_in, tgt, n_tok, tied, drop_masks, grads)

    # ---- h0 = tanh(W_z z + b_z) shared by all layers ----
    dpre = dh0_total * (1.0 - dec["h0"] ** 2)
    grads["dec.z_proj.weight"] = dpre.T @ enc["z"]
    grads["dec.z_proj.bias"] = dpre.sum(0)
    dz = dpre @ params["dec.z_proj.weight"]

    # ---- reparameterisation + KL (SURVEY.md Appendix A "Backward highlights") ----
    kscale = beta / (b_glob * dz_dim)
    dmu = dz + kscale * enc["mu"]
    dlogv = 0.5 * dz * eps * enc["sigma"] + 0.5 * kscale * (np.exp(enc["logv"]) - 1.0)
    dlogv = dlogv * ((enc["logv_raw"] >= -10.0) & (enc["logv_raw"] <= 10.0))
    a_last = enc["acts"][-1]
    grads["enc.mu.weight"] = dmu.T @ a_last
    grads["enc.mu.bias"] = dmu.sum(0)
    grads["enc.logv.weight"] = dlogv.T @ a_last
    grads["enc.logv.bias"] = dlogv.sum(0)
    da = dmu @ params["enc.mu.weight"] + dlogv @ params["enc.logv.weight"]
    for k in range(n_mlp_layers(params) - 1, -1, -1):
        dp = da * gelu_erf_grad(enc["pre"][k])
        grads[f"enc.mlp.{2 * k}.weight"] = dp.T @ enc["acts"][k]
        grads[f"enc.mlp.{2 * k}.bias"] = dp.sum(0)
        da = dp @ params[f"enc.mlp.{2 * k}.weight"]

    # ---- masked mean-pool + embedding gather backward ----
    dg = da / enc["cnt"]                                               # [B,3d]
    dE, dR = grads["enc.e_emb.weight"], grads["enc.r_emb.weight"]
    dm = enc["mask"]
    hh, rr, tt = triples[:, :, 0], triples[:, :, 1], triples[:, :, 2]
    rep = np.broadcast_to(dg[:, None, :], (B, triples.shape[1], 3 * d))
    np.add.at(dE, hh[dm], rep[dm][:, :d])
    np.add.at(dR, rr[dm], rep[dm][:, d:2 * d])
    np.add.at(dE, tt[dm], rep[dm][:, 2 * d:])
    if pad_eid is not None:
        dE[pad_eid] = 0.0          # nn.Embedding(padding_idx=...) never accumulates a gradient
    if pad_rid is not None:
        dR[pad_rid] = 0.0
    if tied and "dec.out.weight" in grads:
        grads["dec.out.weight"] = grads["dec.tok_emb.weight"]
    losses = {"loss": float(fw["loss"]), "ce": float(fw["ce"]), "kl": float(fw["kl"]), "n_tok": n_tok}
    return losses, grads, fw


def ark_forward(params, cfg, seq):
    """ARK.forward + CE — kgvae/model/models.py:340-346,395-405; kgvae/experiments/train.py:44-52."""
    tied = cfg.get("tie_weights", True)
    dec = gru_decoder_forward(params, None, seq[:, :-1], None, tied, decoder_only=True)
    ce, lse, valid, n_tok = cross_entropy_ignore_pad(dec["logits"], seq[:, 1:])
    return {"dec": dec, "ce": ce, "kl": 0.0, "loss": ce, "lse": lse, "valid": valid, "n_tok": n_tok}


def ark_step(params, cfg, seq, n_tok_global=None):
    """Decoder-only CE step (train.py:42-58): forward + hand-written backward.  Returns (losses, grads, fw)."""
    tied = cfg.get("tie_weights", True)
    fw = ark_forward(params, cfg, seq)
    seq_in, tgt = seq[:, :-1], seq[:, 1:]
    B, L = seq_in.shape
    d = params["dec.tok_emb.weight"].shape[1]
    grads = {k: np.zeros_like(v) for k, v in params.items()}
    n_tok = fw["n_tok"] if n_tok_global is None else float(n_tok_global)
    if n_tok_global is not None:
        fw = dict(fw, ce=fw["ce"] * fw["n_tok"] / n_tok)
        fw["loss"] = fw["ce"]
    dy, _ = _decoder_backward(params, fw, seq_in, tgt, n_tok, tied, None, grads)
    grads["dec.pos_emb.weight"][:L] = dy.sum(0)          # pos_emb(arange(L)) broadcast over the batch
    if tied and "dec.out.weight" in grads:
        grads["dec.out.weight"] = grads["dec.tok_emb.weight"]
    return {"loss": float(fw["loss"]), "ce": float(fw["ce"]), "kl": 0.0, "n_tok": n_tok}, grads, fw


def adam_step(p, g, m, v, step, lr, beta1=0.9, beta2=0.999, eps=1e-8):
    """torch.optim.Adam (no weight decay, no amsgrad) — ablation_study.py:571 — one tensor.

    ``step`` is the 1-based step count AFTER this update.  Returns (p, m, v).
    """
    m = beta1 * m + (1.0 - beta1) * g
    v = beta2 * v + (1.0 - beta2) * g * g
    bc1 = 1.0 - beta1 ** step
    bc2 = 1.0 - beta2 ** step
    denom = np.sqrt(v) / math.sqrt(bc2) + eps
    return p - (lr / bc1) * m / denom, m, v


def cosine_lr(base_lr, epoch, t_max, eta_min=1e-6):
    """CosineAnnealingLR closed form — ablation_study.py:577-581 (stepped once per epoch)."""
    return eta_min + (base_lr - eta_min) * (1.0 + math.cos(math.pi * epoch / t_max)) / 2.0


def beta_schedule(beta0, beta1, epoch, num_epochs):
    """ablation_study.py:589-591."""
    return beta0 + (beta1 - beta0) * epoch / num_epochs


def beam_generate(dec_logits_fn, z, layout, beam=4):
    """SAIL.beam_generate — kgvae/model/models.py:283-300.

    ``dec_logits_fn(z, prefix[B,l]) -> logits[B,l,V]``.  The beam is shared by the whole batch
    and ranked by BATCH-MEAN log-prob (models.py:296) — a reference quirk kept on purpose.
    Returns a list (per batch row) of integer triples.
    """
    B = z.shape[0]
    seqs = [(np.full((B, 1), BOS, dtype=np.int64), np.zeros(B))]
    for _ in range(layout["seq_len"] - 1):
        cand = []
        for s, lp in seqs:
            lg = np.asarray(dec_logits_fn(z, s))[:, -1].astype(np.float64)
            m = lg.max(1, keepdims=True)
            logp = lg - m - np.log(np.exp(lg - m).sum(1, keepdims=True))
            ids = np.argsort(-logp, axis=1, kind="stable")[:, :beam]
            for k in range(beam):
                cand.append((np.concatenate([s, ids[:, k:k + 1]], 1),
                             lp + logp[np.arange(B), ids[:, k]]))
        order = sorted(range(len(cand)), key=lambda i: cand[i][1].mean(), reverse=True)[:beam]
        seqs = [cand[i] for i in order]
        if all((s[:, -1] == EOS).all() for s, _ in seqs):
            break
    return [seq_to_triples(row, layout) for row in seqs[0][0]]
