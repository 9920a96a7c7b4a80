"""Drop-in ``kgvae.model.models`` for the KG-VAE hot path, backed by ark_b200's sm_100a kernels.

Same public surface as the reference module (/root/reference/kgvae/model/models.py): ``SAIL(config)`` with
``.enc(triples) -> (z, mu, logv)``, ``.dec(z, tgt) -> logits``, ``.forward(triples, seq_in)``, ``.kl_mean``,
``.decode_latent`` / ``.beam_generate`` / ``.generate_test_graphs`` / ``.posterior_bits`` /
``.bits_per_sequence`` / ``.count_unique_graphs``, identical ``state_dict`` keys and shapes
(models.py:26-27,36,43-44,119-132), identical initialisation (torch's nn.Embedding / nn.Linear / nn.GRU are
used as PARAMETER CONTAINERS only — none of their forward methods runs).

What changed underneath:
  * ``enc`` / ``dec`` / ``forward`` execute on the GPU through the fp32 kernels of libarkb200 (inference:
    generation, validation, beam search — integer outputs must match the reference, so no bf16 here);
  * training goes through ``SAIL.elbo_step`` (new) = ark_b200.elbo.SailEngine: bf16 tcgen05 GEMMs, fused
    softmax-CE, PAD-free packed rows, fused Adam.  ``forward`` never builds an autograd graph.
There is no CPU implementation: modules must live on a CUDA device before they are called.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from ark_b200 import ops
from ark_b200.elbo import SailEngine
from ark_b200.layout import PackedLayout, pack_layout, pack_tlayout
from kgvae.model.utils import canonical_graph_string

K, MN = ops.MAJOR_K, ops.MAJOR_MN


def _need_cuda(t, what):
    if not t.is_cuda:
        raise RuntimeError(f"{what}: this build of kgvae runs on CUDA only (ark_b200 has no CPU path); "
                           "move the model and inputs to a B200 with .to('cuda')")


def _linear_f32(x, lin, epilogue=ops.EPI_NONE):
    """y = epi(x W^T + b) in fp32 on the SIMT kernel (reference: nn.Linear, models.py:36,43-44,120,128)."""
    M, Kd = x.shape
    N = lin.weight.shape[0]
    y = torch.empty(M, N, device=x.device, dtype=torch.float32)
    ops.gemm(x, K, lin.weight.detach(), K, y, M, N, Kd, bias=None if lin.bias is None else lin.bias.detach(),
             epilogue=epilogue, backend="simt")
    return y


class AutoRegEncoderMLP(nn.Module):
    """Embedding gather + masked mean-pool + GELU MLP + (mu, logv) heads + reparameterisation
    (reference: models.py:13-64)."""

    def __init__(self, num_entities, num_relations, d_model, latent_dim, pad_eid=None, pad_rid=None, hidden=None,
                 dropout=0.0, n_layers=2):
        super().__init__()
        if dropout:
            raise NotImplementedError("the reference never enables encoder dropout for SAIL (models.py:151-159)")
        self.pad_rid, self.pad_eid = pad_rid, pad_eid
        self.e_emb = nn.Embedding(num_entities, d_model, padding_idx=pad_eid)
        self.r_emb = nn.Embedding(num_relations, d_model, padding_idx=pad_rid)
        d_in = 3 * d_model
        hidden = hidden or max(d_in, 2 * d_model)
        mods, width = [], d_in
        for _ in range(n_layers):
            mods += [nn.Linear(width, hidden), nn.GELU()]   # indices 0,2,4 hold the weights (state_dict names)
            width = hidden
        self.mlp = nn.Sequential(*mods)
        self.mu = nn.Linear(hidden, latent_dim)
        self.logv = nn.Linear(hidden, latent_dim)

    @torch.no_grad()
    def encode_stats(self, triples):
        _need_cuda(triples, "enc")
        B = triples.shape[0]
        d3 = 3 * self.e_emb.weight.shape[1]
        g = torch.empty(B, d3, device=triples.device)
        inv = torch.empty(B, device=triples.device)
        ops.gather_pool_fwd(triples.contiguous(), None, self.e_emb.weight.detach(), self.r_emb.weight.detach(),
                            self.pad_rid, g, None, inv)
        for m in self.mlp:
            if isinstance(m, nn.Linear):
                g = _linear_f32(g, m, ops.EPI_GELU)
        return _linear_f32(g, self.mu), _linear_f32(g, self.logv).clamp_(-10, 10)

    eps_hook = None     # tests: callable(mu) -> eps replacing the randn_like draw (CPU-generator fixtures)

    @torch.no_grad()
    def forward(self, triples):
        mu, logv = self.encode_stats(triples)
        eps = torch.randn_like(mu) if self.eps_hook is None else self.eps_hook(mu)   # the reference's RNG call (models.py:63)
        z = mu + eps * torch.exp(0.5 * logv)
        return z, mu, logv


def _gru_stack_f32(gru, x, h0, B, Lp, return_states=False):
    """fp32 multi-layer GRU over time-major rows (t, b) on the library's kernels (eval semantics).
    `h0`: one [B, d] tensor shared by every layer (reference models.py:140) or a list with one per layer."""
    dev = x.device
    d = x.shape[1]
    N = B * Lp
    bt = np.full(Lp, B, dtype=np.int32)
    off = (np.arange(Lp + 1, dtype=np.int32) * B).astype(np.int32)
    gh_ws = torch.empty(B, 3 * d, device=dev)
    u = x
    finals = []
    for k in range(gru.num_layers):
        w_ih, w_hh = getattr(gru, f"weight_ih_l{k}").detach(), getattr(gru, f"weight_hh_l{k}").detach()
        b_ih, b_hh = getattr(gru, f"bias_ih_l{k}").detach(), getattr(gru, f"bias_hh_l{k}").detach()
        gi = torch.empty(N, 3 * d, device=dev)
        ops.gemm(u, K, w_ih, K, gi, N, 3 * d, d, bias=b_ih, backend="simt")
        hp = torch.empty(N, d, device=dev)
        hp[:B].copy_(h0[k] if isinstance(h0, (list, tuple)) else h0)
        y = torch.empty(N, d, device=dev)
        ops.gru_layer_fwd(None, hp, w_hh, gi, b_hh, bt, off, Lp, d, y, None, None, gh_ws, 0)
        u = y
        finals.append(y[N - B:])
    return (u, finals) if return_states else u


class AutoRegDecoderGRU(nn.Module):
    """Token embedding + h0 = tanh(z_proj z) + n-layer GRU + tied vocabulary projection
    (reference: models.py:116-142)."""

    def __init__(self, d_model, num_layers, seq_len, vocab_size, latent_dim, dropout=0.1, tie_weights=True):
        super().__init__()
        self.tok_emb = nn.Embedding(vocab_size, d_model)
        self.z_proj = nn.Linear(latent_dim, d_model)
        self.gru = nn.GRU(input_size=d_model, hidden_size=d_model, num_layers=num_layers, batch_first=True,
                          dropout=dropout if num_layers > 1 else 0.0)
        self.out = nn.Linear(d_model, vocab_size)
        if tie_weights and self.out.weight.shape == self.tok_emb.weight.shape:
            self.out.weight = self.tok_emb.weight

    @torch.no_grad()
    def forward(self, z, tgt):
        """logits [B, L', V] for any prefix length L' (fp32, eval semantics: no inter-layer dropout)."""
        _need_cuda(tgt, "dec")
        if self.training and self.gru.dropout > 0:
            raise RuntimeError("dec() is the fp32 inference path; train with SAIL.elbo_step (dropout lives there)")
        B, Lp = tgt.shape
        d = self.tok_emb.weight.shape[1]
        tok = tgt.t().contiguous().view(-1).to(torch.int32)           # time-major rows (t, b)
        x = torch.empty(B * Lp, d, device=tgt.device)
        ops.tok_gather_fwd(self.tok_emb.weight.detach(), tok, x, None)
        h0 = _linear_f32(z.to(torch.float32).contiguous(), self.z_proj, ops.EPI_TANH)
        u = _gru_stack_f32(self.gru, x, h0, B, Lp)
        logits = _linear_f32(u, self.out)
        return logits.view(Lp, B, -1).transpose(0, 1).contiguous()

    # ---- incremental decoding: one GRU step per generated token instead of re-decoding the prefix
    @torch.no_grad()
    def init_state(self, z):
        """Per-layer hidden state before any token: h0 = tanh(z_proj z) for every layer (models.py:139-140)."""
        h0 = _linear_f32(z.to(torch.float32).contiguous(), self.z_proj, ops.EPI_TANH)
        return [h0] * self.gru.num_layers

    @torch.no_grad()
    def step(self, tok, state, pos=None):
        """(logits [B, V] of the NEXT token, new state) after consuming `tok` [B] — bit-identical to
        `self(z, prefix)[:, -1]` with `state` = the state after prefix[:, :-1] (same kernels, same operation order)."""
        B = tok.shape[0]
        d = self.tok_emb.weight.shape[1]
        x = torch.empty(B, d, device=tok.device)
        ops.tok_gather_fwd(self.tok_emb.weight.detach(), tok.to(torch.int32).contiguous(), x, None)
        u, new_state = _gru_stack_f32(self.gru, x, state, B, 1, return_states=True)
        return _linear_f32(u, self.out), new_state


class AutoRegEncoder(nn.Module):
    """t-SAIL encoder PARAMETERS (reference: models.py:66-76): same modules, names, shapes and initialisation as
    the reference so state_dicts interchange; the forward pass lives in ark_b200.tsail (fused ELBO step)."""

    def __init__(self, num_entities, num_relations, d_model, nhead, latent_dim, pad_eid=None, pad_rid=None, n_layers=2):
        super().__init__()
        self.pad_rid = pad_rid
        self.e_emb = nn.Embedding(num_entities, d_model, padding_idx=pad_eid)
        self.r_emb = nn.Embedding(num_relations, d_model, padding_idx=pad_rid)
        layer = nn.TransformerEncoderLayer(d_model * 3, nhead, batch_first=True)
        self.txf = nn.TransformerEncoder(layer, n_layers)
        self.mu = nn.Linear(d_model * 3, latent_dim)
        self.logv = nn.Linear(d_model * 3, latent_dim)

    _owner = None     # the SAIL module whose engine runs the kernels (set by SAIL.__init__, not a submodule)

    @torch.no_grad()
    def encode_stats(self, triples):
        _need_cuda(triples, "enc")
        return self._owner().engine().encode_stats(triples)

    @torch.no_grad()
    def forward(self, triples):
        """(z, mu, logv), eval semantics (no dropout), fp32 kernels; no clamp on logv (reference models.py:92-94)."""
        if self.training and self._owner().engine().p_drop > 0:
            raise RuntimeError("enc() is the fp32 inference path; train with SAIL.elbo_step (dropout lives there)")
        mu, logv = self.encode_stats(triples)
        z = mu + torch.randn_like(mu) * torch.exp(0.5 * logv)
        return z, mu, logv


class AutoRegDecoder(nn.Module):
    """t-SAIL decoder PARAMETERS (reference: models.py:98-106)."""

    def __init__(self, d_model, nhead, num_layers, seq_len, vocab_size, latent_dim):
        super().__init__()
        self.tok_emb = nn.Embedding(vocab_size, d_model)
        self.pos_emb = nn.Embedding(seq_len, d_model)
        self.z_proj = nn.Linear(latent_dim, d_model)
        layer = nn.TransformerDecoderLayer(d_model, nhead, batch_first=True)
        self.txf = nn.TransformerDecoder(layer, num_layers)
        self.out = nn.Linear(d_model, vocab_size)

    _owner = None

    @torch.no_grad()
    def forward(self, z, tgt):
        """logits [B, L', V] for any prefix length (fp32, eval semantics)."""
        _need_cuda(tgt, "dec")
        if self.training and self._owner().engine().p_drop > 0:
            raise RuntimeError("dec() is the fp32 inference path; train with SAIL.elbo_step (dropout lives there)")
        return self._owner().engine().decode_logits(z, tgt)



class _ForwardFn(torch.autograd.Function):
    """Differentiable ``SAIL.forward`` / ``ARK.forward`` (reference models.py:317-320,395-405) on the fused engine.

    The reference trains by ``logits, mu, logv = model(triples, seq_in); loss = CE + b*KL; loss.backward()``
    (ablation_study.py:63-75).  This node makes that loop work unmodified on the CUDA kernels: forward runs the
    engine's bf16 tensor-core path over ALL B x L positions (PAD positions included: the reference returns logits
    for them too) and returns fp32 ``[B, L, V]`` logits; backward receives d(loss)/d(logits, mu, logv), writes the
    logit gradient into the engine's packed buffer and resumes the engine's hand-written backward pass; the
    parameter gradients come back through autograd (``p.grad`` accumulation semantics of torch are kept).
    The materialised fp32 logits make this the COMPATIBILITY path (the fused ``elbo_step`` never builds them)."""

    @staticmethod
    def forward(ctx, model, triples, seq_in, eps, *params):
        from ark_b200.layout import _uniform_layout
        eng = model.engine()
        eng.refresh_shadow()                # a torch optimiser may have stepped the fp32 masters since the last call
        dev = eng.device
        B, L = seq_in.shape
        V, ldv = eng.V, eng.ldv
        seq = torch.cat([seq_in.to(dev), seq_in.new_zeros(B, 1).to(dev)], dim=1).contiguous()   # engine: inputs = seq[:, :-1]
        lay = _uniform_layout(B, L).to(dev)
        tri = None if triples is None else triples.to(dev).contiguous()
        gen = eng._fb_gen(tri, seq, lay, eps, 0.0, train=True, dropout=model.training, autograd=True)
        st = next(gen)
        ctx.gen, ctx.st, ctx.shape, ctx.eng = gen, st, (B, L, V, ldv), eng
        ctx.names = [n for n, _ in model.named_parameters()]
        logits = st["logits"].view(L, B, ldv)[:, :, :V].transpose(0, 1).float().contiguous()
        if st["heads"] is None:
            return (logits,)
        dz = eng.dz
        heads = st["heads"]
        return logits, heads[:, :dz].clone(), heads[:, dz:2 * dz].clamp(-10.0, 10.0)

    @staticmethod
    def backward(ctx, dlogits, dmu=None, dlogv=None):
        eng, st = ctx.eng, ctx.st
        B, L, V, ldv = ctx.shape
        buf = st["logits"].view(L, B, ldv)
        if dlogits is None:
            buf.zero_()
        else:
            buf[:, :, :V].copy_(dlogits.transpose(0, 1))
            if ldv > V:
                buf[:, :, V:].zero_()
        f = eng.flat
        keep, f.grad = f.grad, torch.empty_like(f.grad)      # a private buffer: p.grad (a view of flat.grad) keeps torch's
        try:                                                 # accumulate-into-.grad semantics
            try:
                ctx.gen.send({"dmu": None if dmu is None else dmu.contiguous().float(),
                              "dlogv": None if dlogv is None else dlogv.contiguous().float()})
            except StopIteration:
                pass
            eng._sync_grads()
            grads = tuple(f.g(n) for n in ctx.names)
        finally:
            f.grad = keep
            ctx.gen = ctx.st = None
        return (None, None, None, None) + grads


class _EngineMixin:
    """Plumbing between an nn.Module with the reference's parameters and its fused ark_b200 training engine."""

    def _init_engine_slot(self):
        self._engine = None
        self.register_load_state_dict_post_hook(lambda module, incompatible: module._weights_changed())

    def _attach_engine(self, engine):
        object.__setattr__(self, "_engine", engine)

    def _weights_changed(self):
        if self._engine is not None:
            self._engine.refresh_shadow()

    def engine(self, **kw) -> SailEngine:
        """The fused training engine bound to this module (created on first use; the module must be on CUDA)."""
        if self._engine is None:
            if self.config["model_type"] == "t-SAIL":
                from ark_b200.tsail import TSailEngine
                TSailEngine(self, **kw)     # attaches itself
            elif self.config["model_type"] == "t-ARK":
                from ark_b200.tsail import TArkEngine
                TArkEngine(self, **kw)
            else:
                SailEngine(self, **kw)      # attaches itself
        return self._engine

    def _make_layout(self, triples, seq):
        """Host-side PAD-skipping layout of one batch (time-major packed rows for the GRU models, graph-major
        ragged rows for the Transformer models)."""
        seq_cpu = seq if not seq.is_cuda else seq.cpu()
        if self.config["model_type"] in ("t-SAIL", "t-ARK"):
            if triples is None:     # decoder-only: no encoder rows
                tri_cpu, pad_rid = torch.zeros(seq_cpu.shape[0], 1, 3, dtype=torch.int64), None
            else:
                tri_cpu, pad_rid = (triples if not triples.is_cuda else triples.cpu()), self.config.get("pad_rid")
            return pack_tlayout(tri_cpu, seq_cpu, pad_rid).to(self.engine().device)
        return pack_layout(seq_cpu).to(self.engine().device)


class SAIL(_EngineMixin, nn.Module):
    """The KG-VAE.  Reference: models.py:144-320."""

    def __init__(self, config):
        super().__init__()
        self.config = config
        mt = config["model_type"]
        if mt == "SAIL":
            self.enc = AutoRegEncoderMLP(
                num_entities=config["n_entities"], num_relations=config["n_relations"], d_model=config["d_model"],
                latent_dim=config["d_latent"], pad_eid=config.get("pad_eid"), pad_rid=config.get("pad_rid"),
                n_layers=config["n_layers"])
            self.dec = AutoRegDecoderGRU(
                d_model=config["d_model"], num_layers=config["n_layers"], seq_len=config["seq_len"],
                vocab_size=config["vocab_size"], latent_dim=config["d_latent"],
                dropout=config.get("dec_dropout", 0.1), tie_weights=config.get("tie_weights", True))
        elif mt == "t-SAIL":
            self.enc = AutoRegEncoder(
                num_entities=config["n_entities"], num_relations=config["n_relations"], d_model=config["d_model"],
                nhead=config["n_heads"], latent_dim=config["d_latent"], pad_eid=config.get("pad_eid"),
                pad_rid=config.get("pad_rid"), n_layers=config.get("n_layers", 2))
            self.dec = AutoRegDecoder(
                d_model=config["d_model"], nhead=config["n_heads"], num_layers=config["n_layers"],
                seq_len=config["seq_len"], vocab_size=config["vocab_size"], latent_dim=config["d_latent"])
            import weakref
            self.enc._owner = self.dec._owner = weakref.ref(self)      # their forward() runs on this module's engine
        else:
            raise NotImplementedError(f"Unknown model_type: {mt}")
        self._init_engine_slot()

    def elbo_backward(self, triples, seq, beta, eps=None, layout: PackedLayout = None,
                      n_tok_global=None, batch_global=None):
        """NEW fused entry: forward + backward of ``CE + beta*KL`` (ablation_study.py:59-75) that never
        materialises [B,L,V] probabilities; gradients land in every parameter's ``.grad`` (overwritten).
        ``triples``/``seq`` are the reference's LongTensors (host or device); ``eps`` defaults to
        ``torch.randn(B, d_latent)`` from torch's global CUDA generator — the draw the reference makes at
        models.py:63.  Returns a device tensor [ce, kl] (no host sync)."""
        eng = self.engine()
        if layout is None:
            layout = self._make_layout(triples, seq)
        triples = triples.to(eng.device, non_blocking=True).contiguous()
        seq = seq.to(eng.device, non_blocking=True).contiguous()
        if eps is None:
            eps = torch.randn(triples.shape[0], self.config["d_latent"], device=eng.device)
        out = eng.forward_backward(triples, seq, layout, eps.contiguous(), float(beta), n_tok_global, batch_global)
        eng.stats[0:2] += out
        eng.stats[2] += 1
        return out

    def elbo_step(self, triples, seq, beta, eps=None, layout: PackedLayout = None, lr=None,
                  n_tok_global=None, batch_global=None, graph=False):
        """elbo_backward + the fused Adam update: one full optimisation step (ablation_study.py:43,59-76).
        ``graph=True`` replays a CUDA graph captured per batch layout (fixed-size datasets such as syn-*)."""
        eng = self.engine()
        if layout is None:
            layout = self._make_layout(triples, seq)
        if eps is None:
            eps = torch.randn(triples.shape[0], self.config["d_latent"], device=eng.device)
        if graph:
            return eng.train_step_graphed(triples, seq, layout, eps, float(beta), lr, n_tok_global, batch_global)
        triples = triples.to(eng.device, non_blocking=True).contiguous()
        seq = seq.to(eng.device, non_blocking=True).contiguous()
        # eager step: Adam runs bucket by bucket on a side stream while backward is still going
        return eng.train_step(triples, seq, layout, eps.contiguous(), float(beta), lr, n_tok_global, batch_global)

    # ---- reference interface -----------------------------------------------------------------------
    def kl_mean(self, mu, logv):
        return -0.5 * torch.mean(1 + logv - mu.pow(2) - logv.exp())

    eps_hook = None     # tests: callable(B, dz, device) -> eps, replacing the randn draw of the differentiable forward

    def forward(self, triples, seq_in):
        """(logits [B, L, V], mu, logv) — reference models.py:317-320.  In train() mode with autograd enabled this is a differentiable
        node on the fused engine (`_ForwardFn`: the reference's own loop `model(...)`, `loss.backward()`,
        `optimizer.step()` trains through it); in eval() mode or under `torch.no_grad()` it is the fp32 inference path whose integer
        outputs (beam search, generation) match the reference bit for bit."""
        if (self.training and torch.is_grad_enabled() and self.config["model_type"] == "SAIL"
                and any(p.requires_grad for p in self.parameters())):
            _need_cuda(triples, "forward")
            B, dz = triples.shape[0], self.config["d_latent"]
            dev = self.engine().device
            # the reference's draw: torch.randn_like(mu), first RNG call of the step (models.py:63)
            eps = self.eps_hook(B, dz, dev) if self.eps_hook is not None else torch.randn(B, dz, device=dev)
            return _ForwardFn.apply(self, triples, seq_in, eps.contiguous(), *self.parameters())
        z, mu, logv = self.enc(triples)
        return self.dec(z, seq_in), mu, logv

    def bits_per_sequence(self, seq, z, pad_id=0):
        """AR bits of one sequence under teacher forcing (reference: models.py:202-213).  The GRU decoder is
        causal, so one pass over the full prefix gives every per-position distribution the reference obtains
        from its O(L^2) loop of growing prefixes."""
        seq = seq.unsqueeze(0).to(z.device)
        n = int((seq[0, 1:] != pad_id).long().cumprod(0).sum().item())   # stop at the first PAD target
        if n == 0:
            return 0.0
        logp = F.log_softmax(self.dec(z, seq[:, :n]), dim=-1)[0]
        tgt = seq[0, 1:n + 1]
        return float(-(logp[torch.arange(n, device=z.device), tgt]).sum().item() / math.log(2))

    @torch.no_grad()
    def posterior_bits(self, dataset, device, pad_id=0, sample_frac=0.1, desc="posterior bits"):
        """Reference: models.py:218-260."""
        ln2 = math.log(2)
        n = max(1, int(sample_frac * len(dataset)))
        records = []
        for i in range(n):
            triples, seq = dataset[i]
            triples = triples.unsqueeze(0).to(device)
            z, mu, logv = self.enc(triples)
            ar = self.bits_per_sequence(seq.to(device), z, pad_id)
            kl = float((-0.5 * torch.sum(1 + logv - mu.pow(2) - logv.exp(), dim=1) / ln2).item())
            records.append({"ar_bits": ar, "kl_bits": kl, "total_bits": ar + kl})
        tot = np.array([r["total_bits"] for r in records])
        return {"avg_total_bits": float(tot.mean()), "avg_ar_bits": float(np.mean([r["ar_bits"] for r in records])),
                "avg_kl_bits": float(np.mean([r["kl_bits"] for r in records])), "min_total_bits": float(tot.min()),
                "max_total_bits": float(tot.max()), "records": records}

    @torch.no_grad()
    def decode_latent(self, z, seq_len, special_tokens, seq_to_triples, ent_base, rel_base, beam=4):
        self.eval()
        z = z.to(next(self.parameters()).device, dtype=torch.float32)
        return self.beam_generate(seq_len, special_tokens, seq_to_triples, z, ent_base, rel_base, beam=beam)

    @torch.no_grad()
    def count_unique_graphs(self, latent_dim, decode_latent_fn, num_samples=1000, beam=1):
        self.eval()
        z = torch.randn((num_samples, latent_dim), device=next(self.parameters()).device)
        uniq = {canonical_graph_string(g) for g in decode_latent_fn(z, beam=beam)}
        print(f"\n[Graph Diversity from {num_samples} Random Latents]")
        print(f"  Unique graphs generated: {len(uniq)}")
        print(f"  Diversity ratio: {len(uniq) / num_samples:.3f}")
        return uniq

    @torch.no_grad()
    def beam_generate(self, seq_len, special_tokens, seq_to_triples, z, ent_base, rel_base, beam=4):
        """Batch-shared beam search ranked by the batch-MEAN log-probability — the reference's exact procedure
        (models.py:283-300).  The reference re-decodes the whole prefix for every candidate and step (O(L^2) decoder
        work); here every beam carries its GRU hidden states and a step consumes ONE token (same kernels, same
        arithmetic, so the integer outputs are unchanged — tests/test_elbo_gpu.py)."""
        dev, B = z.device, z.size(0)
        eos = special_tokens["EOS"]
        incremental = hasattr(self.dec, "step")      # GRU decoder: hidden-state cache (one step per token, not O(L^2))
        state0 = self.dec.init_state(z) if incremental else None
        beams = [(torch.full((B, 1), special_tokens["BOS"], dtype=torch.long, device=dev), torch.zeros(B, device=dev), state0)]
        for _ in range(seq_len - 1):
            grown = []
            for prefix, score, state in beams:
                if incremental:      # `state` = the decoder state after prefix[:, :-1]
                    logits, nstate = self.dec.step(prefix[:, -1], state)
                else:
                    logits, nstate = self.dec(z, prefix)[:, -1], None
                logp = F.log_softmax(logits, dim=-1)
                best, idx = logp.topk(beam, dim=-1)
                grown += [(torch.cat([prefix, idx[:, j:j + 1]], 1), score + best[:, j], nstate) for j in range(beam)]
            beams = sorted(grown, key=lambda c: c[1].mean().item(), reverse=True)[:beam]
            if all(bool((p[:, -1] == eos).all()) for p, _, _ in beams):
                break
        return [seq_to_triples(row, special_tokens, ent_base, rel_base) for row in beams[0][0].cpu()]

    @torch.no_grad()
    def generate_test_graphs(self, test_loader, seq_len, special_tokens, seq_to_triples, ent_base, rel_base,
                             beam_width=4, num_generated_test_graphs=1000, device="cuda"):
        graphs = []
        for triples, _ in test_loader:
            z, *_ = self.enc(triples.to(device))
            graphs.extend(self.beam_generate(seq_len, special_tokens, seq_to_triples, z, ent_base, rel_base,
                                             beam=beam_width))
            if len(graphs) >= num_generated_test_graphs:
                return graphs[:num_generated_test_graphs]
        return graphs


class DecoderOnlyGRU(nn.Module):
    """tok_emb + pos_emb -> n-layer GRU (h0 = 0) -> tied vocabulary projection (reference: models.py:323-346)."""

    def __init__(self, d_model, num_layers, seq_len, vocab_size, dropout=0.1, tie_weights=True):
        super().__init__()
        self.tok_emb = nn.Embedding(vocab_size, d_model)
        self.pos_emb = nn.Embedding(seq_len, d_model)
        self.gru = nn.GRU(input_size=d_model, hidden_size=d_model, num_layers=num_layers, batch_first=True,
                          dropout=dropout if num_layers > 1 else 0.0)
        self.out = nn.Linear(d_model, vocab_size)
        if tie_weights and self.out.weight.shape == self.tok_emb.weight.shape:
            self.out.weight = self.tok_emb.weight

    @torch.no_grad()
    def forward(self, seq_in):
        _need_cuda(seq_in, "dec")
        if self.training and self.gru.dropout > 0:
            raise RuntimeError("dec() is the fp32 inference path; train with ARK.ce_backward (dropout lives there)")
        B, Lp = seq_in.shape
        d = self.tok_emb.weight.shape[1]
        tok = seq_in.t().contiguous().view(-1).to(torch.int32)
        x = torch.empty(B * Lp, d, device=seq_in.device)
        ops.tok_gather_fwd(self.tok_emb.weight.detach(), tok, x, None)
        x = (x.view(Lp, B, d) + self.pos_emb.weight.detach()[:Lp, None, :]).view(B * Lp, d).contiguous()
        u = _gru_stack_f32(self.gru, x, torch.zeros(B, d, device=seq_in.device), B, Lp)
        logits = _linear_f32(u, self.out)
        return logits.view(Lp, B, -1).transpose(0, 1).contiguous()

    @torch.no_grad()
    def init_state(self, B, device):
        d = self.tok_emb.weight.shape[1]
        return [torch.zeros(B, d, device=device) for _ in range(self.gru.num_layers)]

    @torch.no_grad()
    def step(self, tok, state, pos):
        """(next-token logits [B, V], new state) after consuming `tok` [B] at position `pos` — bit-identical to
        `self(prefix)[:, -1]` (incremental decoding)."""
        B = tok.shape[0]
        d = self.tok_emb.weight.shape[1]
        x = torch.empty(B, d, device=tok.device)
        ops.tok_gather_fwd(self.tok_emb.weight.detach(), tok.to(torch.int32).contiguous(), x, None)
        x = (x + self.pos_emb.weight.detach()[pos][None, :]).contiguous()
        u, new_state = _gru_stack_f32(self.gru, x, state, B, 1, return_states=True)
        return _linear_f32(u, self.out), new_state


class DecoderOnlyTransformer(nn.Module):
    """t-ARK decoder PARAMETERS (reference: models.py:349-366): same modules / names / init; the fused CE step lives
    in ark_b200.tsail.TArkEngine."""

    def __init__(self, d_model, nhead, num_layers, seq_len, vocab_size, dropout=0.1, tie_weights=True):
        super().__init__()
        self.tok_emb = nn.Embedding(vocab_size, d_model)
        self.pos_emb = nn.Embedding(seq_len, d_model)
        layer = nn.TransformerEncoderLayer(d_model, nhead, batch_first=True, dropout=dropout)
        self.txf = nn.TransformerEncoder(layer, num_layers)
        self.out = nn.Linear(d_model, vocab_size)
        if tie_weights and self.out.weight.shape == self.tok_emb.weight.shape:
            self.out.weight = self.tok_emb.weight

    _owner = None

    @torch.no_grad()
    def forward(self, seq_in):
        _need_cuda(seq_in, "dec")
        if self.training and self._owner().engine().p_drop > 0:
            raise RuntimeError("dec() is the fp32 inference path; train with ARK.ce_step (dropout lives there)")
        return self._owner().engine().decode_logits(None, seq_in)


class ARK(_EngineMixin, nn.Module):
    """Decoder-only autoregressive model — the reference's default ``model_type`` (models.py:368-520).
    Training (``ce_backward`` / ``ce_step``) runs on the same fused engine as SAIL minus the encoder and KL."""

    def __init__(self, config):
        super().__init__()
        self.config = config
        if config["model_type"] == "ARK":
            self.dec = DecoderOnlyGRU(d_model=config["d_model"], num_layers=config["n_layers"], seq_len=config["seq_len"],
                                      vocab_size=config["vocab_size"], dropout=config.get("dec_dropout", 0.1),
                                      tie_weights=config.get("tie_weights", True))
        elif config["model_type"] == "t-ARK":
            self.dec = DecoderOnlyTransformer(d_model=config["d_model"], nhead=config["n_heads"],
                                              num_layers=config["n_layers"], seq_len=config["seq_len"],
                                              vocab_size=config["vocab_size"], dropout=config.get("dec_dropout", 0.1),
                                              tie_weights=config.get("tie_weights", True))
            import weakref
            self.dec._owner = weakref.ref(self)
        else:
            raise NotImplementedError(f"Unknown model_type: {config['model_type']}")
        self._init_engine_slot()

    def forward(self, triples_or_seq, seq_in=None):
        """forward(seq) or forward(triples, seq) — triples are ignored (reference models.py:395-405).  Differentiable
        on the fused engine in train() mode with autograd enabled (GRU model); fp32 inference path in eval() mode or
        under torch.no_grad()."""
        seq = triples_or_seq if seq_in is None else seq_in
        if (self.training and torch.is_grad_enabled() and self.config["model_type"] == "ARK"
                and any(p.requires_grad for p in self.parameters())):
            _need_cuda(seq, "forward")
            return _ForwardFn.apply(self, None, seq, None, *self.parameters())[0]
        return self.dec(seq)

    def ce_backward(self, seq, layout: PackedLayout = None, n_tok_global=None):
        """Fused CE forward + backward over packed rows (reference train step: train.py:42-58)."""
        eng = self.engine()
        if layout is None:
            layout = self._make_layout(None, seq)
        seq = seq.to(eng.device, non_blocking=True).contiguous()
        out = eng.forward_backward(None, seq, layout, None, 0.0, n_tok_global, None)
        eng.stats[0:2] += out
        eng.stats[2] += 1
        return out

    def ce_step(self, seq, layout: PackedLayout = None, lr=None, n_tok_global=None):
        """ce_backward + Adam (bucket by bucket, overlapped with backward): one optimisation step."""
        eng = self.engine()
        if layout is None:
            layout = self._make_layout(None, seq)
        seq = seq.to(eng.device, non_blocking=True).contiguous()
        return eng.train_step(None, seq, layout, None, 0.0, lr, n_tok_global, None)

    @torch.no_grad()
    def generate(self, seq_len, special_tokens, device=None, batch_size=1, beam=1, sample=False, temperature=1.0,
                 top_p=0.0, top_k=0):
        """Greedy or sampled generation (reference: models.py:408-471; its temperature / top-k / top-p filtering verbatim as
        torch ops on the GPU).  The GRU model decodes incrementally from cached hidden states instead of re-decoding
        the prefix at every step."""
        device = device or next(self.parameters()).device
        eos = special_tokens["EOS"]
        seq = torch.full((batch_size, 1), special_tokens["BOS"], dtype=torch.long, device=device)
        incremental = hasattr(self.dec, "step")      # GRU decoder: hidden-state cache, one step per token
        state = self.dec.init_state(batch_size, device) if incremental else None
        for t in range(seq_len - 1):
            if incremental:
                logits, state = self.dec.step(seq[:, -1], state, t)
            else:
                logits = self.dec(seq)[:, -1]
            if not sample:
                nxt = logits.argmax(dim=-1, keepdim=True)
            else:
                if temperature and temperature != 1.0:
                    logits = logits / float(temperature)
                probs = F.softmax(logits, dim=-1)
                if top_k and top_k > 0:
                    _, keep = probs.topk(top_k, dim=-1)
                    probs = probs * torch.zeros_like(probs).scatter_(-1, keep, 1.0)
                    probs = probs / probs.sum(dim=-1, keepdim=True).clamp_min(1e-12)
                if top_p and 0.0 < top_p < 1.0:
                    sp, si = probs.sort(dim=-1, descending=True)
                    drop = sp.cumsum(dim=-1) > top_p
                    drop[..., 1:] = drop[..., :-1].clone()
                    drop[..., 0] = False
                    sp = sp.masked_fill(drop, 0.0)
                    sp = sp / sp.sum(dim=-1, keepdim=True).clamp_min(1e-12)
                    nxt = si.gather(-1, torch.multinomial(sp, 1))
                else:
                    nxt = torch.multinomial(probs, 1)
            seq = torch.cat([seq, nxt], dim=1)
            if bool((seq[:, -1] == eos).all()):
                break
        if seq.size(1) < seq_len:
            seq = torch.cat([seq, torch.full((batch_size, seq_len - seq.size(1)), eos, dtype=torch.long, device=device)], 1)
        return seq[:, :seq_len]

    def bits_per_sequence(self, seq, pad_id=0):
        """AR bits under teacher forcing (reference: models.py:473-486); one causal pass instead of O(L^2)."""
        seq = seq.unsqueeze(0)
        n = int((seq[0, 1:] != pad_id).long().cumprod(0).sum().item())
        if n == 0:
            return 0.0
        logp = F.log_softmax(self(seq[:, :n]), dim=-1)[0]
        return float(-(logp[torch.arange(n, device=seq.device), seq[0, 1:n + 1]]).sum().item() / math.log(2))

    @torch.no_grad()
    def posterior_bits(self, dataset, device, pad_id=0, sample_frac=0.1, desc="Posterior compression"):
        """Reference: models.py:488-520 (decoder-only: KL = 0, total = AR)."""
        n = max(1, int(sample_frac * len(dataset)))
        records = []
        for i in range(n):
            ar = self.bits_per_sequence(dataset[i][1].to(device), pad_id=pad_id)
            records.append({"ar_bits": ar, "kl_bits": 0.0, "total_bits": ar})
        tot = np.array([r["total_bits"] for r in records])
        return {"avg_total_bits": float(tot.mean()), "avg_ar_bits": float(tot.mean()), "avg_kl_bits": 0.0,
                "min_total_bits": float(tot.min()), "max_total_bits": float(tot.max()), "records": records}
