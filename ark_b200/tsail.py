"""Fused ELBO step of the Transformer KG-VAE ('t-SAIL'): hand-scheduled forward + backward on libarkb200.

Reference: kgvae/model/models.py:66-95 (AutoRegEncoder: nn.TransformerEncoder over the [h|r|t] triple
embeddings, width D = 3*d_model, masked mean-pool, mu/logv WITHOUT clamp) and :98-114 (AutoRegDecoder:
tok_emb + pos_emb, memory = z_proj(z) repeated L times, nn.TransformerDecoder with a causal mask, untied
vocabulary projection); loss and optimiser as for SAIL (ablation_study.py:59-76).  Both stacks are PyTorch's
post-LN layers: x1 = LN(x + Drop(SelfAttn(x))), [x2 = LN(x1 + Drop(CrossAttn(x1, mem)))],
x3 = LN(x2 + Drop(W2 Drop(ReLU(W1 x2)))) with dim_feedforward 2048, dropout 0.1, eps 1e-5 (torch defaults —
the reference passes none of them).

What is different here (all exact for loss and gradients, see ark_b200/layout.py::TLayout):
  * rows are PAD-free and graph-major (ragged): encoder rows = real triples, decoder rows = real positions;
  * every memory row of a graph is the same vector, so the decoder's cross-attention softmax is uniform and the
    block collapses to out_proj(v_proj(mem)) broadcast over the graph's rows (its q/k projections receive
    exactly zero gradient); attention dropout on those uniform weights is a Binomial(L, 1-p)/((1-p)L) scale
    per (row, head), drawn explicitly;
  * dense projections run on the tcgen05 GEMM (bf16 operands, fp32 accumulate), the [n x n] score blocks on the
    batched ragged kernel of csrc/attn_ops.cu; residual stream, LayerNorm statistics, mu/logv/KL, loss and all
    parameter gradients stay fp32.
"""
from __future__ import annotations

import math

import torch

from . import ops
from .elbo import SailEngine, _up8
from .flat import FlatParams
from .layout import TLayout

K, MN = ops.MAJOR_K, ops.MAJOR_MN
TOK, SQ = ops.TOK, ops.SQ


def tsail_param_order(model):
    """Gradient-readiness order of t-SAIL's parameters in TSailEngine's backward pass."""
    named = dict(model.named_parameters())
    groups = []
    add = lambda *names: groups.append([(n, named[n]) for n in names])  # noqa: E731
    add("dec.out.bias")
    add("dec.out.weight")
    nl_d = len(model.dec.txf.layers)
    for l in range(nl_d - 1, -1, -1):
        p = f"dec.txf.layers.{l}."
        for n in ("norm3.weight", "norm3.bias", "linear2.weight", "linear2.bias", "linear1.weight", "linear1.bias",
                  "norm2.weight", "norm2.bias", "multihead_attn.out_proj.weight", "multihead_attn.out_proj.bias",
                  "multihead_attn.in_proj_weight", "multihead_attn.in_proj_bias", "norm1.weight", "norm1.bias",
                  "self_attn.out_proj.weight", "self_attn.out_proj.bias", "self_attn.in_proj_weight",
                  "self_attn.in_proj_bias"):
            add(p + n)
    add("dec.tok_emb.weight")
    add("dec.pos_emb.weight")
    add("dec.z_proj.weight")
    add("dec.z_proj.bias")
    add("enc.mu.weight", "enc.logv.weight")
    add("enc.mu.bias", "enc.logv.bias")
    nl_e = len(model.enc.txf.layers)
    for l in range(nl_e - 1, -1, -1):
        p = f"enc.txf.layers.{l}."
        for n in ("norm2.weight", "norm2.bias", "linear2.weight", "linear2.bias", "linear1.weight", "linear1.bias",
                  "norm1.weight", "norm1.bias", "self_attn.out_proj.weight", "self_attn.out_proj.bias",
                  "self_attn.in_proj_weight", "self_attn.in_proj_bias"):
            add(p + n)
    add("enc.r_emb.weight")
    add("enc.e_emb.weight")
    missing = set(named) - {n for g in groups for n, _ in g}
    if missing:
        raise RuntimeError(f"parameters without a slot in the flat layout: {sorted(missing)}")
    return groups


class TSailEngine(SailEngine):
    """Owns the flat parameters of one t-SAIL module and runs its ELBO step on one GPU (or one rank)."""
    MODEL_TYPE = "t-SAIL"

    @staticmethod
    def _dropout_from_config(cfg):
        return float(cfg.get("txf_dropout", 0.1))       # torch default of the reference's layers (models.py:73,104)

    @staticmethod
    def _param_order(model):
        return tsail_param_order(model)

    def __init__(self, model, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, gemm_backend="tc", dist_group=None,
                 bucket_mb=16.0, seed=0):
        cfg = model.config
        if cfg["model_type"] != self.MODEL_TYPE:
            raise NotImplementedError(f"{type(self).__name__} accelerates model_type '{self.MODEL_TYPE}'")
        dev = next(model.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("TSailEngine needs the model on a CUDA device: there is no CPU path")
        self.model, self.cfg, self.device = model, cfg, dev
        self.has_enc = hasattr(model, "enc")
        self.d, self.dz, self.V = cfg["d_model"], cfg.get("d_latent", 0), cfg["vocab_size"]
        self.D = 3 * self.d
        self.H = cfg["n_heads"]
        self.nl_e = len(model.enc.txf.layers) if self.has_enc else 0
        self.nl_d = len(model.dec.txf.layers)
        self.nl = self.nl_d
        self.ln_eps = float(model.dec.txf.layers[0].norm1.eps)
        self.pad_rid, self.pad_eid = cfg.get("pad_rid"), cfg.get("pad_eid")
        self.tied = model.dec.out.weight is model.dec.tok_emb.weight
        self.p_drop = self._dropout_from_config(cfg)
        if self.d % 8 or self.D % self.H or self.d % self.H or (self.D // self.H) % 4 or (self.d // self.H) % 4:
            raise ValueError("the Transformer models need d_model % 8 == 0 and head widths that are multiples of 4")
        self.flat = FlatParams(self._param_order(model), dev)
        self.lr, self.betas, self.eps = float(lr), betas, float(eps)
        self.step_count = 0
        self.backend = gemm_backend
        self.seed, self.philox_offset = int(seed), 0
        self.ldv = _up8(self.V)
        self.group = dist_group
        self.world = torch.distributed.get_world_size(dist_group) if dist_group is not None else 1
        self.bucket_elems = int(bucket_mb * (1 << 20) / 4)
        self.comm_stream = torch.cuda.Stream(device=dev, priority=-1)
        self._upd = None
        self._pending = []
        self._hold_comm, self._held = False, []
        self.prof = None
        self._capturing = False
        self._segment_break = None
        self._graphs = {}
        self.dyn_f = torch.zeros(2, device=dev)
        self.dyn_i = torch.zeros(1, device=dev, dtype=torch.int64)
        self.launches_replayed = 0
        self.force_unfused_gru = False
        self.gru_mode = "auto"
        self.stats = torch.zeros(4, device=dev)
        self.refresh_shadow()
        if hasattr(model, "_attach_engine"):
            model._attach_engine(self)

    # ------------------------------------------------------------------ small helpers
    def _new(self, *s, dtype=torch.float32):
        return torch.empty(*s, device=self.device, dtype=dtype)

    def _drop_off(self, n_elems):
        """Reserve a Philox counter range for one dropout op; returns its start offset."""
        o = self.philox_offset
        self.philox_offset += (int(n_elems) + 3) // 4
        return o

    def _lin(self, x_b, w, bias, out, tag, epilogue=ops.EPI_NONE):
        M, Kd = x_b.shape
        self._gemm(x_b, K, w, K, out, M, w.shape[0], Kd, tag=tag, bias=bias, epilogue=epilogue)
        return out

    def _lin_bwd(self, dy_b, x_b, w, gw, gb, dx, accumulate, tag, dy_f32=None):
        """dW = dY^T X, db = colsum(dY), dX (+)= dY W."""
        M, N = dy_b.shape
        Kd = x_b.shape[1]
        self._gemm(dy_b, MN, x_b, MN, gw, N, Kd, M, tag=tag + "_dW")
        ops.colsum(dy_f32 if dy_f32 is not None else dy_b, M, N, gb)
        if dx is not None:
            self._gemm(dy_b, K, w, MN, dx, M, Kd, N, tag=tag + "_dX", accumulate=accumulate)

    # ------------------------------------------------------------------ attention block
    def _self_attn_fwd(self, x_b, pre, seg, Dm, causal, S, p):
        f, H = self.flat, self.H
        hd = Dm // H
        n = x_b.shape[0]
        qkv = self._lin(x_b, f.s(pre + "in_proj_weight"), f.p(pre + "in_proj_bias"), self._new(n, 3 * Dm, dtype=torch.bfloat16),
                        "attn_in")
        with self._timed("attn_scores", flops=2.0 * seg.sq_total * H * hd):
            ops.attn_bgemm(qkv, TOK, False, 0, qkv, TOK, True, Dm, S, SQ, 0, seg, H, hd, 0, causal, 1.0 / math.sqrt(hd))
        P = self._new(seg.sq_total * H, dtype=torch.bfloat16)
        Pd = self._new(seg.sq_total * H, dtype=torch.bfloat16) if p > 0 else None
        ops.attn_softmax_fwd(S, seg, H, causal, p, self.seed, self._drop_off(seg.sq_total * H) if p > 0 else 0, None, P, Pd)
        o_b = self._new(n, Dm, dtype=torch.bfloat16)
        with self._timed("attn_apply", flops=2.0 * seg.sq_total * H * hd):
            ops.attn_bgemm(Pd if Pd is not None else P, SQ, False, 0, qkv, TOK, False, 2 * Dm, o_b, TOK, 0, seg, H, hd, 1,
                           causal, 1.0)
        a = self._lin(o_b, f.s(pre + "out_proj.weight"), f.p(pre + "out_proj.bias"), self._new(n, Dm), "attn_out")
        return a, (qkv, P, Pd, o_b)

    def _self_attn_bwd(self, d_a_b, x_b, saved, pre, seg, Dm, causal, S, p, dx):
        f, H = self.flat, self.H
        hd = Dm // H
        qkv, P, Pd, o_b = saved
        n = x_b.shape[0]
        do = self._new(n, Dm)
        self._lin_bwd(d_a_b, o_b, f.s(pre + "out_proj.weight"), f.g(pre + "out_proj.weight"), f.g(pre + "out_proj.bias"), do,
                      False, "attn_out")
        dqkv = self._new(n, 3 * Dm, dtype=torch.bfloat16)
        Pa = Pd if Pd is not None else P
        with self._timed("attn_bwd_gemms", flops=8.0 * seg.sq_total * H * hd):
            ops.attn_bgemm(Pa, SQ, True, 0, do, TOK, False, 0, dqkv, TOK, 2 * Dm, seg, H, hd, 1, causal, 1.0)      # dV = P^T dO
            ops.attn_bgemm(do, TOK, False, 0, qkv, TOK, True, 2 * Dm, S, SQ, 0, seg, H, hd, 0, causal, 1.0)        # dP = dO V^T
            dS = self._new(seg.sq_total * H, dtype=torch.bfloat16)
            ops.attn_softmax_bwd(P, Pd, S, seg, H, causal, p, 1.0 / math.sqrt(hd), dS)
            ops.attn_bgemm(dS, SQ, False, 0, qkv, TOK, False, Dm, dqkv, TOK, 0, seg, H, hd, 1, causal, 1.0)        # dQ = dS K
            ops.attn_bgemm(dS, SQ, True, 0, qkv, TOK, False, 0, dqkv, TOK, Dm, seg, H, hd, 1, causal, 1.0)         # dK = dS^T Q
        self._lin_bwd(dqkv, x_b, f.s(pre + "in_proj_weight"), f.g(pre + "in_proj_weight"), f.g(pre + "in_proj_bias"), dx,
                      True, "attn_in")

    # ------------------------------------------------------------------ residual + LayerNorm
    def _add_ln_fwd(self, branch, res, pre, p):
        f = self.flat
        n, Dm = branch.shape
        y, y_b = self._new(n, Dm), self._new(n, Dm, dtype=torch.bfloat16)
        mean, rstd = self._new(n), self._new(n)
        mask = self._new(n, Dm, dtype=torch.uint8) if p > 0 else None
        ops.add_layernorm_fwd(branch, res, f.p(pre + "weight"), f.p(pre + "bias"), self.ln_eps, p, self.seed,
                              self._drop_off(n * Dm) if p > 0 else 0, None, mask, y, y_b, mean, rstd)
        return y, y_b, (branch, mean, rstd, mask)

    def _add_ln_bwd(self, dy, saved, pre, p):
        f = self.flat
        s, mean, rstd, mask = saved
        n, Dm = dy.shape
        d_res, d_br = self._new(n, Dm), self._new(n, Dm, dtype=torch.bfloat16)
        ops.add_layernorm_bwd(dy, s, mean, rstd, f.p(pre + "weight"), p, mask, d_res, None, d_br, f.g(pre + "weight"),
                              f.g(pre + "bias"))
        return d_res, d_br

    # ------------------------------------------------------------------ feed-forward
    def _ffn_fwd(self, x_b, pre, p):
        f = self.flat
        n = x_b.shape[0]
        w1 = f.s(pre + "linear1.weight")
        h_b = self._lin(x_b, w1, f.p(pre + "linear1.bias"), self._new(n, w1.shape[0], dtype=torch.bfloat16), "ffn1",
                        epilogue=ops.EPI_RELU)
        if p > 0:
            ops.dropout_bf16(h_b, p, self.seed, self._drop_off(h_b.numel()), h_b, None)
        out = self._lin(h_b, f.s(pre + "linear2.weight"), f.p(pre + "linear2.bias"), self._new(n, x_b.shape[1]), "ffn2")
        return out, h_b

    def _ffn_bwd(self, d_f_b, x_b, h_b, pre, p, dx):
        f = self.flat
        n = x_b.shape[0]
        dh = self._new(n, h_b.shape[1])
        self._lin_bwd(d_f_b, h_b, f.s(pre + "linear2.weight"), f.g(pre + "linear2.weight"), f.g(pre + "linear2.bias"), dh, False,
                      "ffn2")
        dpre = self._new(n, h_b.shape[1], dtype=torch.bfloat16)
        ops.relu_bwd(dh, h_b, 1.0 / (1.0 - p) if p > 0 else 1.0, dpre)
        self._lin_bwd(dpre, x_b, f.s(pre + "linear1.weight"), f.g(pre + "linear1.weight"), f.g(pre + "linear1.bias"), dx, True,
                      "ffn1")

    # ------------------------------------------------------------------ forward + backward
    def forward_backward(self, triples, seq, lay: TLayout, eps, beta, n_tok_global=None, batch_global=None,
                         train=True, stats_out=None):
        """One ELBO forward+backward.  `lay` (layout.pack_tlayout, moved to the device) carries the PAD-free index
        arrays; `triples` / `seq` are only used for the batch size.  Returns a device tensor [ce, kl]."""
        f, dev, d, D, dz, V, ldv, H = self.flat, self.device, self.d, self.D, self.dz, self.V, self.ldv, self.H
        bf = torch.bfloat16
        B = lay.enc.n_graphs
        N, Ne = lay.n_tok, lay.n_triples
        n_tok_g = float(N if n_tok_global is None else n_tok_global)
        b_g = int(B if batch_global is None else batch_global)
        out = torch.zeros(2, device=dev) if stats_out is None else stats_out
        p = self.p_drop if train else 0.0
        new = self._new
        S = new(max(lay.enc.sq_total, lay.dec.sq_total) * H)          # score / dP scratch, reused by every layer

        # ---------------- encoder (models.py:78-95)
        x, x_b = new(Ne, D), new(Ne, D, dtype=bf)
        ops.triple_embed_fwd(lay.idx_dev, f.p("enc.e_emb.weight"), f.p("enc.r_emb.weight"), x, x_b)
        enc_saved = []
        for l in range(self.nl_e):
            pre = f"enc.txf.layers.{l}."
            a, sa = self._self_attn_fwd(x_b, pre + "self_attn.", lay.enc, D, False, S, p)
            x1, x1_b, ln1 = self._add_ln_fwd(a, x, pre + "norm1.", p)
            ff, h_b = self._ffn_fwd(x1_b, pre, p)
            x2, x2_b, ln2 = self._add_ln_fwd(ff, x1, pre + "norm2.", p)
            enc_saved.append((x_b, sa, ln1, x1_b, h_b, ln2))
            x, x_b = x2, x2_b
        pooled_b = new(B, D, dtype=bf)
        ops.seg_reduce(x, lay.enc, True, None, 0, None, pooled_b)
        w_heads = f.fused(f.shadow, "enc.mu.weight", "enc.logv.weight", (2 * dz, D))
        b_heads = f.fused(f.param, "enc.mu.bias", "enc.logv.bias", (2 * dz,))
        heads = self._lin(pooled_b, w_heads, b_heads, new(B, 2 * dz), "enc_heads")
        z, z_b = new(B, dz), new(B, dz, dtype=bf)
        ops.reparam_kl_fwd(heads, eps, None, dz, False, 1.0 / (b_g * dz), z, z_b, out[1:2])     # no clamp: models.py:93

        # ---------------- decoder (models.py:108-114)
        mem = self._lin(z_b, f.s("dec.z_proj.weight"), f.p("dec.z_proj.bias"), new(B, d), "z_proj")
        mem_b = new(B, d, dtype=bf)
        ops.cast_bf16(mem, mem_b)
        y, y_b = new(N, d), new(N, d, dtype=bf)
        ops.embed_sum_fwd(f.p("dec.tok_emb.weight"), f.p("dec.pos_emb.weight"), lay.tok_dev, lay.pos_dev, y, y_b)
        dec_saved = []
        for l in range(self.nl_d):
            pre = f"dec.txf.layers.{l}."
            a, sa = self._self_attn_fwd(y_b, pre + "self_attn.", lay.dec, d, True, S, p)
            y1, y1_b, ln1 = self._add_ln_fwd(a, y, pre + "norm1.", p)
            # collapsed cross-attention: uniform weights over L identical memory rows
            wv = f.s(pre + "multihead_attn.in_proj_weight")[2 * d:3 * d]
            bv = f.p(pre + "multihead_attn.in_proj_bias")[2 * d:3 * d]
            vmem = self._lin(mem_b, wv, bv, new(B, d), "xattn_v")
            wts = None
            if p > 0:
                wts = new(N, H)
                ops.xattn_weights(N * H, lay.L_pad, p, self.seed, self.philox_offset, None, wts)
                self.philox_offset += N * H * ((lay.L_pad + 3) // 4)
            xa_b = new(N, d, dtype=bf)
            ops.seg_broadcast(vmem, lay.dec, False, wts, H, None, xa_b)
            c = self._lin(xa_b, f.s(pre + "multihead_attn.out_proj.weight"), f.p(pre + "multihead_attn.out_proj.bias"),
                          new(N, d), "xattn_out")
            y2, y2_b, ln2 = self._add_ln_fwd(c, y1, pre + "norm2.", p)
            ff, h_b = self._ffn_fwd(y2_b, pre, p)
            y3, y3_b, ln3 = self._add_ln_fwd(ff, y2, pre + "norm3.", p)
            dec_saved.append((y_b, sa, ln1, y1_b, wts, xa_b, ln2, y2_b, h_b, ln3))
            y, y_b = y3, y3_b
        logits = new(N, ldv, dtype=bf)
        self._gemm(y_b, K, f.s("dec.out.weight"), K, logits, N, V, d, tag="vocab_fwd", bias=f.p("dec.out.bias"))
        with self._timed("softmax_ce", nbytes=2.0 * N * V * 2 + 12.0 * N):
            ops.softmax_ce(logits, V, lay.tgt_dev, 1.0 / n_tok_g, True, out[0:1], None)
        if not train:
            return out

        # ---------------- decoder backward
        self._gemm(logits, MN, y_b, MN, f.g("dec.out.weight"), V, d, N, tag="vocab_dW")
        ops.colsum(logits, N, V, f.g("dec.out.bias"))
        dy = new(N, d)
        self._gemm(logits, K, f.s("dec.out.weight"), MN, dy, N, d, V, tag="vocab_dY")
        del logits
        self._grad_ready("dec.out.bias", "dec.out.weight")
        dmem = None
        for l in range(self.nl_d - 1, -1, -1):
            pre = f"dec.txf.layers.{l}."
            y_in_b, sa, ln1, y1_b, wts, xa_b, ln2, y2_b, h_b, ln3 = dec_saved[l]
            d_y2, d_ff_b = self._add_ln_bwd(dy, ln3, pre + "norm3.", p)
            self._ffn_bwd(d_ff_b, y2_b, h_b, pre, p, d_y2)
            d_y1, d_c_b = self._add_ln_bwd(d_y2, ln2, pre + "norm2.", p)
            # cross-attention: out_proj, segment-sum back to the graph's memory row, v-projection
            dxa = new(N, d)
            self._lin_bwd(d_c_b, xa_b, f.s(pre + "multihead_attn.out_proj.weight"), f.g(pre + "multihead_attn.out_proj.weight"),
                          f.g(pre + "multihead_attn.out_proj.bias"), dxa, False, "xattn_out")
            dvm, dvm_b = new(B, d), new(B, d, dtype=bf)
            ops.seg_reduce(dxa, lay.dec, False, wts, H, dvm, dvm_b)
            g_in, g_inb = f.g(pre + "multihead_attn.in_proj_weight"), f.g(pre + "multihead_attn.in_proj_bias")
            g_in[:2 * d].zero_()          # q / k projections of a uniform softmax: exactly zero gradient
            g_inb[:2 * d].zero_()
            if dmem is None:
                dmem = new(B, d)
                acc = False
            else:
                acc = True
            self._lin_bwd(dvm_b, mem_b, f.s(pre + "multihead_attn.in_proj_weight")[2 * d:3 * d], g_in[2 * d:3 * d],
                          g_inb[2 * d:3 * d], dmem, acc, "xattn_v", dy_f32=dvm)
            d_y, d_a_b = self._add_ln_bwd(d_y1, ln1, pre + "norm1.", p)
            self._self_attn_bwd(d_a_b, y_in_b, sa, pre + "self_attn.", lay.dec, d, True, S, p, d_y)
            self._grad_ready(pre + "norm3.weight", pre + "self_attn.in_proj_bias")
            dy = d_y
        g_tok, g_pos = f.g("dec.tok_emb.weight"), f.g("dec.pos_emb.weight")
        g_tok.zero_()
        g_pos.zero_()
        ops.tok_scatter_add(dy, lay.tok_dev, g_tok)
        ops.tok_scatter_add(dy, lay.pos_dev, g_pos)
        self._grad_ready("dec.tok_emb.weight", "dec.pos_emb.weight")

        # ---------------- z_proj, reparameterisation + KL, heads
        dmem_b = new(B, d, dtype=bf)
        ops.cast_bf16(dmem, dmem_b)
        dz_in = new(B, dz)
        self._lin_bwd(dmem_b, z_b, f.s("dec.z_proj.weight"), f.g("dec.z_proj.weight"), f.g("dec.z_proj.bias"), dz_in, False,
                      "z_proj", dy_f32=dmem)
        ld_dh = _up8(2 * dz)
        dheads = torch.zeros(B, ld_dh, device=dev)
        dheads_b = torch.zeros(B, ld_dh, device=dev, dtype=bf)
        ops.reparam_kl_bwd(heads, eps, None, dz_in, dz, False, beta / (b_g * dz), dheads, dheads_b)
        g_wh = f.fused(f.grad, "enc.mu.weight", "enc.logv.weight", (2 * dz, D))
        g_bh = f.fused(f.grad, "enc.mu.bias", "enc.logv.bias", (2 * dz,))
        self._gemm(dheads_b[:, :2 * dz], MN, pooled_b, MN, g_wh, 2 * dz, D, B, tag="enc_heads_bwd")
        ops.colsum(dheads, B, 2 * dz, g_bh)
        dpool = new(B, D)
        self._gemm(dheads_b[:, :2 * dz], K, w_heads, MN, dpool, B, D, 2 * dz, tag="enc_heads_bwd")
        self._grad_ready("dec.z_proj.weight", "enc.logv.bias")

        # ---------------- encoder backward
        dx = new(Ne, D)
        ops.seg_broadcast(dpool, lay.enc, True, None, 0, dx, None)
        for l in range(self.nl_e - 1, -1, -1):
            pre = f"enc.txf.layers.{l}."
            x_in_b, sa, ln1, x1_b, h_b, ln2 = enc_saved[l]
            d_x1, d_ff_b = self._add_ln_bwd(dx, ln2, pre + "norm2.", p)
            self._ffn_bwd(d_ff_b, x1_b, h_b, pre, p, d_x1)
            d_x, d_a_b = self._add_ln_bwd(d_x1, ln1, pre + "norm1.", p)
            self._self_attn_bwd(d_a_b, x_in_b, sa, pre + "self_attn.", lay.enc, D, False, S, p, d_x)
            self._grad_ready(pre + "norm2.weight", pre + "self_attn.in_proj_bias")
            dx = d_x
        gE, gR = f.g("enc.e_emb.weight"), f.g("enc.r_emb.weight")
        gR.zero_()
        gE.zero_()
        ops.triple_embed_bwd(lay.idx_dev, dx, gE, gR)
        self._grad_ready("enc.r_emb.weight", "enc.e_emb.weight")
        return out

    # ------------------------------------------------------------------ fp32 inference (generation / validation)
    def _lin32(self, x, wname, bname, epilogue=ops.EPI_NONE):
        """y = epi(x W^T + b) in fp32 on the FMA kernel, fp32 master weights (integer outputs of generation must
        match the reference, so no bf16 here)."""
        f = self.flat
        w = wname if torch.is_tensor(wname) else f.p(wname)
        b = bname if torch.is_tensor(bname) else f.p(bname)
        M, Kd = x.shape
        y = self._new(M, w.shape[0])
        ops.gemm(x, K, w, K, y, M, w.shape[0], Kd, bias=b, epilogue=epilogue, backend="simt")
        return y

    def _layer32(self, x, pre, seg, Dm, causal, S, cross=None):
        """One post-LN layer in fp32: self-attention, [collapsed cross-attention row `cross` per graph], ReLU FFN."""
        f, H = self.flat, self.H
        hd, n = Dm // H, x.shape[0]
        zeros = lambda: (self._new(n), self._new(n))  # noqa: E731
        qkv = self._lin32(x, pre + "self_attn.in_proj_weight", pre + "self_attn.in_proj_bias")
        ops.attn_bgemm(qkv, TOK, False, 0, qkv, TOK, True, Dm, S, SQ, 0, seg, H, hd, 0, causal, 1.0 / math.sqrt(hd))
        ops.attn_softmax_inplace(S, seg, H, causal)
        o = self._new(n, Dm)
        ops.attn_bgemm(S, SQ, False, 0, qkv, TOK, False, 2 * Dm, o, TOK, 0, seg, H, hd, 1, causal, 1.0)
        a = self._lin32(o, pre + "self_attn.out_proj.weight", pre + "self_attn.out_proj.bias")
        y = self._new(n, Dm)
        ops.add_layernorm_fwd(a, x, f.p(pre + "norm1.weight"), f.p(pre + "norm1.bias"), self.ln_eps, 0.0, 0, 0, None, None, y,
                              None, *zeros())
        x, k_ff = y, 2
        if cross is not None:
            c = self._new(n, Dm)
            ops.seg_broadcast(cross, seg, False, None, 0, c, None)
            y = self._new(n, Dm)
            ops.add_layernorm_fwd(c, x, f.p(pre + "norm2.weight"), f.p(pre + "norm2.bias"), self.ln_eps, 0.0, 0, 0, None, None,
                                  y, None, *zeros())
            x, k_ff = y, 3
        h = self._lin32(x, pre + "linear1.weight", pre + "linear1.bias", ops.EPI_RELU)
        ff = self._lin32(h, pre + "linear2.weight", pre + "linear2.bias")
        y = self._new(n, Dm)
        ops.add_layernorm_fwd(ff, x, f.p(pre + f"norm{k_ff}.weight"), f.p(pre + f"norm{k_ff}.bias"), self.ln_eps, 0.0, 0, 0,
                              None, None, y, None, *zeros())
        return y

    @torch.no_grad()
    def encode_stats(self, triples):
        """(mu, logv) of AutoRegEncoder.forward in eval mode (models.py:78-93), fp32."""
        from .layout import segments_from_lens
        import numpy as np
        tri = triples.detach().cpu().numpy()
        live = (tri[:, :, 1] != self.pad_rid) if self.pad_rid is not None else np.ones(tri.shape[:2], dtype=bool)
        seg = segments_from_lens(live.sum(1)).to(self.device)
        idx = torch.from_numpy(np.ascontiguousarray(tri[live].astype(np.int32))).to(self.device)
        f, D = self.flat, self.D
        x, xb = self._new(seg.n_rows, D), self._new(seg.n_rows, D, dtype=torch.bfloat16)
        ops.triple_embed_fwd(idx, f.p("enc.e_emb.weight"), f.p("enc.r_emb.weight"), x, xb)
        S = self._new(max(seg.sq_total, 1) * self.H)
        for l in range(self.nl_e):
            x = self._layer32(x, f"enc.txf.layers.{l}.", seg, D, False, S)
        pooled = self._new(seg.n_graphs, D)
        ops.seg_reduce(x, seg, True, None, 0, pooled, None)
        return self._lin32(pooled, "enc.mu.weight", "enc.mu.bias"), self._lin32(pooled, "enc.logv.weight", "enc.logv.bias")

    @torch.no_grad()
    def decode_logits(self, z, tgt):
        """logits [B, L', V] of AutoRegDecoder.forward (models.py:108-114) / DecoderOnlyTransformer.forward (:360-365)
        in eval mode for any prefix length L', fp32.  z is ignored by the decoder-only model."""
        from .layout import segments_from_lens
        import numpy as np
        B, Lp = tgt.shape
        f, d, dev = self.flat, self.d, self.device
        seg = segments_from_lens(np.full(B, Lp, dtype=np.int32)).to(dev)
        tok = tgt.reshape(-1).to(torch.int32).contiguous()
        pos = torch.arange(Lp, device=dev, dtype=torch.int32).repeat(B).contiguous()
        y, yb = self._new(B * Lp, d), self._new(B * Lp, d, dtype=torch.bfloat16)
        ops.embed_sum_fwd(f.p("dec.tok_emb.weight"), f.p("dec.pos_emb.weight"), tok, pos, y, yb)
        mem = self._lin32(z.to(torch.float32).contiguous(), "dec.z_proj.weight", "dec.z_proj.bias") if self.has_enc else None
        S = self._new(seg.sq_total * self.H)
        for l in range(self.nl_d):
            pre = f"dec.txf.layers.{l}."
            cross = None
            if self.has_enc:   # uniform attention over L identical memory rows == out_proj(v_proj(mem))
                w_in, b_in = f.p(pre + "multihead_attn.in_proj_weight"), f.p(pre + "multihead_attn.in_proj_bias")
                vm = self._lin32(mem, w_in[2 * d:3 * d], b_in[2 * d:3 * d])
                cross = self._lin32(vm, pre + "multihead_attn.out_proj.weight", pre + "multihead_attn.out_proj.bias")
            y = self._layer32(y, pre, seg, d, True, S, cross)
        w_out = f.p("dec.tok_emb.weight") if self.tied else f.p("dec.out.weight")
        return self._lin32(y, w_out, f.p("dec.out.bias")).view(B, Lp, -1)

    def train_step_graphed(self, triples, seq, lay, eps, beta, lr=None, n_tok_global=None, batch_global=None):
        """The ragged index arrays of a t-SAIL batch change every step: no graph replay yet, eager launches."""
        return self.train_step(triples, seq, lay, eps, beta, lr, n_tok_global, batch_global)


# ============================================================================================================
# t-ARK: decoder-only Transformer (reference models.py:349-366 DecoderOnlyTransformer inside ARK, :368-405)
# ============================================================================================================
def tark_param_order(model):
    named = dict(model.named_parameters())       # the tied dec.out.weight is deduplicated by torch
    groups = []
    add = lambda *names: groups.append([(n, named[n]) for n in names])  # noqa: E731
    add("dec.out.bias")
    if "dec.out.weight" in named:
        add("dec.out.weight")
    for l in range(len(model.dec.txf.layers) - 1, -1, -1):
        p = f"dec.txf.layers.{l}."
        for n in ("norm2.weight", "norm2.bias", "linear2.weight", "linear2.bias", "linear1.weight", "linear1.bias",
                  "norm1.weight", "norm1.bias", "self_attn.out_proj.weight", "self_attn.out_proj.bias",
                  "self_attn.in_proj_weight", "self_attn.in_proj_bias"):
            add(p + n)
    add("dec.tok_emb.weight")
    add("dec.pos_emb.weight")
    missing = set(named) - {n for g in groups for n, _ in g}
    if missing:
        raise RuntimeError(f"parameters without a slot in the flat layout: {sorted(missing)}")
    return groups


class TArkEngine(TSailEngine):
    """CE-only step (reference train.py:42-58) of the decoder-only Transformer: tok_emb + pos_emb, n_layers post-LN
    nn.TransformerEncoderLayer blocks under a causal mask, (tied) vocabulary projection."""
    MODEL_TYPE = "t-ARK"

    @staticmethod
    def _dropout_from_config(cfg):
        return float(cfg.get("dec_dropout", 0.1))        # models.py:353: dropout=dropout

    @staticmethod
    def _param_order(model):
        return tark_param_order(model)

    def forward_backward(self, triples, seq, lay: TLayout, eps, beta, n_tok_global=None, batch_global=None,
                         train=True, stats_out=None):
        f, dev, d, V, ldv, H = self.flat, self.device, self.d, self.V, self.ldv, self.H
        bf = torch.bfloat16
        N = lay.n_tok
        n_tok_g = float(N if n_tok_global is None else n_tok_global)
        out = torch.zeros(2, device=dev) if stats_out is None else stats_out
        p = self.p_drop if train else 0.0
        new = self._new
        S = new(lay.dec.sq_total * H)
        y, y_b = new(N, d), new(N, d, dtype=bf)
        ops.embed_sum_fwd(f.p("dec.tok_emb.weight"), f.p("dec.pos_emb.weight"), lay.tok_dev, lay.pos_dev, y, y_b)
        saved = []
        for l in range(self.nl_d):
            pre = f"dec.txf.layers.{l}."
            a, sa = self._self_attn_fwd(y_b, pre + "self_attn.", lay.dec, d, True, S, p)
            y1, y1_b, ln1 = self._add_ln_fwd(a, y, pre + "norm1.", p)
            ff, h_b = self._ffn_fwd(y1_b, pre, p)
            y2, y2_b, ln2 = self._add_ln_fwd(ff, y1, pre + "norm2.", p)
            saved.append((y_b, sa, ln1, y1_b, h_b, ln2))
            y, y_b = y2, y2_b
        logits = new(N, ldv, dtype=bf)
        w_out = f.s("dec.tok_emb.weight") if self.tied else f.s("dec.out.weight")
        self._gemm(y_b, K, w_out, K, logits, N, V, d, tag="vocab_fwd", bias=f.p("dec.out.bias"))
        with self._timed("softmax_ce", nbytes=2.0 * N * V * 2 + 12.0 * N):
            ops.softmax_ce(logits, V, lay.tgt_dev, 1.0 / n_tok_g, True, out[0:1], None)
        if not train:
            return out
        g_wout = f.g("dec.tok_emb.weight") if self.tied else f.g("dec.out.weight")
        self._gemm(logits, MN, y_b, MN, g_wout, V, d, N, tag="vocab_dW")
        ops.colsum(logits, N, V, f.g("dec.out.bias"))
        dy = new(N, d)
        self._gemm(logits, K, w_out, MN, dy, N, d, V, tag="vocab_dY")
        del logits
        self._grad_ready("dec.out.bias", "dec.out.bias" if self.tied else "dec.out.weight")
        for l in range(self.nl_d - 1, -1, -1):
            pre = f"dec.txf.layers.{l}."
            y_in_b, sa, ln1, y1_b, h_b, ln2 = saved[l]
            d_y1, d_ff_b = self._add_ln_bwd(dy, ln2, pre + "norm2.", p)
            self._ffn_bwd(d_ff_b, y1_b, h_b, pre, p, d_y1)
            d_y, d_a_b = self._add_ln_bwd(d_y1, ln1, pre + "norm1.", p)
            self._self_attn_bwd(d_a_b, y_in_b, sa, pre + "self_attn.", lay.dec, d, True, S, p, d_y)
            self._grad_ready(pre + "norm2.weight", pre + "self_attn.in_proj_bias")
            dy = d_y
        g_tok, g_pos = f.g("dec.tok_emb.weight"), f.g("dec.pos_emb.weight")
        if not self.tied:
            g_tok.zero_()
        g_pos.zero_()
        ops.tok_scatter_add(dy, lay.tok_dev, g_tok)
        ops.tok_scatter_add(dy, lay.pos_dev, g_pos)
        self._grad_ready("dec.tok_emb.weight", "dec.pos_emb.weight")
        return out
