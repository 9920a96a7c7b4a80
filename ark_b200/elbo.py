"""The fused ELBO training step of the KG-VAE (SAIL): hand-scheduled forward + backward + Adam.

This is the B200-native replacement of the body of the reference's training loop
(kgvae/experiments/ablation_study.py:59-76):

    logits, mu, logv = model(triples, seq[:, :-1]); ce = F.cross_entropy(..., ignore_index=PAD)
    kl = model.kl_mean(mu, logv); loss = ce + b*kl; loss.backward(); optimizer.step()

Instead of an autograd graph over ~60 ATen ops, the step is a fixed schedule of C-ABI kernel calls
(include/arkb200.h) over packed, PAD-free token rows (ark_b200/layout.py).  The [B, L, V] logits tensor is
produced once in bf16, turned into its own gradient in place by the fused softmax-CE kernel and consumed by
the two backward GEMMs — the probability matrix never exists.  torch provides memory, streams and NCCL only.

Numerics: GEMM operands bf16 (weights from the flat bf16 shadow, activations written in bf16 by the
producing kernel), fp32 accumulation, fp32 master weights / recurrent state / gates / mu / logv / KL / loss.
"""
from __future__ import annotations

import math
import os
import sys

import numpy as np
import torch

from . import ops
from .flat import FlatParams, ark_param_order, merge_span, sail_param_order
from .layout import PackedLayout, pack_layout

K, MN = ops.MAJOR_K, ops.MAJOR_MN


def _up8(n):
    return (n + 7) // 8 * 8


class SailEngine:
    """Owns the flat parameter storage of one SAIL module and runs its ELBO step on one GPU.

    ``gemm_backend``: "tc" (tcgen05/TMA, default, the product path) or "simt" (fp32-FMA kernel with the same
    bf16 operands — a debugging cross-check, never chosen automatically for TMA-eligible shapes).
    """
    # defaults of the data-parallel / GRU-driver switches, also for subclasses with their own __init__
    # (ark_b200.tsail: the Transformer engines reuse the bucket machinery below)
    _hold_comm, _held = False, ()
    mm, _mm_spans, dp_mm_ctas, _factor_wss = None, frozenset(), 32, {}     # multicast gradient exchange (csrc/dp_reduce.cu): SailEngine only
    dp_hold_comm, dp_factor_gather, dp_emb_min_bytes, dp_mlp_pipe = False, True, None, True
    _gru_cluster_ws = None
    _leaf_used, use_leaf_stream, leaf_stream, leaf_embedding, capture_nccl, post_stream = False, False, None, False, False, None     # (the Transformer engines do not fork leaf work)
    prof_stream = {}                 # tag -> "main" | "leaf" | "side": the stream a timed op was launched on (bench.py)
    nvtx = os.environ.get("ARK_NVTX", "0") != "0"   # NVTX range per kernel family (same tags as bench.py's `kernels` list)
    logits_chunk_rows = 16384        # packed rows per logits workspace chunk (see forward_backward)
    max_graphs = 8                   # captured step graphs kept per engine (oldest evicted first)
    keep = None                      # tests: a dict that receives references to the GRU stack's internal tensors

    def __init__(self, model, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, gemm_backend="tc", dist_group=None,
                 bucket_mb=16.0, seed=0):
        cfg = model.config
        if cfg["model_type"] not in ("SAIL", "ARK"):
            raise NotImplementedError("SailEngine accelerates model_type 'SAIL' (MLP encoder + GRU decoder) and its "
                                      "decoder-only sibling 'ARK'")
        if str(cfg.get("precision", "bf16")).lower() != "bf16":
            raise NotImplementedError(f"precision {cfg.get('precision')!r}: the fused step is built for bf16 GEMM operands with "
                                      "fp32 accumulation / masters / state only (a tcgen05 kind::tf32 path does not exist)")
        self.has_enc = cfg["model_type"] == "SAIL"
        dev = next(model.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("SailEngine needs the model on a CUDA device: there is no CPU path")
        self.model, self.cfg, self.device = model, cfg, dev
        self.d, self.dz, self.V = cfg["d_model"], cfg["d_latent"], cfg["vocab_size"]
        self.nl = model.dec.gru.num_layers
        self.n_mlp = len([m for m in model.enc.mlp if isinstance(m, torch.nn.Linear)]) if self.has_enc else 0
        self.pad_rid, self.pad_eid = cfg.get("pad_rid"), cfg.get("pad_eid")
        self.tied = model.dec.out.weight is model.dec.tok_emb.weight
        self.p_drop = float(cfg.get("dec_dropout", 0.1)) if self.nl > 1 else 0.0
        if self.d % 8:
            raise ValueError("d_model must be a multiple of 8")
        self.group = dist_group
        self.world = torch.distributed.get_world_size(dist_group) if dist_group is not None else 1
        # data parallel on an NVSwitch node: the flat buffers live in symmetric multicast memory and the gradient
        # buckets are reduced IN THE SWITCH by a kernel that also applies Adam to the slice this rank owns and
        # multicasts the new parameters (csrc/dp_reduce.cu).  ARK_DP_MULTIMEM=0: NCCL all-reduce + replicated Adam;
        # =1: fail instead of falling back to NCCL when the allocation has no multicast mapping
        self.mm, self._mm_spans, self._factor_wss = None, set(), {}
        mm_mode = os.environ.get("ARK_DP_MULTIMEM", "auto")
        if (self.world >= 4 and mm_mode != "0") or (self.world > 1 and mm_mode == "1"):
            try:
                from .symm import SymmFlat
                self.mm = SymmFlat.__new__(SymmFlat)
                def alloc(numel, mm=self.mm):
                    mm.__init__(numel, dev, dist_group)
                    return mm.buffers()
                self.flat = FlatParams(sail_param_order(model) if self.has_enc else ark_param_order(model), dev, alloc)
            except Exception as exc:      # no multicast (PCIe box, MIG, old driver): the NCCL path is the same math
                if mm_mode == "1":
                    raise
                print(f"[ark_b200] multicast gradient exchange unavailable ({type(exc).__name__}: {exc}); using NCCL",
                      file=sys.stderr, flush=True)
                self.mm = None
        if self.mm is None:
            self.flat = FlatParams(sail_param_order(model) if self.has_enc else ark_param_order(model), dev)
        # CTAs of the exchange kernels: at 8 ranks 16 already saturate the switch (profiles/r02_dp_multicast.md) and
        # 128 + 16 <= 148 lets them run NEXT TO the one-CTA-per-SM GRU kernels
        self.dp_mm_ctas = int(os.environ.get("ARK_DP_MM_CTAS", "16" if self.world >= 8 else "32"))
        self.lr, self.betas, self.eps = float(lr), betas, float(eps)
        self.step_count = 0
        self.backend = gemm_backend
        self.seed, self.philox_offset = int(seed), 0
        self.ldv = _up8(self.V)
        if self.world > 1 and bucket_mb == 16.0:
            bucket_mb = 48.0        # with NVSwitch a few large all-reduces beat many launch-latency-bound small ones
        bucket_mb = float(os.environ.get("ARK_BUCKET_MB", bucket_mb))
        self.bucket_elems = int(bucket_mb * (1 << 20) / 4)
        # gradient all-reduce + per-bucket Adam, overlapping backward.  High priority: its (few) NCCL CTAs must not queue
        # behind the full grids of the main stream's GEMMs
        self.comm_stream = torch.cuda.Stream(device=dev, priority=-1 if os.environ.get("ARK_COMM_PRIORITY", "1") != "0" else 0)
        # weight-gradient GEMMs / bias column sums that nothing in the backward pass consumes ("leaves") run on their own
        # stream, concurrently with the dependent chain (dY -> GRU backward -> dX -> scatter -> encoder backward): they fill
        # the SMs the latency-bound chain kernels leave idle (and each other's partial last waves)
        self.leaf_stream = torch.cuda.Stream(device=dev)
        # data parallel: the side stream carries ONLY the NCCL calls, back to back; what follows a collective (Adam on the
        # reduced slice, the global dW GEMM on gathered factors) runs on a third stream so that the HBM-bound update of
        # bucket i overlaps the NVLink-bound all-reduce of bucket i+1
        self.post_stream = torch.cuda.Stream(device=dev) if self.world > 1 else self.comm_stream
        self.use_leaf_stream = os.environ.get("ARK_LEAF_STREAM", "1") != "0"
        self.leaf_embedding = os.environ.get("ARK_LEAF_EMB", "0") != "0"
        # data parallel + CUDA graphs: capture the NCCL collectives INTO the step graph (one replay per step) instead of
        # cutting the graph into segments with eager collectives in between
        # (ARK_CAPTURE_NCCL=0: the round-1 scheme — a chain of graph segments with eager collectives between them)
        self.capture_nccl = os.environ.get("ARK_CAPTURE_NCCL", "1") != "0"   # also fork layer-0 dX + the embedding scatter
        self._leaf_used = False
        self._upd = None                 # (mode, lr) while a train step is in flight: buckets are updated as they finish
        self._pending = []
        self._hold_comm, self._held = False, []   # see _comm_action: no NCCL next to the 128-CTA cooperative GRU kernels
        self.prof, self.prof_stream = None, {}
        self._capturing = False
        self._segment_break = None
        self._graphs = {}
        # per-step scalars of a replayed graph, ONE 24-byte device buffer (one H2D copy per step):
        self._dyn_raw = torch.zeros(24, device=dev, dtype=torch.uint8)
        self.dyn_f = self._dyn_raw[:8].view(torch.float32)           # [lr/(1-b1^t), 1/sqrt(1-b2^t)]
        self.dyn_i = self._dyn_raw[8:16].view(torch.int64)           # running Philox offset
        self.dyn_beta = self._dyn_raw[16:20].view(torch.float32)     # beta of the ELBO (changes every epoch: NOT a graph key)
        self.launches_replayed = 0       # kernels of libarkb200 executed through graph replays
        self.force_unfused_gru = False   # tests: compare the persistent GRU kernel with the per-step path
        self.gru_mode = os.environ.get("ARK_GRU_MODE", "auto")   # "auto": cluster stack kernel (long chains of short batch tiles), else the
                                         # wavefront stack kernel when it fits, else per-layer persistent;
                                         # "wave": never the cluster kernel; "layer": neither (tests / A-B timing)
        self._gru_cluster_ws = None      # scratch of the cluster GRU kernels (gi^T / dx^T slices), grown on demand
        self.dp_factor_gather = os.environ.get("ARK_DP_FACTOR_GATHER", "1") != "0"   # data parallel: all-gather the [B, 3d] FACTORS of the encoder-MLP weight
                                         # gradients (dW = dY^T X, rank <= global batch) instead of all-reducing dW
        self.dp_hold_comm = os.environ.get("ARK_DP_HOLD_COMM", "0") != "0"   # opt-in: see _release_comm
        self.dp_mlp_pipe = os.environ.get("ARK_DP_MLP_PIPE", "1") != "0"     # encoder-MLP factor exchange, layer by layer
        self.dp_emb_min_bytes = None     # ... and (opt-in: a byte threshold) of the entity-embedding gradient of a large
                                         # table.  Off by default: the scatter's atomics make the ranks' results differ
                                         # in the last bits, so replicas would drift apart without a periodic re-sync
        self.stats = torch.zeros(4, device=dev)  # [ce, kl, steps, unused] accumulated on device
        self.refresh_shadow()
        if hasattr(model, "_attach_engine"):
            model._attach_engine(self)

    # ------------------------------------------------------------------ parameter plumbing
    def refresh_shadow(self):
        """bf16 operand copies <- fp32 masters (after init / load_state_dict; Adam keeps them in sync)."""
        ops.cast_bf16(self.flat.param, self.flat.shadow)

    def _w(self, name):       # bf16 shadow view
        return self.flat.s(name)

    def _gemm(self, A, am, B, bm, C, M, N, Kd, tag="gemm", **kw):
        backend = "tc" if (self.backend == "tc" and ops.tc_eligible(A, B)) else "simt"
        with self._timed(f"gemm_{backend}:{tag}", flops=2.0 * M * N * Kd):
            ops.gemm(A, am, B, bm, C, M, N, Kd, backend=backend, **kw)

    # ------------------------------------------------------------------ GRU driver choice
    def _use_gru_cluster(self, d, b0, nl, L):
        """Batch-tile rows of the cluster kernel (0 = do not use it).  It wins where the chain is long and the batch
        tile short (its DSMEM exchange costs ~0.8 us per step; the wavefront kernel's global-memory exchange ~2.3 us)."""
        if self.gru_mode != "cluster" and L < 32:
            return 0
        return ops.gru_cluster_supported(d, b0, nl, L)

    def _cluster_ws(self, L, b0, d, nl):
        need = ops.gru_cluster_workspace_bytes(L, b0, d, nl)
        if self._gru_cluster_ws is None or self._gru_cluster_ws.numel() < need:
            if self._capturing:
                raise RuntimeError("cluster GRU scratch must be allocated before graph capture")
            self._gru_cluster_ws = torch.empty(need, device=self.device, dtype=torch.uint8)
        return self._gru_cluster_ws

    # ------------------------------------------------------------------ optional per-op device timing
    class _Timer:
        def __init__(self, eng, tag, flops, nbytes):
            self.eng, self.tag, self.flops, self.nbytes = eng, tag, flops, nbytes

        def __enter__(self):
            if SailEngine.nvtx:
                torch.cuda.nvtx.range_push(self.tag)
            if self.eng.prof is not None:
                # inside a stream capture the events become EXTERNAL event-record nodes of the graph: every replay
                # re-records them, so elapsed_time() after a replay is the kernel's time inside the replayed graph
                # (no host launch gaps — what bench.py's per-kernel list is built from)
                ext = bool(self.eng._capturing)
                self.e0 = torch.cuda.Event(enable_timing=True, external=ext)
                self.e1 = torch.cuda.Event(enable_timing=True, external=ext)
                self.e0.record()

        def __exit__(self, *exc):
            if self.eng.prof is not None:
                self.e1.record()
                eng, cur = self.eng, torch.cuda.current_stream()
                where = ("side" if cur in (eng.comm_stream, eng.post_stream) else "leaf" if cur == eng.leaf_stream else "main")
                eng.prof_stream[self.tag] = where
                eng.prof.append((self.tag, self.e0, self.e1, self.flops, self.nbytes))
            if SailEngine.nvtx:
                torch.cuda.nvtx.range_pop()

    def _timed(self, tag, flops=0.0, nbytes=0.0):
        """CUDA events around one op on the launching stream when `self.prof` is a list (bench.py's roofline
        pass); free otherwise."""
        return SailEngine._Timer(self, tag, flops, nbytes)

    def profile_summary(self, prof=None, agg=None):
        """tag -> dict(ms, calls, flops, bytes) from the recorded events (synchronises).  `prof`: the event list of a
        captured profiling graph (train_step_graphed with self.prof set), read after EACH replay; `agg` accumulates."""
        torch.cuda.synchronize()
        agg = {} if agg is None else agg
        for tag, e0, e1, fl, nb in (self.prof if prof is None else prof):
            a = agg.setdefault(tag, {"ms": 0.0, "calls": 0, "flops": 0.0, "bytes": 0.0})
            a["ms"] += e0.elapsed_time(e1)
            a["calls"] += 1
            a["flops"] += fl
            a["bytes"] += nb
        return agg

    # ------------------------------------------------------------------ forward + backward
    def forward_backward(self, triples, seq, lay: PackedLayout, eps, beta, n_tok_global=None, batch_global=None,
                         train=True, stats_out=None):
        """One ELBO forward+backward over a device-resident batch.  Gradients land in ``self.flat.grad``
        (overwritten, never accumulated).  Returns a device tensor [ce, kl] (already globally normalised when
        the *_global normalisers are given; the caller sums them over ranks)."""
        gen = self._fb_gen(triples, seq, lay, eps, beta, n_tok_global, batch_global, train, stats_out)
        try:
            next(gen)
        except StopIteration as stop:
            return stop.value
        raise RuntimeError("forward_backward: the fused step must not suspend")

    def _fb_gen(self, triples, seq, lay: PackedLayout, eps, beta, n_tok_global=None, batch_global=None,
                train=True, stats_out=None, dropout=None, autograd=False):
        """The step as a GENERATOR.  Fused mode (autograd=False) runs to completion on the first next().  With
        autograd=True it suspends once, between forward and backward: it yields {"logits": bf16 [N, ldv] packed
        time-major rows, "heads": f32 [B, 2dz] = (mu | raw logv)} and is resumed with
        {"dmu": f32 [B, dz] | None, "dlogv": ...} AFTER the caller has overwritten `logits` in place with
        d(loss)/d(logits) — the activations of the forward pass stay alive in the suspended frame exactly as long as
        the autograd node that owns the generator (kgvae.model.models._ForwardFn: the differentiable
        SAIL.forward / ARK.forward of the reference interface, models.py:317-320,395-405).
        `dropout` (default: `train`) switches the inter-layer GRU dropout independently of running the backward."""
        dropout = train if dropout is None else bool(dropout)
        f, dev, d, dz, V, ldv, nl = self.flat, self.device, self.d, self.dz, self.V, self.ldv, self.nl
        B = seq.shape[0]
        N, L = lay.n_tok, lay.L
        d3 = 3 * d
        bf, f32 = torch.bfloat16, torch.float32
        n_tok_g = float(N if n_tok_global is None else n_tok_global)
        b_g = int(B if batch_global is None else batch_global)
        out = torch.zeros(2, device=dev) if stats_out is None else stats_out
        self._drop_calls = 0
        use_tc = 1 if self.backend == "tc" else 0
        self._hold_comm, self._held = False, []
        if train and self.world > 1 and use_tc and not self.force_unfused_gru and self.dp_hold_comm:
            b0_ = int(lay.bt[0])
            stack = ((self.gru_mode in ("auto", "cluster") and self._use_gru_cluster(d, b0_, nl, L) > 0)
                     or (self.gru_mode in ("auto", "wave") and ops.gru_wave_supported(d, b0_, nl) > 0))
            self._hold_comm = (not stack) and ops.gru_persist_supported(d, b0_) > 0
        new = lambda *s, dtype=f32: torch.empty(*s, device=dev, dtype=dtype)  # noqa: E731

        # ---------------- decoder inputs: token packing + embedding gather + (per-layer GRU path) the layer-0 input
        # projection do not depend on the encoder, so they run on the leaf stream WHILE the encoder (a chain of
        # latency-bound small-M GEMMs) runs on this one; joined in front of the GRU
        b0 = int(lay.bt[0])
        cl_nb = (self._use_gru_cluster(d, b0, nl, L)
                 if (use_tc and self.gru_mode in ("auto", "cluster") and not self.force_unfused_gru) else 0)
        wave_ok = use_tc and not self.force_unfused_gru and ops.gru_wave_supported(d, b0, nl) > 0
        # A wavefront grid larger than the GPU (d = 512 with two batch tiles: 32 x 2 x 3 = 192 CTAs) runs its tile groups
        # back to back; there the per-layer kernels (half-tile forward, K-split backward) are faster although they take
        # nl * L dependent steps (measured on syn-paths, L = 10: 0.82 vs 0.89 ms per training step; NOT for long chains:
        # wd-articles at 256 graphs per GPU, L = 637: 47.9 vs 42.2 ms)
        if (wave_ok and self.gru_mode == "auto" and (d // 16) * ((b0 + 127) // 128) * nl > 148 and L <= 32
                and ops.gru_persist_bwd_ksplit(d, b0) and b0 > 64):
            wave_ok = False
        cluster = cl_nb > 0
        # the two stack kernels share every tensor of their contract, so the direction can be chosen separately: with
        # 64-row batch tiles the cluster backward (16 work items per epilogue thread) is no faster than the wavefront's
        cluster_bwd = cluster and (cl_nb <= 32 or self.gru_mode == "cluster" or not wave_ok)
        wave = cluster or (self.gru_mode in ("auto", "wave") and wave_ok)
        persist = use_tc and ops.gru_persist_supported(d, b0) > 0 and not self.force_unfused_gru and not wave
        tok, tgt = new(N, dtype=torch.int32), new(N, dtype=torch.int32)
        row_t = None if self.has_enc else new(N, dtype=torch.int32)
        x_b = new(N, d, dtype=bf)
        gi0 = None if wave else new(N, d3)

        def decoder_inputs():
            ops.pack_tokens(seq, lay.perm_dev, lay.bt_dev, lay.off_dev, L, tok, tgt, row_t)
            if self.has_enc:
                ops.tok_gather_fwd(self._w("dec.tok_emb.weight"), tok, None, x_b)
            else:   # token + position embedding (reference models.py:340-342)
                ops.tok_pos_gather_fwd(self._w("dec.tok_emb.weight"), self._w("dec.pos_emb.weight"), tok, row_t, x_b)
            if gi0 is not None:
                self._gemm(x_b, K, self._w("dec.gru.weight_ih_l0"), K, gi0, N, d3, d, tag="gru_gi",
                           bias=f.p("dec.gru.bias_ih_l0"))
        if self.has_enc:
            self._leaf(decoder_inputs)
        else:
            decoder_inputs()

        # ---------------- encoder forward (models.py:46-64)
        if self.has_enc:
            g_b, inv_cnt = new(B, d3, dtype=bf), new(B)
            with self._timed("gather_pool_fwd", nbytes=3.0 * lay.n_triples * d * 4 + 3.0 * lay.n_triples * 8 + B * d3 * 2):
                ops.gather_pool_fwd(triples, lay.perm_dev, f.p("enc.e_emb.weight"), f.p("enc.r_emb.weight"), self.pad_rid,
                                    None, g_b, inv_cnt)
            acts, pres = [g_b], []
            for k in range(self.n_mlp):
                a_next, pre = new(B, d3, dtype=bf), new(B, d3)
                self._gemm(acts[-1], K, self._w(f"enc.mlp.{2 * k}.weight"), K, a_next, B, d3, d3, tag="enc_mlp",
                           bias=f.p(f"enc.mlp.{2 * k}.bias"), epilogue=ops.EPI_GELU, aux=pre)
                acts.append(a_next)
                pres.append(pre)
            # data parallel: dW_k = dY_k^T X_k has rank <= global batch << 3d, so the ranks exchange the FACTORS
            # ([B, 3d] bf16 each) instead of all-reducing the [3d, 3d] fp32 gradient: X_k now (hidden behind the whole
            # decoder), dY_k at the end of backward; every rank then forms the global dW_k with one K = world*B GEMM
            fg = (train and self.world > 1 and self.dp_factor_gather and (batch_global is None or batch_global == self.world * B))
            # the entity-embedding gradient is a scatter of B x T rows into a [nE, d] table: when the table is large and
            # the global batch touches few of its rows, every rank gathers the [B, 3d] upstream gradient + the triples
            # of all ranks and runs the scatter over the GLOBAL batch instead of all-reducing the (mostly zero) table
            nE_rows = f.p("enc.e_emb.weight").shape[0]
            fg_emb = (fg and self.dp_emb_min_bytes is not None and nE_rows * d * 4 >= self.dp_emb_min_bytes
                      and self.world * B * triples.shape[1] <= nE_rows)
            if fg:
                if self.mm is not None:     # gathered by multicast stores into symmetric memory (csrc/dp_reduce.cu)
                    x_all, dy_all = self._factor_views(self._factor_ws(self.world * B, d3), self.world * B, d3)
                else:
                    x_all, dy_all = [new(self.world * B, d3, dtype=bf) for _ in range(self.n_mlp)], None
                # pipelined exchange (default): the X factors leave NOW (hidden behind the whole decoder), each dY_k as
                # soon as the backward chain has produced it.  Legacy (ARK_DP_MLP_PIPE=0, or NCCL outside the graph):
                # one grouped call for all 2 nl factors after the chain
                pipe = self.dp_mlp_pipe and self.mm is None and not (self._capturing and not self.capture_nccl)
                gathers = [(x_all[k], acts[k]) for k in range(self.n_mlp)] if pipe else []
                if fg_emb:
                    tri_p = triples.index_select(0, lay.perm_dev.long()).contiguous()      # rows in packed order
                    tri_all = torch.empty((self.world * B,) + tuple(triples.shape[1:]), device=dev, dtype=triples.dtype)
                    inv_all = new(self.world * B)
                    gathers += [(tri_all, tri_p), (inv_all, inv_cnt)]
                if gathers:
                    self._comm_action(("gathers", gathers))
            w_heads = f.fused(f.shadow, "enc.mu.weight", "enc.logv.weight", (2 * dz, d3))
            b_heads = f.fused(f.param, "enc.mu.bias", "enc.logv.bias", (2 * dz,))
            heads = new(B, 2 * dz)
            self._gemm(acts[-1], K, w_heads, K, heads, B, 2 * dz, d3, tag="enc_heads", bias=b_heads)
            z, z_b = new(B, dz), new(B, dz, dtype=bf)
            ops.reparam_kl_fwd(heads, eps, lay.perm_dev, dz, True, 1.0 / (b_g * dz), z, z_b, out[1:2])
            h0 = new(B, d)
            self._gemm(z_b, K, self._w("dec.z_proj.weight"), K, h0, B, d, dz, tag="z_proj", bias=f.p("dec.z_proj.bias"),
                       epilogue=ops.EPI_TANH)
        else:
            h0 = torch.zeros(B, d, device=dev)          # decoder-only ARK: nn.GRU's default initial state

        # ---------------- decoder forward (models.py:136-142) over packed rows
        self._leaf_join()             # token rows / embeddings / layer-0 input projection are ready
        saved = []
        u_b = x_b
        gh_ws = None if (persist or wave) else new(b0, d3)
        sync_ws = new(32 * nl * ((b0 + 15) // 16), dtype=torch.int32) if (persist or wave) else None   # cluster: 2*nl*tiles*16
        cl_ws = self._cluster_ws(L, b0, d, nl) if cluster else None
        if wave:
            # the whole stack in ONE cooperative launch: layers run as a diagonal wavefront (L+nl-1 dependent
            # steps instead of nl*L), W_ih u_t inside the recurrence (no gi buffer), dropout in the epilogue
            hp_all, out_all = new(nl, N, d, dtype=bf), new(nl, N, d, dtype=bf)
            h0_b = new(b0, d, dtype=bf)
            ops.cast_bf16(h0[:b0], h0_b)
            hp_all[:, :b0].copy_(h0_b)
            wave_gates = tuple(new(nl, N, d, dtype=bf) for _ in range(4))
            dropping = dropout and self.p_drop > 0 and nl > 1
            wave_mask = new(nl - 1, N, d, dtype=torch.uint8) if dropping else None
            stride = (N * d + 3) // 4
            with self._timed("gru_cluster_fwd" if cluster else "gru_wave_fwd", flops=4.0 * N * d * d3 * nl):
                fwd_args = (x_b, hp_all, out_all, h0 if self.has_enc else None,
                            [self._w(f"dec.gru.weight_ih_l{k}") for k in range(nl)],
                            [self._w(f"dec.gru.weight_hh_l{k}") for k in range(nl)],
                            [f.p(f"dec.gru.bias_ih_l{k}") for k in range(nl)],
                            [f.p(f"dec.gru.bias_hh_l{k}") for k in range(nl)],
                            lay.bt_dev, lay.off_dev, L, b0, d, wave_gates, wave_mask,
                            self.p_drop if dropping else 0.0, self.seed,
                            self._drop_calls * stride if self._capturing else self.philox_offset,
                            self.dyn_i if (self._capturing and dropping) else None, sync_ws)
                if cluster:
                    ops.gru_cluster_fwd(*fwd_args, cl_ws)
                else:
                    ops.gru_wave_fwd(*fwd_args)
            if dropping:
                if not self._capturing:
                    self.philox_offset += (nl - 1) * stride
                self._drop_calls += nl - 1
            u_b = out_all[nl - 1]
        for k in range(0 if not wave else nl, nl):
            if k == 0:
                gi = gi0
            else:
                gi = new(N, d3)
                self._gemm(u_b, K, self._w(f"dec.gru.weight_ih_l{k}"), K, gi, N, d3, d, tag="gru_gi",
                           bias=f.p(f"dec.gru.bias_ih_l{k}"))
            hp_b, y_b = new(N, d, dtype=bf), new(N, d, dtype=bf)
            mask = None
            if persist:
                # one cooperative launch for all L steps: W_hh slice resident in smem, state in registers; the
                # inter-layer dropout of the output rows (same Philox draw as the stand-alone kernel) is fused in
                ops.cast_bf16(h0[:b0], hp_b[:b0])
                gates = tuple(new(N, d, dtype=bf) for _ in range(4))
                drop_k = dropout and self.p_drop > 0 and k < nl - 1
                stride = (N * d + 3) // 4
                if drop_k:
                    mask = new(N, d, dtype=torch.uint8)
                with self._timed("gru_persist_fwd", flops=2.0 * N * d * d3):
                    ops.gru_persist_fwd(hp_b, h0, self._w(f"dec.gru.weight_hh_l{k}"), gi, f.p(f"dec.gru.bias_hh_l{k}"),
                                        lay.bt_dev, lay.off_dev, L, b0, d, y_b, gates, sync_ws, mask=mask,
                                        p_drop=self.p_drop if drop_k else 0.0, seed=self.seed,
                                        offset=self._drop_calls * stride if self._capturing else self.philox_offset,
                                        offset_dev=self.dyn_i if (self._capturing and drop_k) else None)
                if drop_k:
                    if not self._capturing:
                        self.philox_offset += stride
                    self._drop_calls += 1
                hp_f = None
            else:
                hp_f = new(N, d)
                hp_f[:b0].copy_(h0[:b0])
                ops.cast_bf16(hp_f[:b0], hp_b[:b0])
                y = new(N, d)
                gates = tuple(new(N, d) for _ in range(4))
                with self._timed("gru_layer_fwd", flops=2.0 * N * d * d3):
                    ops.gru_layer_fwd(hp_b, hp_f, self._w(f"dec.gru.weight_hh_l{k}"), gi, f.p(f"dec.gru.bias_hh_l{k}"),
                                      lay.bt, lay.off, L, d, y, y_b, gates, gh_ws, use_tc)
            if dropout and self.p_drop > 0 and k < nl - 1 and not persist:
                mask = new(N, d, dtype=torch.uint8)
                if self._capturing:   # replayed graphs read the running Philox offset from device memory
                    ops.dropout_bf16(y_b, self.p_drop, self.seed, self._drop_calls * ((N * d + 3) // 4), y_b, mask,
                                     self.dyn_i)
                else:
                    ops.dropout_bf16(y_b, self.p_drop, self.seed, self.philox_offset, y_b, mask)
                    self.philox_offset += (N * d + 3) // 4
                self._drop_calls += 1
            saved.append((u_b, hp_f, hp_b, gates, mask))
            u_b = y_b
            del gi
        w_out = self._w("dec.tok_emb.weight") if self.tied else self._w("dec.out.weight")
        g_wout = f.g("dec.tok_emb.weight") if self.tied else f.g("dec.out.weight")
        # Token-chunked logits: beyond `logits_chunk_rows` packed rows the [N, V] bf16 logits never exist as
        # a whole — a fixed workspace of <= chunk rows is projected, turned into its gradient by the fused softmax-CE and
        # consumed by dY and dW (accumulated over chunks) before the next chunk overwrites it.  wd-articles at 256 graphs
        # per GPU: 81 k rows x 60 944 columns = 9.9 GB of logits become a 2.0 GB workspace.
        chunk = int(self.logits_chunk_rows)
        chunked = (not autograd) and N > chunk
        deferred, keep_alive = [], []
        dx0_src, stack_weight_grads = None, None
        ext = None
        if chunked:
            ws = new(chunk, ldv, dtype=bf)
            dy = new(N, d) if train else None
            if train and self.tied:
                g_wout.zero_()        # the vocabulary dW accumulates here chunk by chunk; the embedding scatter adds on top
            for ci, r0 in enumerate(range(0, N, chunk)):
                rows = min(chunk, N - r0)
                lg = ws[:rows]
                self._gemm(u_b[r0:r0 + rows], K, w_out, K, lg, rows, V, d, tag="vocab_fwd", bias=f.p("dec.out.bias"))
                with self._timed("softmax_ce", nbytes=2.0 * rows * V * 2 + 12.0 * rows):
                    ops.softmax_ce(lg, V, tgt[r0:r0 + rows], 1.0 / n_tok_g, True, out[0:1], None)
                if train:
                    self._gemm(lg, K, w_out, MN, dy[r0:r0 + rows], rows, d, V, tag="vocab_dY")
                    self._gemm(lg, MN, u_b[r0:r0 + rows], MN, g_wout, V, d, rows, tag="vocab_dW",
                               accumulate=(ci > 0 or self.tied))
                    ops.colsum(lg, rows, V, f.g("dec.out.bias"), accumulate=ci > 0)
            if not train:
                return out

            def vocab_weight_grads():
                self._grad_ready("dec.out.bias", "dec.out.weight" if not self.tied else "dec.out.bias")
                if self.tied:
                    self._grad_ready("dec.tok_emb.weight", "dec.tok_emb.weight")
        else:
            logits = new(N, ldv, dtype=bf)
            self._gemm(u_b, K, w_out, K, logits, N, V, d, tag="vocab_fwd", bias=f.p("dec.out.bias"))
            # CE forward+backward in place (ablation_study.py:64-69): logits -> (softmax-onehot)/N_tok
            if autograd:
                # hand the logits out; the caller turns the buffer into d(loss)/d(logits) in place and resumes us
                ext = yield {"logits": logits, "heads": heads if self.has_enc else None}
                ext = ext or {}
                beta = 0.0                       # the KL term (if any) arrives through dmu / dlogv
            else:
                with self._timed("softmax_ce", nbytes=2.0 * N * V * 2 + 12.0 * N):
                    ops.softmax_ce(logits, V, tgt, 1.0 / n_tok_g, True, out[0:1], None)
            if not train:
                return out

            # ---------------- decoder backward
            # Order of the backward pass: the CHAIN first (dY -> GRU backward -> dX -> embedding scatter -> encoder), the
            # LEAVES (weight-gradient GEMMs nothing in this pass consumes) on the leaf stream, so that the two large late
            # buckets (token / entity embedding tables: all-reduce + dense Adam on the side stream) overlap them.
            dy = new(N, d)
            self._gemm(logits, K, w_out, MN, dy, N, d, V, tag="vocab_dY")                         # dY = dLogits . W

            def vocab_weight_grads(logits=logits):
                # dW = dLogits^T . Y; tied weights: on top of the embedding scatter already in the slot
                self._gemm(logits, MN, u_b, MN, g_wout, V, d, N, tag="vocab_dW", accumulate=self.tied)
                ops.colsum(logits, N, V, f.g("dec.out.bias"))
                self._grad_ready("dec.out.bias", "dec.out.weight" if not self.tied else "dec.out.bias")
                if self.tied:
                    self._grad_ready("dec.tok_emb.weight", "dec.tok_emb.weight")
            del logits
        dh0 = new(b0, d)
        if wave:
            w_t = new(2 * nl, d, d3, dtype=bf)
            for k in range(nl):
                ops.transpose_bf16(self._w(f"dec.gru.weight_hh_l{k}"), w_t[k])
                if k > 0:
                    ops.transpose_bf16(self._w(f"dec.gru.weight_ih_l{k}"), w_t[nl + k])
            dgi_all, dgh_all = new(nl, N, d3, dtype=bf), new(nl, N, d3, dtype=bf)
            with self._timed("gru_cluster_bwd" if cluster_bwd else "gru_wave_bwd", flops=4.0 * N * d * d3 * nl - 2.0 * N * d * d3):
                bwd_args = (dy, wave_gates, hp_all, wave_mask, self.p_drop if wave_mask is not None else 0.0,
                            [w_t[k] for k in range(nl)], [None] + [w_t[nl + k] for k in range(1, nl)],
                            lay.bt_dev, lay.off_dev, L, b0, d, dgi_all, dgh_all, dh0 if self.has_enc else None,
                            sync_ws)
                if cluster_bwd:
                    ops.gru_cluster_bwd(*bwd_args, cl_ws)
                else:
                    ops.gru_wave_bwd(*bwd_args)
            if self.keep is not None:
                self.keep.update(gru_out=out_all, gru_dgi=dgi_all, gru_dgh=dgh_all, gru_dh0=dh0)
            dx0_src = dgi_all[0]

            def gru_weight_grads():
                for k in range(nl - 1, -1, -1):
                    u_in = x_b if k == 0 else out_all[k - 1]
                    self._gemm(dgi_all[k], MN, u_in, MN, f.g(f"dec.gru.weight_ih_l{k}"), d3, d, N, tag="gru_dWih")
                    self._gemm(dgh_all[k], MN, hp_all[k], MN, f.g(f"dec.gru.weight_hh_l{k}"), d3, d, N, tag="gru_dWhh")
                    ops.colsum(dgi_all[k], N, d3, f.g(f"dec.gru.bias_ih_l{k}"))
                    ops.colsum(dgh_all[k], N, d3, f.g(f"dec.gru.bias_hh_l{k}"))
                    self._grad_ready(f"dec.gru.weight_ih_l{k}", f"dec.gru.bias_hh_l{k}")
            stack_weight_grads = gru_weight_grads
        else:
            if not persist:
                dgi, dgh = new(N, d3, dtype=bf), new(N, d3, dtype=bf)
            ks_bwd = persist and ops.gru_persist_bwd_ksplit(d, b0)    # K-split cluster kernel: W_hh untransposed
            if persist:
                whh_t = None if ks_bwd else new(d, d3, dtype=bf)
            else:
                dh_a, dh_b = new(b0, d), new(b0, d)
        for k in range(nl - 1, -1 if not wave else nl - 1, -1):
            u_in, hp_f, hp_b, gates, mask = saved[k]
            if mask is not None and not persist:
                ops.dropout_bwd(dy, mask, self.p_drop, dy)
            if persist:
                dgi, dgh = new(N, d3, dtype=bf), new(N, d3, dtype=bf)   # per layer: read later by the leaf stream
                if not ks_bwd:
                    ops.transpose_bf16(self._w(f"dec.gru.weight_hh_l{k}"), whh_t)
                with self._timed("gru_persist_bwd", flops=2.0 * N * d * d3):
                    # (the backward of the inter-layer dropout is applied to dy on the fly)
                    ops.gru_persist_bwd(dy, gates, hp_b, whh_t, lay.bt_dev, lay.off_dev, L, b0, d, dgi, dgh, dh0,
                                        k != nl - 1, sync_ws, Whh_b=self._w(f"dec.gru.weight_hh_l{k}") if ks_bwd else None,
                                        dy_mask=mask, p_drop=self.p_drop if mask is not None else 0.0)
            else:
                with self._timed("gru_layer_bwd", flops=2.0 * N * d * d3):
                    dh_k = ops.gru_layer_bwd(dy, gates, hp_f, self._w(f"dec.gru.weight_hh_l{k}"), lay.bt, lay.off, L, d,
                                             dgi, dgh, dh_a, dh_b, use_tc)
                if k == nl - 1:
                    dh0.copy_(dh_k)
                else:
                    ops.add_(dh0, dh_k, dh0, None)
            def layer_weight_grads(k=k, dgi=dgi, dgh=dgh, u_in=u_in, hp_b=hp_b):
                self._gemm(dgi, MN, u_in, MN, f.g(f"dec.gru.weight_ih_l{k}"), d3, d, N, tag="gru_dWih")
                self._gemm(dgh, MN, hp_b, MN, f.g(f"dec.gru.weight_hh_l{k}"), d3, d, N, tag="gru_dWhh")
                ops.colsum(dgi, N, d3, f.g(f"dec.gru.bias_ih_l{k}"))
                ops.colsum(dgh, N, d3, f.g(f"dec.gru.bias_hh_l{k}"))
                self._grad_ready(f"dec.gru.weight_ih_l{k}", f"dec.gru.bias_hh_l{k}")
            if k == 0 and persist:
                dx0_src = dgi         # layer 0: dX only feeds the embedding scatter -> with it on the leaf stream
            else:
                self._gemm(dgi, K, self._w(f"dec.gru.weight_ih_l{k}"), MN, dy, N, d, d3, tag="gru_dX")   # the chain
            if persist:
                keep_alive.append((dgi, dgh))
                self._leaf(layer_weight_grads)
            else:
                layer_weight_grads()
        self._release_comm()          # (no-op unless collectives were held back for the persistent GRU kernels)

        def embedding_grads():
            # d(loss)/d(decoder inputs) -> token (+ position) embedding scatter -> vocabulary weight gradients: nothing
            # in the rest of the pass reads them, so they leave the chain (the encoder backward continues on the main
            # stream with dh0 only)
            if dx0_src is not None:
                self._gemm(dx0_src, K, self._w("dec.gru.weight_ih_l0"), MN, dy, N, d, d3, tag="gru_dX")
            if not (chunked and self.tied):       # (chunked + tied: the slot already holds the accumulated vocabulary dW)
                f.g("dec.tok_emb.weight").zero_()
            with self._timed("tok_scatter_add", nbytes=N * d * 12.0):
                ops.tok_scatter_add(dy, tok, f.g("dec.tok_emb.weight"))
            if not self.tied:             # (tied: final once the vocabulary dW has been added on top of the scatter)
                self._grad_ready("dec.tok_emb.weight", "dec.tok_emb.weight")
            if not self.has_enc:
                g_pos = f.g("dec.pos_emb.weight")
                g_pos.zero_()
                ops.tok_scatter_add(dy, row_t, g_pos)      # d pos_emb[t] = sum of dX over the rows of step t
                self._grad_ready("dec.pos_emb.weight", "dec.pos_emb.weight")
            if self.leaf_embedding:
                vocab_weight_grads()
                if stack_weight_grads is not None:
                    stack_weight_grads()

        def tail_weight_grads():
            vocab_weight_grads()
            if stack_weight_grads is not None:
                stack_weight_grads()
        if self.leaf_embedding:
            self._leaf(embedding_grads)
        else:           # dX_0 + scatter stay on the chain's stream; only the weight gradients fork
            embedding_grads()
            self._leaf(tail_weight_grads)

        if not self.has_enc:
            for fn in deferred:
                fn()
            self._leaf_join()
            return out

        # ---------------- h0 = tanh(W_z z + b_z), reparameterisation, KL
        dpre, dpre_b = new(B, d), new(B, d, dtype=bf)
        ops.tanh_bwd(dh0, h0, dpre, dpre_b)
        def z_proj_weight_grads():
            self._gemm(dpre_b, MN, z_b, MN, f.g("dec.z_proj.weight"), d, dz, B, tag="z_proj_bwd")
            ops.colsum(dpre, B, d, f.g("dec.z_proj.bias"))
        self._leaf(z_proj_weight_grads)
        dz_in = new(B, dz)
        self._gemm(dpre_b, K, self._w("dec.z_proj.weight"), MN, dz_in, B, dz, d, tag="z_proj_bwd")
        ld_dh = _up8(2 * dz)
        dheads = torch.zeros(B, ld_dh, device=dev)
        dheads_b = torch.zeros(B, ld_dh, device=dev, dtype=bf)
        if self._capturing:      # replayed graphs read beta from device memory (ablation_study.py:589-591: per-epoch schedule)
            ops.reparam_kl_bwd(heads, eps, lay.perm_dev, dz_in, dz, True, 1.0 / (b_g * dz), dheads, dheads_b,
                               beta_dev=self.dyn_beta)
        else:
            ops.reparam_kl_bwd(heads, eps, lay.perm_dev, dz_in, dz, True, beta / (b_g * dz), dheads, dheads_b,
                               dmu_ext=None if ext is None else ext.get("dmu"),
                               dlogv_ext=None if ext is None else ext.get("dlogv"))
        g_wh = f.fused(f.grad, "enc.mu.weight", "enc.logv.weight", (2 * dz, d3))
        g_bh = f.fused(f.grad, "enc.mu.bias", "enc.logv.bias", (2 * dz,))
        def heads_weight_grads():
            self._gemm(dheads_b[:, :2 * dz], MN, acts[-1], MN, g_wh, 2 * dz, d3, B, tag="enc_heads_bwd")
            ops.colsum(dheads, B, 2 * dz, g_bh)
            self._grad_ready("dec.z_proj.weight", "enc.logv.bias")
            if self.mm is not None:      # every decoder gradient is final: exchange them NOW, behind the encoder-MLP chain,
                self._flush_bucket()     # instead of at the end of the step together with the embedding rows
        da = new(B, d3)
        self._gemm(dheads_b[:, :2 * dz], K, w_heads, MN, da, B, d3, 2 * dz, tag="enc_heads_bwd")
        self._leaf(heads_weight_grads)

        # ---------------- encoder MLP + pooled gather backward
        wb = self.world * B
        mlp_items = []
        for k in range(self.n_mlp - 1, -1, -1):
            dp_b = new(B, d3, dtype=bf)
            ops.gelu_bwd(da, pres[k], None, dp_b)
            if fg:
                # data parallel: the layers' dY / X factors are gathered (one grouped call after the chain below), the
                # GLOBAL dW_k = dY_all^T X_all formed and Adam applied, all on the side streams: nothing on this stream
                # waits for it before the step ends
                dp_all_k = dy_all[k] if dy_all is not None else new(wb, d3, dtype=bf)
                keep_alive.append((dp_b, dp_all_k))
                if pipe:        # dY_k leaves now; the global dW_k GEMM follows it on the side streams
                    self._comm_action(("mlp_dw_gather", k, dp_all_k, dp_b, x_all[k]))
            da = new(B, d3)
            self._gemm(dp_b, K, self._w(f"enc.mlp.{2 * k}.weight"), MN, da, B, d3, d3, tag="enc_mlp_dX")
            if fg and pipe:     # W_k has been read for the last time: Adam may rewrite it
                self._comm_action(("mlp_dw_adam", k))
            elif fg:
                mlp_items.append((k, dp_all_k, dp_b, x_all[k], acts[k]))
            if not fg:
                def mlp_weight_grads(k=k, dp_b=dp_b):
                    self._gemm(dp_b, MN, acts[k], MN, f.g(f"enc.mlp.{2 * k}.weight"), d3, d3, B, tag="enc_mlp_dW")
                    ops.colsum(dp_b, B, d3, f.g(f"enc.mlp.{2 * k}.bias"))
                    self._grad_ready(f"enc.mlp.{2 * k}.weight", f"enc.mlp.{2 * k}.bias")
                keep_alive.append(dp_b)
                self._leaf(mlp_weight_grads)
        if fg and not pipe:      # AFTER the last dX GEMM is queued (the side streams' Adam rewrites W_k, which those GEMMs read): ONE
            self._comm_action(("mlp_dw", mlp_items))      # grouped all-gather of every layer's two factors (2 nl small
                                                          # messages are launch-latency bound: ~0.1 ms per NCCL call)
        gE, gR = f.g("enc.e_emb.weight"), f.g("enc.r_emb.weight")
        if fg and fg_emb:       # (opt-in) the embedding scatter over the GLOBAL batch from gathered factors
            da_all = new(wb, d3)
            self._comm_action(("gather", da_all, da))
            self._flush_bucket()
            self._comm_action(("join",))            # every rank's factors have arrived
            with self._timed("gather_pool_bwd", nbytes=gE.numel() * 4.0 + wb * d3 * 4 + 3.0 * lay.n_triples * self.world * d * 8):
                gR.zero_()
                gE.zero_()
                ops.gather_pool_bwd(da_all, tri_all, None, inv_all, self.pad_rid, self.pad_eid, gE, gR)
            self._grad_ready("enc.r_emb.weight", "enc.e_emb.weight", reduced=True)
        else:
            with self._timed("gather_pool_bwd", nbytes=gE.numel() * 4.0 + B * d3 * 4 + 3.0 * lay.n_triples * d * 8):
                gR.zero_()
                gE.zero_()
                ops.gather_pool_bwd(da, triples, lay.perm_dev, inv_cnt, self.pad_rid, self.pad_eid, gE, gR)
            self._grad_ready("enc.r_emb.weight", "enc.e_emb.weight")
        for fn in deferred:
            fn()
        self._leaf_join()
        del keep_alive
        return out

    # ------------------------------------------------------------------ leaf stream
    def _leaf(self, fn):
        """Run `fn` (launches of leaf work: nothing later on the current stream reads what it writes before `_leaf_join`)
        on the leaf stream, ordered after everything queued on the current stream so far.  Inline when the leaf stream is
        off, or while a data-parallel step is captured (its graph segments end at arbitrary points of the pass and a
        capture cannot end with un-joined forked work)."""
        if (not self.use_leaf_stream or (self._capturing and self.world > 1 and not self.capture_nccl)
                or self.prof is not None and not self._capturing):
            fn()
            return
        ev = torch.cuda.Event()
        ev.record()
        self.leaf_stream.wait_event(ev)
        with torch.cuda.stream(self.leaf_stream):
            fn()
        self._leaf_used = True

    def _leaf_join(self):
        if self._leaf_used:
            torch.cuda.current_stream().wait_stream(self.leaf_stream)
            self._leaf_used = False

    def _join_side(self):
        """The current stream waits for the side stream(s): collectives and the updates chained behind them."""
        cur = torch.cuda.current_stream()
        cur.wait_stream(self.comm_stream)
        if self.post_stream is not None and self.post_stream is not self.comm_stream:
            if self._capturing:      # a step that queued nothing on the post stream: bring it into the capture first
                self.post_stream.wait_stream(self.comm_stream)
            cur.wait_stream(self.post_stream)

    def _post(self):
        """Stream context for work that follows a collective (ordered after everything queued on the side stream)."""
        post = self.post_stream if self.post_stream is not None else self.comm_stream
        if post is not self.comm_stream:
            ev = torch.cuda.Event()
            ev.record(self.comm_stream)
            post.wait_event(ev)
        return torch.cuda.stream(post)

    def _comm_waits_for_leaf(self):
        """Buckets may contain gradients written on the leaf stream: the side stream waits for it too."""
        if self._leaf_used:
            ev = torch.cuda.Event()
            ev.record(self.leaf_stream)
            self.comm_stream.wait_event(ev)

    # ------------------------------------------------------------------ gradient exchange + bucketed update
    def _grad_ready(self, first, last, reduced=False):
        """Gradient slots first..last (contiguous in the flat layout) are final AND their weights are no longer read
        by the rest of this backward pass.  They are merged into buckets; a full bucket is handed to the side stream,
        which (data parallel) sums it over ranks with NCCL and (inside a train step) applies Adam to exactly that
        slice of the flat buffers — both overlap the remaining backward kernels on the main stream.
        reduced=True: the slots already hold the GLOBAL gradient (factor all-gather): Adam only, no all-reduce."""
        if self.world == 1 and self._upd is None:
            return
        s, e = self.flat.span(first, last)
        if reduced:
            self._flush_bucket()
            self._comm_action(("bucket", [(s, e)], False))
            return
        merge_span(self._pending, s, e)
        if sum(b - a for a, b in self._pending) >= self.bucket_elems:
            self._flush_bucket()

    def _flush_bucket(self):
        if not self._pending:
            return
        spans, self._pending = self._pending, []
        self._comm_action(("bucket", spans, True))

    def _comm_action(self, action):
        """Side-stream work that may contain NCCL calls: ("bucket", spans, allreduce) | ("gather", out, inp) |
        ("join",).  Graph mode under data parallelism ends the graph segment here; the replay loop runs the action
        eagerly while the NEXT segment runs (NCCL kernels are never captured: capturing them next to the cooperative
        GRU kernels hung at 2 ranks)."""
        if self._hold_comm:
            self._held.append(action)
            return
        if self._capturing and self.world > 1 and not self.capture_nccl:
            self._segment_break([action])
            return
        self._run_action(action, self._upd)

    def _comm_actions(self, actions):
        if self._hold_comm:
            self._held.extend(actions)
            return
        if self._capturing and self.world > 1 and not self.capture_nccl:
            self._segment_break(list(actions))      # one segment break for all of them
            return
        for a in actions:
            self._run_action(a, self._upd)

    def _release_comm(self):
        """End of the stretch of per-layer persistent GRU kernels: everything that was held back goes out at once.
        Those kernels are cooperative launches of 128 CTAs; an NCCL kernel that holds a few SMs keeps them from
        starting, so a collective issued between two of them is serialised with the recurrence (and with the
        slowest rank) instead of overlapping it — measured at 8 GPUs on syn-types: 7 bucket all-reduces took
        1.5 ms and stretched gru_persist_bwd 0.41 -> 1.08 ms.  Held back, the same gradients leave in one or two
        large all-reduces that overlap the (non-cooperative) encoder backward."""
        self._hold_comm = False
        held, self._held = self._held, []
        if held:
            self._comm_actions(held)

    def _run_action(self, action, upd):
        kind = action[0]
        if kind == "bucket":
            self._bucket_async(action[1], upd, allreduce=action[2])
        elif kind in ("gather", "gathers"):      # all ranks' [B, n] rows -> [world * B, n], on the side stream
            pairs = [(action[1], action[2])] if kind == "gather" else list(action[1])
            ev = torch.cuda.Event()
            ev.record()
            self.comm_stream.wait_event(ev)
            self._comm_waits_for_leaf()
            with torch.cuda.stream(self.comm_stream):
                with self._timed("nccl_all_gather", nbytes=float(sum(o.numel() * o.element_size() for o, _ in pairs))):
                    if len(pairs) == 1:
                        torch.distributed.all_gather_into_tensor(pairs[0][0], pairs[0][1], group=self.group)
                    else:   # ONE NCCL group launch for all of them (small messages are launch-latency bound)
                        with torch.distributed._coalescing_manager(group=self.group, device=self.device, async_ops=False):
                            for o, i in pairs:
                                torch.distributed.all_gather_into_tensor(o, i, group=self.group)
        elif kind == "mlp_dw":      # [(k, dY_all, dY_local, X_all, X_local)]: gathers -> global dW_k GEMMs -> bias -> Adam
            items = action[1]
            ev = torch.cuda.Event()
            ev.record()
            self.comm_stream.wait_event(ev)
            with torch.cuda.stream(self.comm_stream):
                if self.mm is not None:     # every rank multicasts its rows into its slot of all ranks' gathered matrices
                    wb, d3 = items[0][1].shape
                    ws, B_ = self._factor_ws(wb, d3), wb // self.world
                    slot = lambda t: t.data_ptr() - ws.buf.data_ptr() + self.mm.rank * B_ * d3 * 2
                    pairs = [p for _, dp_all_k, dp_b, x_all_k, x_k in items for p in ((dp_b, slot(dp_all_k)), (x_k, slot(x_all_k)))]
                    for i in range(0, len(pairs), 8):
                        with self._timed("dp_allgather_mc", nbytes=float(sum(t.numel() * 2 for t, _ in pairs[i:i + 8]) * self.world)):
                            ops.dp_allgather_mc(self.mm, ws, pairs[i:i + 8], ctas=self.dp_mm_ctas)
                else:
                    with self._timed("nccl_all_gather", nbytes=float(sum(it[1].numel() * 4 for it in items))):
                        with torch.distributed._coalescing_manager(group=self.group, device=self.device, async_ops=False):
                            for _, dp_all_k, dp_b, x_all_k, x_k in items:
                                torch.distributed.all_gather_into_tensor(dp_all_k, dp_b, group=self.group)     # dY factor
                                torch.distributed.all_gather_into_tensor(x_all_k, x_k, group=self.group)       # X factor
            f = self.flat
            # the global dW_k GEMMs (tensor-bound) on the leaf stream, each layer's Adam (HBM-bound) on the post stream
            # behind its GEMM: GEMM k+1 overlaps Adam k
            split = self.use_leaf_stream and self.post_stream is not self.comm_stream and not (self.prof is not None and not self._capturing)
            gstream = self.leaf_stream if split else (self.post_stream if self.post_stream is not None else self.comm_stream)
            ev = torch.cuda.Event()
            ev.record(self.comm_stream)
            gstream.wait_event(ev)
            if split:
                self.post_stream.wait_event(ev)
                self._leaf_used = True
            for k, dp_all_k, _, x_all_k, _ in items:
                d3 = dp_all_k.shape[1]
                with torch.cuda.stream(gstream):
                    self._gemm(dp_all_k, MN, x_all_k, MN, f.g(f"enc.mlp.{2 * k}.weight"), d3, d3, dp_all_k.shape[0],
                               tag="enc_mlp_dW_global")
                    ops.colsum(dp_all_k, dp_all_k.shape[0], d3, f.g(f"enc.mlp.{2 * k}.bias"), deterministic=True)   # ranks agree bitwise
                    if split:
                        evk = torch.cuda.Event()
                        evk.record(gstream)
                if upd is not None:
                    if split:
                        self.post_stream.wait_event(evk)
                    with torch.cuda.stream(self.post_stream if split else gstream):
                        s_, e_ = f.span(f"enc.mlp.{2 * k}.weight", f"enc.mlp.{2 * k}.bias")
                        self._adam_slice(s_, e_, upd)
        elif kind == "mlp_dw_gather":   # (k, dY_all, dY_local, X_all): dY_k -> all ranks; global dW_k + bias on the post stream
            _, k, dp_all_k, dp_b, x_all_k = action
            ev = torch.cuda.Event()
            ev.record()
            self.comm_stream.wait_event(ev)
            with torch.cuda.stream(self.comm_stream):
                with self._timed("nccl_all_gather", nbytes=float(dp_all_k.numel() * 2)):
                    torch.distributed.all_gather_into_tensor(dp_all_k, dp_b, group=self.group)
            f, d3 = self.flat, dp_all_k.shape[1]
            with self._post():
                self._gemm(dp_all_k, MN, x_all_k, MN, f.g(f"enc.mlp.{2 * k}.weight"), d3, d3, dp_all_k.shape[0],
                           tag="enc_mlp_dW_global")
                ops.colsum(dp_all_k, dp_all_k.shape[0], d3, f.g(f"enc.mlp.{2 * k}.bias"), deterministic=True)   # ranks agree bitwise
        elif kind == "mlp_dw_adam":     # (k,): everything queued on the current stream (the dX GEMM that reads W_k) first
            if upd is not None:
                k, f = action[1], self.flat
                post = self.post_stream if self.post_stream is not None else self.comm_stream
                ev = torch.cuda.Event()
                ev.record()
                post.wait_event(ev)
                with torch.cuda.stream(post):
                    s_, e_ = f.span(f"enc.mlp.{2 * k}.weight", f"enc.mlp.{2 * k}.bias")
                    self._adam_slice(s_, e_, upd)
        elif kind == "join":        # the main stream needs what the side stream produced
            self._join_side()
        else:
            raise ValueError(kind)

    def _bucket_async(self, spans, upd=None, allreduce=True):
        """On the side stream, ordered after everything queued on the current stream: all-reduce (world > 1) and,
        inside a train step, Adam over the slices."""
        upd = self._upd if upd is None else upd
        ev = torch.cuda.Event()
        ev.record()
        self.comm_stream.wait_event(ev)
        self._comm_waits_for_leaf()
        f = self.flat
        if self.world > 1 and allreduce and self.mm is not None:
            # switch-reduced gradients -> Adam on the owned slice -> multicast parameters, ONE kernel (K11); without an
            # update in flight (gradients only) the same kernel multicasts the sum back into the gradient buffer
            spans4 = [(s, min(f.numel, (e + 3) // 4 * 4)) for s, e in spans]
            with torch.cuda.stream(self.comm_stream):
                for i in range(0, len(spans4), 16):
                    part = spans4[i:i + 16]
                    with self._timed("dp_reduce_adam", nbytes=4.0 * sum(e - s for s, e in part)):
                        if upd is None:
                            ops.dp_reduce_adam(self.mm, part, f.exp_avg, f.exp_avg_sq, 0, ctas=self.dp_mm_ctas)
                        elif upd[0] == "dyn":
                            ops.dp_reduce_adam(self.mm, part, f.exp_avg, f.exp_avg_sq, 1, 0.0, self.betas[0], self.betas[1],
                                               self.eps, hyper=self.dyn_f, ctas=self.dp_mm_ctas)
                        else:
                            ops.dp_reduce_adam(self.mm, part, f.exp_avg, f.exp_avg_sq, 1, upd[1], self.betas[0], self.betas[1],
                                               self.eps, step=self.step_count, ctas=self.dp_mm_ctas)
                    if upd is not None:
                        self._mm_spans.update(part)
            return
        if self.world > 1 and allreduce:
            with torch.cuda.stream(self.comm_stream):
                with self._timed("nccl_all_reduce", nbytes=4.0 * sum(e - s for s, e in spans)):
                    if len(spans) == 1:
                        torch.distributed.all_reduce(f.grad[spans[0][0]:spans[0][1]], group=self.group)
                    else:   # the slices of one bucket leave in ONE grouped NCCL launch (they are launch-latency bound)
                        with torch.distributed._coalescing_manager(group=self.group, device=self.device, async_ops=False):
                            for (s, e) in spans:
                                torch.distributed.all_reduce(f.grad[s:e], group=self.group)
        if upd is not None:
            with self._post():
                for (s, e) in spans:
                    self._adam_slice(s, e, upd)

    def _adam_slice(self, s, e, upd):
        f = self.flat
        with self._timed("adam_flat", nbytes=30.0 * (e - s)):
            if upd[0] == "dyn":     # graph replay: lr / bias corrections live in device memory
                ops.adam_flat_dyn(f.param[s:e], f.grad[s:e], f.exp_avg[s:e], f.exp_avg_sq[s:e], f.shadow[s:e],
                                  self.dyn_f, self.betas[0], self.betas[1], self.eps)
            else:
                ops.adam_flat(f.param[s:e], f.grad[s:e], f.exp_avg[s:e], f.exp_avg_sq[s:e], f.shadow[s:e],
                              upd[1], self.betas[0], self.betas[1], self.eps, self.step_count)

    def _factor_ws(self, wb, d3):
        """Symmetric multicast home of the gathered encoder-MLP factors: [2 n_mlp, world * B, 3d] bf16 (X_all_k, dY_all_k).
        Creating it is a COLLECTIVE host-side rendezvous: before any capture, in the same order on every rank."""
        ws = self._factor_wss.get((wb, d3))
        if ws is None:
            if self._capturing:
                raise RuntimeError("the symmetric factor workspace must exist before the step is captured")
            from .symm import SymmBuf
            ws = self._factor_wss[(wb, d3)] = SymmBuf(2 * self.n_mlp * wb * d3 * 2, self.device, self.group)
        return ws

    def _factor_views(self, ws, wb, d3):
        nb = wb * d3 * 2
        mats = [ws.buf[i * nb:(i + 1) * nb].view(torch.bfloat16).view(wb, d3) for i in range(2 * self.n_mlp)]
        return mats[0::2], mats[1::2]          # X_all_k, dY_all_k

    def gather_adam_state(self):
        """COLLECTIVE (every rank).  With the multicast exchange a rank maintains exp_avg / exp_avg_sq only for the
        slices it owns; before the optimiser state is read as a whole (checkpoint on rank 0, a switch to another
        exchange) every owner broadcasts its slices."""
        if self.mm is None or not self._mm_spans:
            return
        self._sync_grads()
        torch.cuda.synchronize(self.device)
        f, dist = self.flat, torch.distributed
        for s, e in sorted(self._mm_spans):
            for r, (lo, hi) in enumerate(self.mm.owned(s, e)):
                if hi > lo:
                    src = dist.get_global_rank(self.group, r)
                    dist.broadcast(f.exp_avg[lo:hi], src, group=self.group)
                    dist.broadcast(f.exp_avg_sq[lo:hi], src, group=self.group)
        torch.cuda.synchronize(self.device)

    def _sync_grads(self):
        self._leaf_join()
        self._flush_bucket()
        self._join_side()

    # ------------------------------------------------------------------ optimiser
    def adam_step(self, lr=None):
        """Dense Adam over the whole flat buffer (torch.optim.Adam defaults, ablation_study.py:571) — the separate
        optimizer.step() of the reference's loop.  train_step() instead updates bucket by bucket during backward."""
        self._sync_grads()
        if self._mm_spans:       # earlier fused steps kept a sharded Adam state: make it whole before the dense update
            self.gather_adam_state()
            self._mm_spans.clear()
        self.step_count += 1
        f = self.flat
        with self._timed("adam_flat", nbytes=30.0 * f.numel):
            ops.adam_flat(f.param, f.grad, f.exp_avg, f.exp_avg_sq, f.shadow, self.lr if lr is None else lr,
                          self.betas[0], self.betas[1], self.eps, self.step_count)

    def train_step(self, triples, seq, lay, eps, beta, lr=None, n_tok_global=None, batch_global=None):
        """zero_grad + forward + backward (+ all-reduce) + Adam: ablation_study.py:43,59-76.  Adam runs per gradient
        bucket on the side stream as soon as the bucket is final, overlapping the rest of backward."""
        self.step_count += 1
        self._upd = ("host", self.lr if lr is None else float(lr))
        try:
            out = self.forward_backward(triples, seq, lay, eps, beta, n_tok_global, batch_global, train=True)
            self._sync_grads()          # last bucket + join: the next forward must see every updated weight
        finally:
            self._upd = None
        self.stats[0:2] += out
        self.stats[2] += 1
        return out

    # ------------------------------------------------------------------ CUDA-graph replay of the whole step
    def train_step_graphed(self, triples, seq, lay, eps, beta, lr=None, n_tok_global=None, batch_global=None):
        """Same as train_step, but the ~190 launches of the step are captured ONCE per batch layout
        (B, T, per-step row counts, normalisers — NOT beta, lr or the Philox offset: those live in a small device
        buffer refreshed by one H2D copy per step) into CUDA graphs and replayed: the host cost of a step
        drops to a few small input copies + a handful of graph launches.  Layouts that never repeat (ragged real
        data) should use train_step.
        Single GPU: the per-bucket Adam launches are a parallel branch of the one captured graph.  Under data
        parallelism the step is captured as a CHAIN of graph segments cut at the gradient-bucket boundaries; between
        two segments the replay loop issues that bucket's NCCL all-reduce + Adam eagerly on the side stream, so they
        overlap the following segment of backward exactly as in the eager step."""
        key = (None if triples is None else tuple(triples.shape), tuple(seq.shape), lay.bt.tobytes(),
               n_tok_global, batch_global, self.prof is not None)     # (a profiling capture carries event nodes)
        ent = self._graphs.get(key)
        self.step_count += 1
        lr = self.lr if lr is None else float(lr)
        b1, b2 = self.betas
        # pageable sources: the driver stages them at call time, so the next step may overwrite nothing in flight
        host = np.zeros(24, dtype=np.uint8)
        host[:8].view(np.float32)[:] = (lr / (1.0 - b1 ** self.step_count), 1.0 / math.sqrt(1.0 - b2 ** self.step_count))
        host[8:16].view(np.int64)[0] = self.philox_offset
        host[16:20].view(np.float32)[0] = beta
        self._dyn_raw.copy_(torch.from_numpy(host))
        if ent is None:
            dev = self.device
            st = {"triples": None if triples is None else triples.to(dev, copy=True), "seq": seq.to(dev, copy=True),
                  "eps": None if eps is None else eps.to(dev, copy=True),
                  "lay": PackedLayout(perm=lay.perm, lens=lay.lens, bt=lay.bt, off=lay.off, n_tok=lay.n_tok,
                                      n_triples=lay.n_triples, L=lay.L, perm_dev=lay.perm_dev.clone(),
                                      bt_dev=lay.bt_dev.clone(), off_dev=lay.off_dev.clone())}
            from . import _C
            segs = []                     # [(CUDAGraph, [(s, e) gradient spans to all-reduce after it])]
            pool = torch.cuda.graph_pool_handle()
            cur = {"g": None}

            def begin():
                cur["g"] = torch.cuda.CUDAGraph()
                cur["g"].capture_begin(pool=pool)

            def brk(actions):
                cur["g"].capture_end()
                segs.append((cur["g"], list(actions)))
                begin()

            # warm-up outside capture is NOT wanted (it would apply an extra optimiser step): capture directly
            b0 = int(lay.bt[0])
            if (self.backend == "tc" and self.gru_mode in ("auto", "cluster") and not self.force_unfused_gru
                    and self._use_gru_cluster(self.d, b0, self.nl, lay.L)):
                self._cluster_ws(lay.L, b0, self.d, self.nl)      # scratch must exist before capture
            if (self.mm is not None and self.has_enc and self.dp_factor_gather and triples is not None
                    and (batch_global is None or batch_global == self.world * triples.shape[0])):
                self._factor_ws(self.world * triples.shape[0], 3 * self.d)      # (a collective rendezvous)
            torch.cuda.synchronize()
            n0 = _C.lib().launch_count()
            cap = torch.cuda.Stream(device=dev)
            cap.wait_stream(torch.cuda.current_stream())
            self._capturing, self._segment_break, self._upd = True, brk, ("dyn", None)
            try:
                with torch.cuda.stream(cap):
                    begin()
                    out = self.forward_backward(st["triples"], st["seq"], st["lay"], st["eps"], beta, n_tok_global,
                                                batch_global, train=True)
                    self._flush_bucket()              # the last gradient slices (world > 1: cuts a segment)
                    if self.world == 1 or self.capture_nccl:
                        self._join_side()    # join the Adam (+ NCCL) branches
                    self.stats[0:2] += out
                    self.stats[2] += 1
                    cur["g"].capture_end()
                    segs.append((cur["g"], []))
            finally:
                self._capturing, self._upd, self._segment_break = False, None, None     # (brk's closure pins the last graph)
            torch.cuda.current_stream().wait_stream(cap)
            st["out"], st["segs"] = out, segs
            st["philox_per_step"] = self._drop_calls * ((lay.n_tok * self.d + 3) // 4)
            st["n_launch"] = _C.lib().launch_count() - n0
            if self.prof is not None:
                st["prof"], self.prof = list(self.prof), []
            if len(self._graphs) >= self.max_graphs:       # bounded: every entry pins a private pool of activations
                self._graphs.pop(next(iter(self._graphs)))
            self._graphs[key] = ent = st
        else:
            if triples is not None:
                ent["triples"].copy_(triples, non_blocking=True)
                ent["eps"].copy_(eps, non_blocking=True)
            ent["seq"].copy_(seq, non_blocking=True)
            ent["lay"].perm_dev.copy_(lay.perm_dev, non_blocking=True)
        segs = ent["segs"]
        for g, actions in segs:
            g.replay()
            for a in actions:
                self._run_action(a, ("dyn", None))
        if self.world > 1:
            self._join_side()     # the next forward needs every updated weight
        self.launches_replayed += ent["n_launch"]
        self.philox_offset += ent["philox_per_step"]
        self.last_graph = ent
        return ent["out"]

    def release_graphs(self):
        """Drop every captured step graph (and the private pools they pin).  Under data parallelism the graphs contain
        NCCL kernels: release them BEFORE torch.distributed.destroy_process_group(), which otherwise waits forever."""
        torch.cuda.synchronize()
        self._graphs.clear()
        self.last_graph = None
        self._segment_break = None
        import gc
        gc.collect()

    def eval_step(self, triples, seq, lay, eps, beta):
        """Forward only (validation loss, ablation_study.py:92-187 without the generation part)."""
        return self.forward_backward(triples, seq, lay, eps, beta, train=False)

    def read_stats(self, beta, reset=True):
        """(avg_loss, avg_ce, avg_kl) since the last reset — ONE device->host read instead of the reference's
        three .item() calls per step (ablation_study.py:78-80)."""
        st = self.stats.clone()
        if self.world > 1:
            # ce/kl were normalised by the GLOBAL counts, so the global value is the sum over ranks
            torch.distributed.all_reduce(st[0:2], group=self.group)
        ce, kl, n = st[:3].tolist()
        if reset:
            self.stats.zero_()
        n = max(n, 1.0)
        return (ce + beta * kl) / n, ce / n, kl / n
