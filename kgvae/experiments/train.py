"""``python -m kgvae.experiments.train --config configs/autoreg_<dataset>.yaml`` — KG-VAE training on B200.

Same CLI, YAML schema, epoch flow, logged keys and checkpoint dictionary as the reference trainer
(/root/reference/kgvae/experiments/train.py:241-624), with the SAIL / ELBO branches the reference keeps in
ablation_study.py:59-81,589-591 merged in (SURVEY.md finding 3).  Kept: 1-based `epoch` in logs and checkpoints,
`dataset_meta = {dataset, n_entities, n_relations}`, `save_every` default 10, `objective` = best posterior
compression bits updated every `compression_log_every` epochs (`val/compression_*` keys), the final evaluation on
the validation (and, with `use_test_for_final_eval`, test) split logged as `final_val/*` / `final_test/*`.
NOT reproduced (needs the `intelligraphs` verifiers, which cannot be installed here): the `verify_every` semantic
validity / novelty evaluation of generated graphs and its `verification/*` keys.  Underneath:

  * the train step is the fused ark_b200 ELBO engine (no autograd graph, no [B,L,V] probabilities);
  * the optimiser is the fused flat Adam (same `optimizer_state_dict` layout);
  * loss scalars are accumulated on the device and read back once per epoch instead of 3 .item() per step;
  * data parallelism: run the same module under ``torchrun --nproc_per_node=G`` (one rank per GPU; NCCL
    gradient all-reduce overlapped with backward; global CE/KL normalisers so the result equals the
    single-process step on the concatenated batch);
  * when `intelligraphs` is not installed (or ``synthetic: true``) the five dataset names resolve to
    synthetic graphs of the same shape (ark_b200/synthetic.py) so the path runs offline.
"""
from __future__ import annotations

import argparse
import math
import os
import time
import warnings

import numpy as np
import torch
import yaml

from ark_b200.optim import FusedAdam
from ark_b200.synthetic import DATASET_SHAPES
from kgvae.model.models import ARK, SAIL
from kgvae.model.utils import (GraphSeqDataset, build_batch, canonical_graph_string, canonicalize, ints_to_labels,
                               seq_to_triples)

SPECIAL = {"PAD": 0, "BOS": 1, "EOS": 2}


# --------------------------------------------------------------------------------------------- logging
class _NullRun:
    id = "offline"


class _Log:
    """wandb when it is importable and not disabled; otherwise a dict-printing stand-in with the same calls."""

    def __init__(self, project, entity, config, name, rank):
        self.wandb, self.config, self.run = None, dict(config), _NullRun()
        if rank == 0 and os.environ.get("WANDB_MODE", "") != "disabled":
            try:
                import wandb
                kw = dict(project=project, config=config, name=name, anonymous="allow")
                if entity:
                    kw["entity"] = entity
                wandb.init(**kw)
                self.wandb, self.config, self.run = wandb, dict(wandb.config), wandb.run
            except Exception as e:  # no network / not installed
                print(f"[train] wandb unavailable ({type(e).__name__}); logging to stdout")

    def log(self, d):
        if self.wandb is not None:
            self.wandb.log(d)
        else:
            print("[log]", {k: (round(v, 6) if isinstance(v, float) else v) for k, v in d.items()})

    def finish(self):
        if self.wandb is not None:
            self.wandb.finish()


# --------------------------------------------------------------------------------------------- data
def load_graphs(config, seed=1234):
    """(train, val, test, (e2i,i2e), (r2i,i2r), (min_edges,max_edges)) — intelligraphs when present, else
    synthetic graphs shaped like the named dataset."""
    name = config["dataset"]
    mode = str(config.get("synthetic", "auto")).lower()
    if mode not in ("true", "1"):
        try:
            from intelligraphs.data_loaders import load_data_as_list
            tr, va, te, ent, rel, edges, *_ = load_data_as_list(name)
            return tr, va, te, ent, rel, edges
        except ImportError:
            if mode == "false":
                raise
            print(f"[train] intelligraphs not installed: using synthetic {name}-shaped graphs")
    nE, nR, lo, hi, _ = DATASET_SHAPES[name]
    rng = np.random.default_rng(seed)

    def make(n):
        sizes = rng.integers(lo, hi + 1, n)
        return [[(int(rng.integers(nE)), int(rng.integers(nR)), int(rng.integers(nE))) for _ in range(k)] for k in sizes]

    tr = make(int(config.get("synthetic_train_graphs", 4096)))
    va = make(int(config.get("synthetic_val_graphs", 256)))
    te = make(int(config.get("synthetic_val_graphs", 256)))
    e2i = {f"e{i}": i for i in range(nE)}
    r2i = {f"r{i}": i for i in range(nR)}
    return tr, va, te, (e2i, {v: k for k, v in e2i.items()}), (r2i, {v: k for k, v in r2i.items()}), (lo, hi)


class BatchLoader:
    """Rank-sharded, vectorised replacement of DataLoader(GraphSeqDataset) for the train step.

    The whole split is tensorised ONCE (reference rules: utils.py:102-108,131-146) into pinned host tensors
    `triples [G, T, 3]` and `seq [G, seq_len]`; an epoch is a (shuffled) row order of those tensors and a batch is a
    contiguous slice — no per-graph Python in the step loop (the reference's `num_workers=0` DataLoader builds every
    item with `random.sample` / `torch.tensor`, which caps real-data training two orders below the GPU step).
    Global batch g covers rows [g*B*W, (g+1)*B*W) of the epoch order; rank r owns the r-th slice of B graphs
    (drop_last).  Every rank can see all lengths, so the GLOBAL non-PAD token count of a step needs no collective.
    Yields (triples, seq, n_tok_global, batch_global) with pinned host tensors in the reference's format.
    """

    def __init__(self, graphs, vocab, batch_size, rank=0, world=1, shuffle=False, drop_last=True, permute=False,
                 seed=0, triple_order="keep", i2e=None, i2r=None, device=None):
        if triple_order != "keep":       # reference: GraphSeqDataset canonicalises every graph (utils.py:96-99,115)
            graphs = [canonicalize(g, i2e, i2r, triple_order) for g in graphs]
        self.graphs = graphs             # (posterior_bits iterates a GraphSeqDataset over the same canonical graphs)
        self.v, self.B, self.rank, self.world = vocab, batch_size, rank, world
        self.shuffle, self.drop_last, self.permute, self.epoch, self.seed = shuffle, drop_last, permute, 0, seed
        self.G = len(graphs)
        self.n = np.fromiter((len(g) for g in graphs), dtype=np.int64, count=self.G)
        self.lens = 3 * self.n + 1
        if not vocab["use_padding"] and self.G and not (self.n == self.n[0]).all():
            raise ValueError("without padding all graphs need the same number of triples (reference: default collate)")
        # one vectorised pass over the split; chunked so the temporary Python list stays small
        tri, seq = [], []
        for c0 in range(0, self.G, 8192):
            t_, s_ = build_batch(graphs[c0:c0 + 8192], special_tokens=SPECIAL, ent_base=vocab["ENT_BASE"],
                                 rel_base=vocab["REL_BASE"], seq_len=vocab["seq_len"], max_triples=vocab["max_edges"],
                                 use_padding=vocab["use_padding"], pad_eid=vocab["pad_eid"], pad_rid=vocab["pad_rid"])
            tri.append(t_)
            seq.append(s_)
        pin = torch.cuda.is_available()
        self.tri = torch.cat(tri) if tri else torch.zeros(0, 1, 3, dtype=torch.int64)
        self.seq = torch.cat(seq) if seq else torch.zeros(0, vocab["seq_len"], dtype=torch.int64)
        self.tri_e, self.seq_e = torch.empty_like(self.tri), torch.empty_like(self.seq)     # this epoch's row order
        if pin:
            self.tri, self.seq = self.tri.pin_memory(), self.seq.pin_memory()
            self.tri_e, self.seq_e = self.tri_e.pin_memory(), self.seq_e.pin_memory()
        # ON-DEVICE batch assembly: the tensorised split lives in HBM (wd-articles: ~10 KB per graph); an
        # epoch's order / triple permutation is applied there and a batch is a slice of device memory — no per-step
        # host->device copy of tokens at all.  The packed layout needs only the per-graph lengths, which the host has.
        self.device = torch.device(device) if device is not None else None
        if self.device is not None:
            self.tri_d, self.seq_d = self.tri.to(self.device), self.seq.to(self.device)

    def __len__(self):
        per = self.B * self.world
        return self.G // per if self.drop_last else math.ceil(self.G / per)

    def __iter__(self):
        self.epoch += 1
        tri, seq, lens = self.tri, self.seq, self.lens
        tri_d, seq_d = (self.tri_d, self.seq_d) if self.device is not None else (None, None)
        if self.shuffle or (self.permute and not self.v["use_padding"]):
            rng = np.random.default_rng(self.seed + self.epoch)
            order = torch.from_numpy(rng.permutation(self.G) if self.shuffle else np.arange(self.G))
            lens = self.lens[order.numpy()]
            T = tri.shape[1]
            perm_tri = self.permute and not self.v["use_padding"] and T > 1                 # reference: utils.py:133-134
            p = torch.from_numpy(np.argsort(rng.random((self.G, T)), axis=1)) if perm_tri else None   # a random order per graph
            if self.device is None:
                if torch.cuda.is_available():
                    torch.cuda.synchronize()     # last epoch's async H2D copies read tri_e / seq_e: finish them before rewriting
                torch.index_select(self.tri, 0, order, out=self.tri_e)
                torch.index_select(self.seq, 0, order, out=self.seq_e)
                tri, seq = self.tri_e, self.seq_e
                if perm_tri:
                    tri.copy_(torch.gather(tri, 1, p[:, :, None].expand(-1, -1, 3)))
                    tok = seq[:, 1:1 + 3 * T].reshape(self.G, T, 3)
                    seq[:, 1:1 + 3 * T] = torch.gather(tok, 1, p[:, :, None].expand(-1, -1, 3)).reshape(self.G, 3 * T)
            else:                                # the same order / permutation, applied to the device-resident copy
                order_d = order.to(self.device)
                tri_d, seq_d = self.tri_d.index_select(0, order_d), self.seq_d.index_select(0, order_d)
                if perm_tri:
                    p_d = p.to(self.device)[:, :, None].expand(-1, -1, 3)
                    tri_d = torch.gather(tri_d, 1, p_d)
                    tok_d = seq_d[:, 1:1 + 3 * T].reshape(self.G, T, 3)
                    seq_d[:, 1:1 + 3 * T] = torch.gather(tok_d, 1, p_d).reshape(self.G, 3 * T)
        per = self.B * self.world
        for g in range(len(self)):
            lo = g * per
            hi = min(lo + per, self.G)
            a, b = lo + self.rank * self.B, min(lo + (self.rank + 1) * self.B, hi)
            if b <= a:
                continue
            if self.device is not None:
                from ark_b200.layout import pack_layout_from_lens
                yield (tri_d[a:b], seq_d[a:b], int(lens[lo:hi].sum()), hi - lo,
                       pack_layout_from_lens(lens[a:b]).to(self.device))
            else:
                yield tri[a:b], seq[a:b], int(lens[lo:hi].sum()), hi - lo


# --------------------------------------------------------------------------------------------- loops
def train_epoch(model, dataloader, optimizer, config, device, b=1.0, eps_fn=None):
    """One epoch of ELBO steps (reference: ablation_study.py:31-88, SAIL branch :59-81).
    Returns (avg_loss, avg_recon, avg_kl, avg_entity_loss) like the reference."""
    mt = config.get("model_type", "ARK")
    if mt not in ("SAIL", "ARK", "t-SAIL", "t-ARK"):
        raise NotImplementedError(f"unknown model_type {mt!r}")
    model.train()
    eng = model.engine()
    eng.stats.zero_()
    fused = hasattr(optimizer, "sync_from_engine")        # FusedAdam: one fused call per step, Adam overlapped with backward
    # fixed-size datasets (syn-*) repeat one batch layout: replay the captured CUDA graph(s) instead of ~150 launches
    replay = fused and mt == "SAIL" and not config.get("use_padding", False) and bool(config.get("cuda_graph", True))
    for i, batch in enumerate(dataloader):
        triples, seq = batch[0], batch[1]
        ntg = batch[2] if len(batch) > 2 else None
        bg = batch[3] if len(batch) > 3 else None
        lay = batch[4] if len(batch) > 4 else None          # device-resident loader: layout from host-side lengths
        eps = None if eps_fn is None else eps_fn(i)
        if fused:
            lr = float(optimizer.param_groups[0]["lr"])
            if mt in ("ARK", "t-ARK"):       # decoder-only: loss = CE, KL = 0 (reference train.py:42-58)
                model.ce_step(seq, layout=lay, lr=lr, n_tok_global=ntg)
            else:
                model.elbo_step(triples, seq, b, eps=eps, layout=lay, lr=lr, n_tok_global=ntg, batch_global=bg, graph=replay)
            continue
        optimizer.zero_grad()
        if mt in ("ARK", "t-ARK"):
            model.ce_backward(seq, n_tok_global=ntg)
        else:
            model.elbo_backward(triples, seq, b, eps=eps, n_tok_global=ntg, batch_global=bg)
        optimizer.step()
    if fused:
        optimizer.sync_from_engine()
    loss, ce, kl = eng.read_stats(b)            # one device->host read for the whole epoch
    return loss, ce, kl, 0.0


@torch.no_grad()
def validate(model, dataloader, config, device, b=1.0):
    """Validation loss terms (reference: ablation_study.py:92-187, loss part), forward only."""
    model.eval()
    eng = model.engine()
    acc, n = torch.zeros(2, device=eng.device), 0
    for batch in dataloader:
        lay = model._make_layout(batch[0], batch[1])
        triples, seq = batch[0].to(eng.device), batch[1].to(eng.device)
        eps = torch.randn(triples.shape[0], config["d_latent"], device=eng.device) if eng.has_enc else None
        acc += eng.eval_step(triples.contiguous(), seq.contiguous(), lay, eps, b)
        n += 1
    ce, kl = (acc / max(n, 1)).tolist()
    return ce + b * kl, ce, kl


@torch.no_grad()
def posterior_compression(model, dataset, config, device):
    """Posterior compression bits of `sample_frac` of a split (reference: validate(compute_compression=True),
    ablation_study.py:151-186 -> SAIL/ARK.posterior_bits, models.py:218-260,488-520)."""
    model.eval()
    return model.posterior_bits(dataset, device, pad_id=SPECIAL["PAD"], sample_frac=float(config.get("sample_frac", 0.1)),
                                desc="Posterior compression")


def cosine_lr(base, epoch, t_max, eta_min):
    return eta_min + (base - eta_min) * (1 + math.cos(math.pi * epoch / t_max)) / 2


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=str, required=True, help="Path to config file")
    ap.add_argument("--wandb-project", type=str, default="submission", help="Weights & Biases project name")
    ap.add_argument("--wandb-entity", type=str, default=None, help="Weights & Biases entity")
    ap.add_argument("--checkpoint-dir", type=str, default="checkpoints", help="Directory to save checkpoints")
    ap.add_argument("--max-epochs", type=int, default=None, help="(new) stop early, for smoke runs")
    args = ap.parse_args(argv)

    with open(args.config) as f:
        config = yaml.safe_load(f)
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
    if not torch.cuda.is_available():
        raise RuntimeError("kgvae.experiments.train runs on B200 GPUs only (ark_b200 has no CPU path)")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    group = None
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=device)
        group = torch.distributed.group.WORLD

    log = _Log(args.wandb_project, args.wandb_entity or os.getenv("WANDB_ENTITY"), config,
               config.get("experiment_name", "ARK_experiment"), rank)
    config.update(log.config)                                   # sweep overrides (reference train.py:273)
    if world > 1:       # wandb runs on rank 0 only: every rank must train with rank 0's effective hyper-parameters
        box = [config]
        torch.distributed.broadcast_object_list(box, src=0)
        config = box[0]
    config["learning_rate"] = float(config.get("learning_rate", 1e-3))
    run_dir = os.path.join(args.checkpoint_dir, str(log.run.id))
    if rank == 0:
        os.makedirs(run_dir, exist_ok=True)
        with open(os.path.join(run_dir, "effective_config.yaml"), "w") as f:
            yaml.safe_dump(config, f)
    best_comp_bits = 1e12
    log.log({"objective": best_comp_bits})
    if config.get("use_test_for_final_eval", False) and rank == 0:
        warnings.warn("Test set evaluation ENABLED! Only use for final evaluation, NOT for hyperparameter tuning!",
                      UserWarning, stacklevel=2)

    model_type = config.get("model_type", "ARK")
    if model_type not in ("SAIL", "ARK", "t-SAIL", "t-ARK"):
        raise NotImplementedError(f"Unknown model_type: {model_type}")      # reference: models.py:197,390

    train_g, val_g, test_g, (e2i, i2e), (r2i, i2r), (min_edges, max_edges) = load_graphs(config)
    n_ent, n_rel = len(e2i), len(r2i)
    use_padding = config.get("use_padding", config["dataset"].startswith("wd-"))
    pad_eid = pad_rid = None
    if use_padding:                                            # reference: ablation_study.py:436-447
        pad_eid, pad_rid = n_ent, n_rel
        n_ent, n_rel = n_ent + 1, n_rel + 1
    ent_base = 3
    rel_base = ent_base + n_ent
    vocab = {"ENT_BASE": ent_base, "REL_BASE": rel_base, "seq_len": 3 * max_edges + 2, "max_edges": max_edges,
             "use_padding": use_padding, "pad_eid": pad_eid, "pad_rid": pad_rid}
    config.update({"n_entities": n_ent, "n_relations": n_rel, "pad_eid": pad_eid, "pad_rid": pad_rid,
                   "seq_len": vocab["seq_len"], "vocab_size": rel_base + n_rel, "special_tokens": SPECIAL,
                   "ENT_BASE": ent_base, "REL_BASE": rel_base})

    B = config["batch_size"]
    order = config.get("triple_order", "keep")
    train_loader = BatchLoader(train_g, vocab, B, rank, world, shuffle=config["shuffle_train"], drop_last=True,
                               permute=config.get("permute_triples", False), triple_order=order, i2e=i2e, i2r=i2r,
                               device=device if config.get("device_resident_data", True) else None)
    val_loader = BatchLoader(val_g, vocab, B, 0, 1, drop_last=False, triple_order=order, i2e=i2e, i2r=i2r)
    test_loader = BatchLoader(test_g, vocab, B, 0, 1, drop_last=False, triple_order=order, i2e=i2e, i2r=i2r)

    def seq_dataset(loader):     # what the reference hands to posterior_bits: dataloader.dataset (ablation_study.py:151-157)
        return GraphSeqDataset(loader.graphs, i2e, i2r, triple_order="keep", permute=False, use_padding=use_padding,
                               pad_eid=pad_eid, pad_rid=pad_rid, max_triples=max_edges, special_tokens=SPECIAL,
                               ent_base=ent_base, rel_base=rel_base, seq_len=vocab["seq_len"])
    if rank == 0:
        print(f"Dataset: {config['dataset']}  Entities: {n_ent}, Relations: {n_rel}")
        print(f"Train batches: {len(train_loader)}, Val batches: {len(val_loader)}  world={world}")

    torch.manual_seed(0)                                       # identical initial weights on every rank
    model = (ARK if model_type in ("ARK", "t-ARK") else SAIL)(config).to(device)
    # ... but independent noise per rank: eps (torch's CUDA generator) and the Philox dropout stream.  With one common
    # seed every rank would draw the SAME eps / masks for its different shard, which no single-process run does
    torch.manual_seed(1 + rank)
    optimizer = FusedAdam(model, lr=config["learning_rate"], dist_group=group,
                          bucket_mb=float(config.get("ddp_bucket_mb", 32)), seed=rank)
    scheduler = None
    if config.get("lr_scheduler", False):
        scheduler = torch.optim.lr_scheduler.CosineAnnealingLR(optimizer, T_max=config["num_epochs"],
                                                               eta_min=config.get("eta_min", 1e-6))
    best_val = float("inf")
    n_epochs = config["num_epochs"] if args.max_epochs is None else min(config["num_epochs"], args.max_epochs)
    for epoch in range(n_epochs):
        b = config["beta0"] + (config["beta1"] - config["beta0"]) * epoch / config["num_epochs"]
        t0 = time.perf_counter()
        tr = train_epoch(model, train_loader, optimizer, config, device, b)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        n_tri = sum(len(g) for g in train_g[:len(train_loader) * B * world])
        va = validate(model, val_loader, config, device, b) if rank == 0 else (0.0, 0.0, 0.0)
        if rank == 0:
            print(f"\nEpoch {epoch + 1}/{config['num_epochs']}  loss {tr[0]:.4f} ce {tr[1]:.4f} kl {tr[2]:.4f}  "
                  f"val {va[0]:.4f}  {n_tri / dt:,.0f} triples/s")
            if (epoch + 1) % int(config.get("compression_log_every", 5)) == 0:     # ablation_study.py:595-621
                stats = posterior_compression(model, seq_dataset(val_loader), config, device)
                log.log({"val/compression_bits": stats["avg_total_bits"], "val/compression_kl_bits": stats["avg_kl_bits"],
                         "val/compression_edge_bits": stats["avg_ar_bits"], "val/compression_entity_bits": stats["avg_ar_bits"]})
                if math.isfinite(stats["avg_total_bits"]) and stats["avg_total_bits"] < best_comp_bits:
                    best_comp_bits = stats["avg_total_bits"]
            log.log({"objective": best_comp_bits})
            log.log({"epoch": epoch + 1, "train/loss": tr[0], "train/reconstruction_loss": tr[1], "train/kl_loss": tr[2],
                     "val/loss": va[0], "val/reconstruction_loss": va[1], "val/kl_loss": va[2], "beta": b,
                     "learning_rate": optimizer.param_groups[0]["lr"], "triples_per_sec": n_tri / dt,
                     "step_ms": 1e3 * dt / max(len(train_loader), 1)})
        if scheduler is not None:
            scheduler.step()
        if world > 1:       # every rank: owners broadcast their slices of the (sharded) Adam state before rank 0 reads it
            model.engine().gather_adam_state()
        if rank == 0:
            ckpt = {"epoch": epoch + 1, "model_state_dict": {k: v.detach().cpu() for k, v in model.state_dict().items()},
                    "optimizer_state_dict": optimizer.state_dict(),
                    "scheduler_state_dict": scheduler.state_dict() if scheduler else None, "val_loss": va[0],
                    "config": config, "vocabs": {"e2i": e2i, "i2e": i2e, "r2i": r2i, "i2r": i2r},
                    "dataset_meta": {"dataset": config["dataset"], "n_entities": len(i2e), "n_relations": len(i2r)}}
            tag = f"{config['dataset']}_{model_type}"
            if va[0] < best_val:
                best_val = va[0]
                torch.save(ckpt, os.path.join(run_dir, f"{tag}_best_model.pt"), _use_new_zipfile_serialization=False)
            if (epoch + 1) % int(config.get("save_every", 10)) == 0:
                torch.save(ckpt, os.path.join(run_dir, f"{tag}_checkpoint_epoch_{epoch + 1}.pt"),
                           _use_new_zipfile_serialization=False)
        if world > 1:       # rank 0 alone validated / logged / saved: the others wait HERE, on the host, not inside the
            torch.distributed.barrier()     # next step's exchange kernel (its cross-rank spin is bounded: csrc/dp_reduce.cu)
    if rank == 0:     # final evaluation (reference final_validation, ablation_study.py:190-346, minus the intelligraphs parts)
        final = {}
        splits = [("final_val", val_loader)] + ([("final_test", test_loader)] if config.get("use_test_for_final_eval") else [])
        for tag, loader in splits:
            fl = validate(model, loader, config, device, 1.0)
            stats = posterior_compression(model, seq_dataset(loader), config, device)
            final.update({f"{tag}/loss": fl[0], f"{tag}/reconstruction_loss": fl[1], f"{tag}/kl_loss": fl[2],
                          f"{tag}/compression_bits": stats["avg_total_bits"], f"{tag}/compression_kl_bits": stats["avg_kl_bits"],
                          f"{tag}/compression_ar_bits": stats["avg_ar_bits"]})
        log.log(final)
    log.finish()
    if world > 1:
        model.engine().release_graphs()     # (captured NCCL kernels must be gone before the group is destroyed)
        torch.distributed.destroy_process_group()
    return model


if __name__ == "__main__":
    main()
