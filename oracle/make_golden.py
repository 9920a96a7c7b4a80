"""Generate tests/golden/* by RUNNING THE UNMODIFIED REFERENCE in the build container.

TEST INFRASTRUCTURE.  Usage (from the repository root, build container only):

    python oracle/make_golden.py

It imports /root/reference's ``kgvae`` (through oracle/ref_loader.py: an ``intelligraphs`` stub
plus sys.path), executes the reference's own SAIL model / loss / Adam / train_epoch /
beam search / token indexing on small seeded inputs and stores inputs + outputs as fixtures.
/root/reference does not exist on the GPU box, so the fixtures — not this script — are what the
tests read.  Nothing here is imported by the product path.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
# make sure this repository's own `kgvae` is NOT importable ahead of the reference's
sys.path = [p for p in sys.path if os.path.abspath(p or ".") != os.path.dirname(HERE)]

import ref_loader  # noqa: E402

ref_models, ref_utils = ref_loader.load_reference()

import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
os.makedirs(OUT, exist_ok=True)
torch.set_num_threads(1)


def layout_from_reference_rules(n_ent, n_rel, max_edges, use_padding):
    # executed copy of the arithmetic in ablation_study.py:436-454 (no function to import there)
    num_entities, num_relations = n_ent, n_rel
    if use_padding:
        PAD_EID, PAD_RID = num_entities, num_relations
        num_entities += 1
        num_relations += 1
    else:
        PAD_EID = PAD_RID = None
    special = {"PAD": 0, "BOS": 1, "EOS": 2}
    ENT_BASE = 3
    REL_BASE = ENT_BASE + num_entities
    return {"n_entities": num_entities, "n_relations": num_relations, "pad_eid": PAD_EID,
            "pad_rid": PAD_RID, "special_tokens": special, "ENT_BASE": ENT_BASE, "REL_BASE": REL_BASE,
            "vocab_size": REL_BASE + num_relations, "seq_len": 1 + max_edges * 3 + 1,
            "max_edges": max_edges}


def random_graphs(rng, n_graphs, n_ent, n_rel, lo, hi):
    out = []
    for _ in range(n_graphs):
        n = int(rng.integers(lo, hi + 1))
        out.append([(int(rng.integers(n_ent)), int(rng.integers(n_rel)), int(rng.integers(n_ent)))
                    for _ in range(n)])
    return out


def dataset_batch(graphs, lay, use_padding):
    ds = ref_utils.GraphSeqDataset(
        graphs=graphs, i2e=None, i2r=None, triple_order="keep", permute=False,
        use_padding=use_padding, pad_eid=lay["pad_eid"], pad_rid=lay["pad_rid"],
        max_triples=lay["max_edges"], special_tokens=lay["special_tokens"],
        ent_base=lay["ENT_BASE"], rel_base=lay["REL_BASE"], seq_len=lay["seq_len"])
    loader = torch.utils.data.DataLoader(ds, batch_size=len(graphs), shuffle=False)
    (triples, seq), = list(loader)
    return triples, seq


def sail_case(name, *, n_ent, n_rel, lo, hi, use_padding, d, dz, nl, B, seed, beta,
              tie=True, logv_bias=None, lengths=None):
    rng = np.random.default_rng(seed)
    lay = layout_from_reference_rules(n_ent, n_rel, hi, use_padding)
    graphs = random_graphs(rng, B, n_ent, n_rel, lo, hi)
    if lengths is not None:
        graphs = [g[:n] + random_graphs(rng, 1, n_ent, n_rel, max(n - len(g), 0), max(n - len(g), 0))[0]
                  for g, n in zip(graphs, lengths)]
    triples, seq = dataset_batch(graphs, lay, use_padding)
    cfg = dict(lay, model_type="SAIL", d_model=d, d_latent=dz, n_heads=2, n_layers=nl,
               dec_dropout=0.0, tie_weights=tie)
    torch.manual_seed(seed)
    model = ref_models.SAIL(cfg)
    if logv_bias is not None:
        with torch.no_grad():
            model.enc.logv.bias.copy_(torch.linspace(-logv_bias, logv_bias, dz))
    model.train()
    torch.manual_seed(1000 + seed)
    eps = torch.randn(B, dz)
    torch.manual_seed(1000 + seed)          # SAIL's eps is the first draw
    logits, mu, logv = model(triples, seq[:, :-1])
    z_check, _, _ = (mu + eps * torch.exp(0.5 * logv)), None, None
    ce = F.cross_entropy(logits.reshape(-1, logits.size(-1)), seq[:, 1:].reshape(-1), ignore_index=0)
    kl = model.kl_mean(mu, logv)
    loss = ce + beta * kl
    loss.backward()
    # confirm eps really is what the reference drew
    torch.manual_seed(1000 + seed)
    z_ref, mu2, logv2 = model.enc(triples)
    assert torch.equal(z_ref, mu2 + eps * torch.exp(0.5 * logv2)), "eps replay mismatch"
    arrays = {"triples": triples.numpy(), "seq": seq.numpy(), "eps": eps.numpy(),
              "beta": np.float64(beta), "z": z_check.detach().numpy(), "mu": mu.detach().numpy(),
              "logv": logv.detach().numpy(), "logits": logits.detach().numpy(),
              "ce": np.float64(ce.item()), "kl": np.float64(kl.item()), "loss": np.float64(loss.item())}
    sd = model.state_dict()
    for k, v in sd.items():
        arrays["param::" + k] = v.detach().numpy().copy()
    for k, p in model.named_parameters():
        arrays["grad::" + k] = p.grad.detach().numpy().copy()
    # two Adam steps of the reference trainer body (ablation_study.py:43,59-76)
    opt = torch.optim.Adam(model.parameters(), lr=1e-2)
    model.zero_grad()
    step_losses = []
    for s in range(2):
        opt.zero_grad()
        torch.manual_seed(2000 + seed + s)
        lg, m_, lv_ = model(triples, seq[:, :-1])
        c_ = F.cross_entropy(lg.reshape(-1, lg.size(-1)), seq[:, 1:].reshape(-1), ignore_index=0)
        k_ = model.kl_mean(m_, lv_)
        l_ = c_ + beta * k_
        l_.backward()
        opt.step()
        step_losses.append([l_.item(), c_.item(), k_.item()])
        torch.manual_seed(2000 + seed + s)
        arrays[f"adam_eps{s}"] = torch.randn(B, dz).numpy()
    arrays["adam_losses"] = np.asarray(step_losses)
    for k, v in model.state_dict().items():
        arrays["adam_param::" + k] = v.detach().numpy().copy()
    # batch-shared beam search from fixed latents (models.py:262-300), eval mode
    zg = torch.from_numpy(np.random.default_rng(seed + 7).standard_normal((3, dz)).astype(np.float32))
    decoded = model.decode_latent(zg, lay["seq_len"], lay["special_tokens"], ref_utils.seq_to_triples,
                                  lay["ENT_BASE"], lay["REL_BASE"], beam=3)
    model.eval()
    with torch.no_grad():
        arrays["beam_z"] = zg.numpy()
        arrays["eval_logits_prefix4"] = model.dec(zg, seq[:3, :4]).numpy()
    meta = {"cfg": cfg, "graphs": graphs, "beam_decoded": [[list(t) for t in g] for g in decoded],
            "beam": 3, "adam_lr": 1e-2}
    np.savez_compressed(os.path.join(OUT, f"sail_{name}.npz"), **arrays)
    with open(os.path.join(OUT, f"sail_{name}.json"), "w") as f:
        json.dump(meta, f)
    print(f"[golden] sail_{name}: loss={loss.item():.6f} ce={ce.item():.6f} kl={kl.item():.6f} "
          f"V={lay['vocab_size']} L={lay['seq_len'] - 1}")


def utils_golden():
    rng = np.random.default_rng(5)
    rec = []
    for (n_ent, n_rel, lo, hi, pad) in [(7, 2, 3, 3, False), (19, 4, 1, 5, True), (5, 1, 0, 2, True)]:
        lay = layout_from_reference_rules(n_ent, n_rel, hi, pad)
        graphs = random_graphs(rng, 4, n_ent, n_rel, lo, hi)
        seqs = [ref_utils.triples_to_seq(g, lay["special_tokens"], lay["ENT_BASE"], lay["REL_BASE"],
                                         lay["seq_len"]).tolist() for g in graphs]
        back = [[list(t) for t in ref_utils.seq_to_triples(torch.tensor(s), lay["special_tokens"],
                                                           lay["ENT_BASE"], lay["REL_BASE"])] for s in seqs]
        # seq_to_triples on sequences WITHOUT an EOS (generation may emit those) and truncated ones
        odd = [s[:-1] for s in seqs] + [[1] + s[1:5] for s in seqs]
        odd_back = [[list(t) for t in ref_utils.seq_to_triples(s, lay["special_tokens"], lay["ENT_BASE"],
                                                               lay["REL_BASE"])] for s in odd]
        tri, seq = dataset_batch(graphs, lay, pad)
        rec.append({"n_ent": n_ent, "n_rel": n_rel, "max_edges": hi, "use_padding": pad, "layout": lay,
                    "graphs": graphs, "seqs": seqs, "seq_to_triples": back, "odd": odd,
                    "odd_back": odd_back, "batch_triples": tri.tolist(), "batch_seq": seq.tolist()})
    with open(os.path.join(OUT, "utils_indexing.json"), "w") as f:
        json.dump(rec, f)
    print(f"[golden] utils_indexing: {len(rec)} layouts")


def train_epoch_golden():
    """Reference ablation_study.train_epoch on 3 batches (dec_dropout=0 so it is deterministic)."""
    import importlib
    abl = importlib.import_module("kgvae.experiments.ablation_study")
    rng = np.random.default_rng(11)
    lay = layout_from_reference_rules(17, 3, 4, True)
    cfg = dict(lay, model_type="SAIL", d_model=16, d_latent=5, n_heads=2, n_layers=2,
               dec_dropout=0.0, tie_weights=True)
    batches = []
    for _ in range(3):
        graphs = random_graphs(rng, 4, 17, 3, 1, 4)
        batches.append(dataset_batch(graphs, lay, True))
    torch.manual_seed(3)
    model = ref_models.SAIL(cfg)
    arrays = {}
    for k, v in model.state_dict().items():
        arrays["param::" + k] = v.detach().numpy().copy()
    opt = torch.optim.Adam(model.parameters(), lr=5e-3)
    torch.manual_seed(99)
    res = abl.train_epoch(model, batches, opt, cfg, torch.device("cpu"), 0.37)
    torch.manual_seed(99)
    for i, (t, s) in enumerate(batches):
        arrays[f"triples{i}"], arrays[f"seq{i}"] = t.numpy(), s.numpy()
        arrays[f"eps{i}"] = torch.randn(4, 5).numpy()
    arrays["result"] = np.asarray(res[:3], dtype=np.float64)
    for k, v in model.state_dict().items():
        arrays["after::" + k] = v.detach().numpy().copy()
    np.savez_compressed(os.path.join(OUT, "train_epoch.npz"), **arrays)
    with open(os.path.join(OUT, "train_epoch.json"), "w") as f:
        json.dump({"cfg": cfg, "lr": 5e-3, "beta": 0.37}, f)
    print(f"[golden] train_epoch: {res[:3]}")


def ark_case(name, *, n_ent, n_rel, lo, hi, use_padding, d, nl, B, seed, tie=True, model_type="ARK", heads=2):
    """Decoder-only ARK / t-ARK (models.py:323-405) with the CE-only step of train.py:42-58."""
    rng = np.random.default_rng(seed)
    lay = layout_from_reference_rules(n_ent, n_rel, hi, use_padding)
    graphs = random_graphs(rng, B, n_ent, n_rel, lo, hi)
    triples, seq = dataset_batch(graphs, lay, use_padding)
    cfg = dict(lay, model_type=model_type, d_model=d, d_latent=4, n_heads=heads, n_layers=nl, dec_dropout=0.0, tie_weights=tie)
    torch.manual_seed(seed)
    model = ref_models.ARK(cfg)
    if model_type == "t-ARK":
        with torch.no_grad():   # LayerNorm affine / biases away from (1, 0) so their gradients are exercised
            for n_, p_ in model.named_parameters():
                if "norm" in n_ or n_.endswith("bias"):
                    p_.add_(0.1 * torch.randn_like(p_))
    model.train()
    logits = model(seq[:, :-1])
    ce = F.cross_entropy(logits.reshape(-1, logits.size(-1)), seq[:, 1:].reshape(-1), ignore_index=0)
    ce.backward()
    arrays = {"triples": triples.numpy(), "seq": seq.numpy(), "logits": logits.detach().numpy(),
              "ce": np.float64(ce.item())}
    for k, v in model.state_dict().items():
        arrays["param::" + k] = v.detach().numpy().copy()
    for k, p in model.named_parameters():
        arrays["grad::" + k] = p.grad.detach().numpy().copy()
    opt = torch.optim.Adam(model.parameters(), lr=1e-2)
    losses = []
    for s in range(2):
        opt.zero_grad()
        lg = model(triples, seq[:, :-1])                      # the (triples, seq) call form, models.py:395-405
        c_ = F.cross_entropy(lg.reshape(-1, lg.size(-1)), seq[:, 1:].reshape(-1), ignore_index=0)
        c_.backward()
        opt.step()
        losses.append(c_.item())
    arrays["adam_losses"] = np.asarray(losses)
    for k, v in model.state_dict().items():
        arrays["adam_param::" + k] = v.detach().numpy().copy()
    model.eval()
    with torch.no_grad():
        gen = model.generate(lay["seq_len"], lay["special_tokens"], batch_size=3, sample=False)   # greedy
        arrays["greedy"] = gen.numpy()
        arrays["eval_logits_prefix5"] = model(seq[:2, :5]).numpy()
    if model_type == "t-ARK":      # keep the fixture small: the Adam'd weights are not needed for the Transformer case
        arrays = {k: v for k, v in arrays.items() if not k.startswith("adam_param::")}
    np.savez_compressed(os.path.join(OUT, f"ark_{name}.npz"), **arrays)
    with open(os.path.join(OUT, f"ark_{name}.json"), "w") as f:
        json.dump({"cfg": cfg, "adam_lr": 1e-2}, f)
    print(f"[golden] ark_{name}: ce={ce.item():.6f} V={lay['vocab_size']} L={lay['seq_len'] - 1}")


def tsail_case(name, *, n_ent, n_rel, lo, hi, use_padding, d, dz, heads, nl, B, seed, beta, lengths=None):
    """Transformer KG-VAE (models.py:66-114) with the ELBO step of ablation_study.py:59-75.  The reference's
    layers hard-code torch's default dropout 0.1; for an exact fixture every dropout probability of the
    instantiated reference modules is set to 0 (hyper-parameter change on the live modules, no code change)."""
    rng = np.random.default_rng(seed)
    lay = layout_from_reference_rules(n_ent, n_rel, hi, use_padding)
    graphs = random_graphs(rng, B, n_ent, n_rel, lo, hi)
    if lengths is not None:
        graphs = [g[:n] + random_graphs(rng, 1, n_ent, n_rel, max(n - len(g), 0), max(n - len(g), 0))[0]
                  for g, n in zip(graphs, lengths)]
    triples, seq = dataset_batch(graphs, lay, use_padding)
    cfg = dict(lay, model_type="t-SAIL", d_model=d, d_latent=dz, n_heads=heads, n_layers=nl, txf_dropout=0.0)
    torch.manual_seed(seed)
    model = ref_models.SAIL(cfg)
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
        if isinstance(m, torch.nn.MultiheadAttention):
            m.dropout = 0.0
    with torch.no_grad():      # LayerNorm affine / biases away from their (1, 0) init so their gradients are exercised
        for n_, p_ in model.named_parameters():
            if "norm" in n_ or n_.endswith("bias"):
                p_.add_(0.1 * torch.randn_like(p_))
    model.train()
    torch.manual_seed(1000 + seed)
    eps = torch.randn(B, dz)
    torch.manual_seed(1000 + seed)
    logits, mu, logv = model(triples, seq[:, :-1])
    ce = F.cross_entropy(logits.reshape(-1, logits.size(-1)), seq[:, 1:].reshape(-1), ignore_index=0)
    kl = model.kl_mean(mu, logv)
    loss = ce + beta * kl
    loss.backward()
    torch.manual_seed(1000 + seed)
    z_ref, mu2, logv2 = model.enc(triples)
    assert torch.equal(z_ref, mu2 + eps * torch.exp(0.5 * logv2)), "eps replay mismatch (dropout consumed RNG?)"
    arrays = {"triples": triples.numpy(), "seq": seq.numpy(), "eps": eps.numpy(), "beta": np.float64(beta),
              "mu": mu.detach().numpy(), "logv": logv.detach().numpy(), "logits": logits.detach().numpy(),
              "ce": np.float64(ce.item()), "kl": np.float64(kl.item()), "loss": np.float64(loss.item())}
    for k, v in model.state_dict().items():
        arrays["param::" + k] = v.detach().numpy().copy()
    for k, p_ in model.named_parameters():
        arrays["grad::" + k] = p_.grad.detach().numpy().copy()
    np.savez_compressed(os.path.join(OUT, f"tsail_{name}.npz"), **arrays)
    with open(os.path.join(OUT, f"tsail_{name}.json"), "w") as f:
        json.dump({"cfg": cfg}, f)
    print(f"[golden] tsail_{name}: loss={loss.item():.6f} ce={ce.item():.6f} kl={kl.item():.6f} "
          f"V={lay['vocab_size']} L={lay['seq_len'] - 1} params={sum(p_.numel() for p_ in model.parameters())}")


def tsail_golden():
    tsail_case("syn", n_ent=11, n_rel=3, lo=3, hi=3, use_padding=False, d=8, dz=6, heads=2, nl=2, B=5, seed=21, beta=0.5)
    tsail_case("wd", n_ent=23, n_rel=4, lo=1, hi=6, use_padding=True, d=8, dz=8, heads=2, nl=2, B=6, seed=22,
               beta=0.25, lengths=[6, 1, 3, 4, 2, 6])


def tark_golden():
    ark_case("t_wd", n_ent=23, n_rel=4, lo=1, hi=6, use_padding=True, d=16, nl=2, B=6, seed=13, model_type="t-ARK", heads=4)


def ark_golden():
    ark_case("syn", n_ent=11, n_rel=3, lo=3, hi=3, use_padding=False, d=16, nl=3, B=5, seed=11)
    ark_case("wd", n_ent=23, n_rel=4, lo=1, hi=6, use_padding=True, d=16, nl=2, B=6, seed=12)


def eval_golden():
    """Evaluation-time paths of the reference: `posterior_bits` / `bits_per_sequence`
    (models.py:202-260, 473-520) and SAMPLED `ARK.generate` (temperature / top-k / top-p, models.py:408-471).
    The reference samples with torch.multinomial on the CPU generator, which no CUDA run can replay; the fixture
    therefore records the FILTERED distributions the reference hands to torch.multinomial (what the filtering code
    computes) with multinomial replaced by argmax for the duration of the call, so the token path is deterministic."""
    arrays, meta = {}, {}
    # ---- SAIL posterior bits (padded, ragged)
    rng = np.random.default_rng(31)
    lay = layout_from_reference_rules(23, 4, 6, True)
    graphs = random_graphs(rng, 7, 23, 4, 1, 6)
    cfg = dict(lay, model_type="SAIL", d_model=32, d_latent=8, n_heads=2, n_layers=2, dec_dropout=0.0, tie_weights=True)
    ds = ref_utils.GraphSeqDataset(graphs=graphs, i2e=None, i2r=None, triple_order="keep", permute=False, use_padding=True,
                                   pad_eid=lay["pad_eid"], pad_rid=lay["pad_rid"], max_triples=lay["max_edges"],
                                   special_tokens=lay["special_tokens"], ent_base=lay["ENT_BASE"], rel_base=lay["REL_BASE"],
                                   seq_len=lay["seq_len"])
    torch.manual_seed(31)
    model = ref_models.SAIL(cfg).eval()
    torch.manual_seed(77)
    draws, real_randn_like = [], torch.randn_like

    def recording_randn_like(t, *a, **k):       # record the reference's own eps draws (models.py:63), one per graph
        e = real_randn_like(t, *a, **k)
        draws.append(e.detach().clone())
        return e
    torch.randn_like = recording_randn_like
    try:
        stats = model.posterior_bits(ds, torch.device("cpu"), pad_id=0, sample_frac=1.0)
    finally:
        torch.randn_like = real_randn_like
    assert len(draws) == len(ds)
    arrays["sail_eps"] = torch.cat(draws, 0).numpy()
    for k, v in model.state_dict().items():
        arrays["sail_param::" + k] = v.detach().numpy().copy()
    arrays["sail_ar_bits"] = np.asarray([r["ar_bits"] for r in stats["records"]])
    arrays["sail_kl_bits"] = np.asarray([r["kl_bits"] for r in stats["records"]])
    z1 = torch.from_numpy(arrays["sail_eps"][:1])
    arrays["sail_bits_seq0_z"] = np.float64(model.bits_per_sequence(ds[0][1], z1, 0))
    meta["sail"] = {"cfg": cfg, "graphs": graphs, "stats": {k: v for k, v in stats.items() if k != "records"}}
    # ---- ARK posterior bits + sampled generation
    lay2 = layout_from_reference_rules(23, 4, 5, True)
    graphs2 = random_graphs(rng, 5, 23, 4, 1, 5)
    cfg2 = dict(lay2, model_type="ARK", d_model=32, d_latent=4, n_heads=2, n_layers=2, dec_dropout=0.0, tie_weights=True)
    ds2 = ref_utils.GraphSeqDataset(graphs=graphs2, i2e=None, i2r=None, triple_order="keep", permute=False, use_padding=True,
                                    pad_eid=lay2["pad_eid"], pad_rid=lay2["pad_rid"], max_triples=lay2["max_edges"],
                                    special_tokens=lay2["special_tokens"], ent_base=lay2["ENT_BASE"], rel_base=lay2["REL_BASE"],
                                    seq_len=lay2["seq_len"])
    torch.manual_seed(32)
    ark = ref_models.ARK(cfg2).eval()
    st2 = ark.posterior_bits(ds2, torch.device("cpu"), pad_id=0, sample_frac=1.0)
    for k, v in ark.state_dict().items():
        arrays["ark_param::" + k] = v.detach().numpy().copy()
    arrays["ark_ar_bits"] = np.asarray([r["ar_bits"] for r in st2["records"]])
    meta["ark"] = {"cfg": cfg2, "graphs": graphs2, "stats": {k: v for k, v in st2.items() if k != "records"}}
    real_multinomial = torch.multinomial
    gens = []
    for gi, kw in enumerate([dict(temperature=1.0, top_p=0.9, top_k=0), dict(temperature=0.7, top_p=0.0, top_k=5),
                             dict(temperature=1.3, top_p=0.8, top_k=7), dict(temperature=1.0, top_p=0.0, top_k=0)]):
        seen = []

        def fake(probs, n, *a, **k):
            seen.append(probs.detach().clone().reshape(-1, probs.shape[-1]))
            return probs.argmax(dim=-1, keepdim=True)
        torch.multinomial = fake
        try:
            out = ark.generate(lay2["seq_len"], lay2["special_tokens"], batch_size=3, sample=True, **kw)
        finally:
            torch.multinomial = real_multinomial
        arrays[f"gen{gi}_seq"] = out.numpy()
        # top-p: one call per batch row (sorted probabilities); otherwise one call per step ([B, V])
        arrays[f"gen{gi}_probs"] = torch.cat(seen, 0).numpy()
        gens.append(dict(kw, n_calls=len(seen)))
    meta["gen"] = gens
    np.savez_compressed(os.path.join(OUT, "eval_bits.npz"), **arrays)
    with open(os.path.join(OUT, "eval_bits.json"), "w") as f:
        json.dump(meta, f)
    print(f"[golden] eval_bits: sail avg_total={stats['avg_total_bits']:.4f} ark avg_total={st2['avg_total_bits']:.4f} "
          f"gen calls {[g['n_calls'] for g in gens]}")


if __name__ == "__main__":
    if "--eval-only" in sys.argv:
        eval_golden()
        sys.exit(0)
    if "--ark-only" in sys.argv:
        ark_golden()
        sys.exit(0)
    if "--tark-only" in sys.argv:
        tark_golden()
        sys.exit(0)
    if "--tsail-only" in sys.argv:
        tsail_golden()
        sys.exit(0)
    utils_golden()
    sail_case("syn", n_ent=11, n_rel=3, lo=3, hi=3, use_padding=False, d=16, dz=6, nl=3, B=5, seed=1, beta=0.5)
    sail_case("wd", n_ent=23, n_rel=4, lo=1, hi=6, use_padding=True, d=16, dz=8, nl=2, B=6, seed=2,
              beta=0.25, lengths=[6, 1, 3, 4, 2, 6])
    sail_case("wd_clamp", n_ent=23, n_rel=4, lo=1, hi=5, use_padding=True, d=8, dz=12, nl=1, B=4, seed=3,
              beta=1.0, logv_bias=14.0)
    sail_case("untied", n_ent=9, n_rel=2, lo=2, hi=2, use_padding=False, d=8, dz=4, nl=2, B=3, seed=4,
              beta=0.1, tie=False)
    train_epoch_golden()
    ark_golden()
    tsail_golden()
    tark_golden()
    eval_golden()
