"""Pins oracle/torch_cpu_port.py (the CPU baseline that bench.py times) to golden outputs of the reference."""
import numpy as np
import torch

from conftest import load_sail_golden
from oracle.torch_cpu_port import CpuSail, train_steps


def _load(cfg, params):
    m = CpuSail(cfg)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in params.items()}, strict=True)
    return m


def test_port_loss_and_grads_match_reference(sail_golden):
    name, arr, meta, params, grads = sail_golden
    m = _load(meta["cfg"], params).train()
    loss, ce, kl = m.elbo(torch.from_numpy(arr["triples"]), torch.from_numpy(arr["seq"]), float(arr["beta"]),
                          torch.from_numpy(arr["eps"]))
    loss.backward()
    np.testing.assert_allclose([loss.item(), ce.item(), kl.item()],
                               [float(arr["loss"]), float(arr["ce"]), float(arr["kl"])], rtol=1e-5)
    for k, p in m.named_parameters():
        np.testing.assert_allclose(p.grad.numpy(), grads[k], rtol=1e-4, atol=1e-6)


def test_port_two_adam_steps_match_reference():
    arr, meta, params, _ = load_sail_golden("wd")
    m = _load(meta["cfg"], params)
    opt = torch.optim.Adam(m.parameters(), lr=meta["adam_lr"])
    b = (torch.from_numpy(arr["triples"]), torch.from_numpy(arr["seq"]))
    rec = train_steps(m, opt, [b, b], float(arr["beta"]), [torch.from_numpy(arr["adam_eps0"]), torch.from_numpy(arr["adam_eps1"])])
    np.testing.assert_allclose(np.asarray(rec), arr["adam_losses"], rtol=1e-5)
    for k, v in m.state_dict().items():
        np.testing.assert_allclose(v.numpy(), arr["adam_param::" + k], rtol=1e-4, atol=1e-6)


def test_one_pass_posterior_bits_equal_the_reference_prefix_loop():
    """The reference computes AR bits with an O(L^2) loop over growing prefixes (models.py:202-213); the product
    (kgvae.model.models.SAIL.bits_per_sequence) uses ONE teacher-forced pass.  Pinned here on the CPU port against
    the reference's own posterior_bits output (tests/golden/eval_bits.*)."""
    import json
    import os

    from conftest import GOLDEN
    from oracle.torch_cpu_port import posterior_bits_port
    from kgvae.model.utils import GraphSeqDataset
    arr = dict(np.load(os.path.join(GOLDEN, "eval_bits.npz")))
    with open(os.path.join(GOLDEN, "eval_bits.json")) as f:
        meta = json.load(f)["sail"]
    cfg = meta["cfg"]
    m = _load(cfg, {k[len("sail_param::"):]: v for k, v in arr.items() if k.startswith("sail_param::")})
    ds = GraphSeqDataset([[tuple(t) for t in g] for g in meta["graphs"]], None, None, use_padding=True, pad_eid=cfg["pad_eid"],
                         pad_rid=cfg["pad_rid"], max_triples=cfg["max_edges"], special_tokens=cfg["special_tokens"],
                         ent_base=cfg["ENT_BASE"], rel_base=cfg["REL_BASE"], seq_len=cfg["seq_len"])
    got = [posterior_bits_port(m, *ds[i], torch.from_numpy(arr["sail_eps"][i])) for i in range(len(ds))]
    np.testing.assert_allclose([g[0] for g in got], arr["sail_ar_bits"], rtol=1e-5)
    np.testing.assert_allclose([g[1] for g in got], arr["sail_kl_bits"], rtol=1e-5)
    np.testing.assert_allclose(np.mean([a + k for a, k in got]), meta["stats"]["avg_total_bits"], rtol=1e-5)
