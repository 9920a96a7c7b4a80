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
