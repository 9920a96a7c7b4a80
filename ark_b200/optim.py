"""torch.optim-compatible front of the fused Adam kernel (K10).

`FusedAdam(model, lr)` looks like `torch.optim.Adam(model.parameters(), lr)` to the rest of the trainer
(param_groups for LR schedulers, state_dict()/load_state_dict() with per-parameter step/exp_avg/exp_avg_sq so
checkpoints keep the reference's `optimizer_state_dict` layout, ablation_study.py:571,735-791), but step() is
ONE kernel launch over the engine's flat fp32 buffers that also refreshes the bf16 weight shadow.
"""
from __future__ import annotations

import torch


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, model, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, **engine_kw):
        self.engine = model.engine(lr=lr, betas=betas, eps=eps, **engine_kw)
        self.engine.lr, self.engine.betas, self.engine.eps = float(lr), tuple(betas), float(eps)
        # parameters are registered in MODULE order (model.parameters()), exactly like the reference's
        # Adam(model.parameters()) (ablation_study.py:571): torch indexes optimizer_state_dict entries by position, so
        # reference checkpoints load into FusedAdam and vice versa.  The flat buffers are ordered differently
        # (gradient-readiness order); the per-parameter views below do not care.
        named = list(model.named_parameters())
        names = [n for n, _ in named]
        params = [p for _, p in named]
        missing = set(names) ^ set(self.engine.flat.order)
        if missing:
            raise RuntimeError(f"optimizer / flat-buffer parameter mismatch: {sorted(missing)}")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=0, amsgrad=False))
        f = self.engine.flat
        for n, p in zip(names, params):   # expose the flat Adam state through the usual per-parameter dicts
            self.state[p] = {"step": torch.tensor(0.0), "exp_avg": f.view(f.exp_avg, n),
                             "exp_avg_sq": f.view(f.exp_avg_sq, n)}

    def zero_grad(self, set_to_none=False):
        """No-op: the fused backward OVERWRITES every gradient slot each step (nothing accumulates)."""

    @torch.no_grad()
    def step(self, closure=None):
        g = self.param_groups[0]
        self.engine.betas, self.engine.eps = tuple(g["betas"]), float(g["eps"])
        self.engine.adam_step(lr=float(g["lr"]))
        for st in self.state.values():
            st["step"] = torch.tensor(float(self.engine.step_count))

    def sync_from_engine(self):
        """After fused steps (SAIL.elbo_step / ARK.ce_step run Adam inside the engine): publish the engine's step
        count through the per-parameter state so checkpoints and LR schedulers see it."""
        g = self.param_groups[0]
        self.engine.betas, self.engine.eps = tuple(g["betas"]), float(g["eps"])
        for st in self.state.values():
            st["step"] = torch.tensor(float(self.engine.step_count))
        self._opt_called = True      # (LR schedulers check that the optimiser stepped before they do)

    def load_state_dict(self, state_dict):
        f = self.engine.flat
        keep = {p: (st["exp_avg"], st["exp_avg_sq"]) for p, st in self.state.items()}
        super().load_state_dict(state_dict)
        steps = []
        for p, st in self.state.items():   # copy loaded moments back INTO the flat buffers, keep the views
            ea, es = keep[p]
            ea.copy_(st["exp_avg"].to(ea.device))
            es.copy_(st["exp_avg_sq"].to(es.device))
            steps.append(float(st["step"]))
            st["exp_avg"], st["exp_avg_sq"] = ea, es
        if steps:
            self.engine.step_count = int(max(steps))
