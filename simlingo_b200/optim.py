"""Fused AdamW over the flat parameter store.

Drop-in for ``torch.optim.AdamW(self.parameters(), lr, weight_decay, betas)`` as configured by the reference
(``simlingo_training/models/driving.py:718-724``) plus ``Trainer(gradient_clip_val=0.3)`` (``train.py:206``): one
squared-norm kernel and one update kernel over the whole flat range (fp32 master weights and moments, bf16 gradients,
bf16 model copy written back in the same pass), with the global-norm clip and the data-parallel 1/world averaging
folded in as a gradient scale.  ``param_groups[0]['lr']`` / ``['betas']`` are re-read every step, so
``torch.optim.lr_scheduler.OneCycleLR`` (which also cycles beta1, the reference keeps ``cycle_momentum=True``)
drives it unchanged."""
from __future__ import annotations

from typing import Iterable, Optional

import torch

from . import lib
from .training import ParamStore


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, params: Iterable[torch.nn.Parameter], store: ParamStore, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 1e-2, max_grad_norm: float = 0.0):
        params = list(params)
        super().__init__(params, dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay))
        # torch's LR schedulers with cycle_momentum look for 'betas' in defaults: present above
        self.store = store
        known = {id(p) for p in store.params.values()}
        missing = [p for p in params if id(p) not in known]
        if missing:
            raise RuntimeError(f"{len(missing)} parameters handed to FusedAdamW are not in the flat parameter store")
        self.max_grad_norm = float(max_grad_norm)
        n = store.numel
        dev = store.flat_param.device
        self.master = store.flat_param.float()
        self.exp_avg = torch.zeros(n, device=dev, dtype=torch.float32)
        self.exp_avg_sq = torch.zeros(n, device=dev, dtype=torch.float32)
        self.sqnorm = torch.zeros(1, device=dev, dtype=torch.float32)
        self.step_count = 0
        self.launches = 0

    def resync_master(self) -> None:
        """After ``load_state_dict`` on the model: take the bf16 parameters as the new fp32 master copy."""
        self.master.copy_(self.store.flat_param)

    def zero_grad(self, set_to_none: bool = True) -> None:
        self.store.zero_grad()

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        st = self.store
        st.wait_exchange()
        g = self.param_groups[0]
        self.step_count += 1
        clip = self.max_grad_norm
        sq: Optional[torch.Tensor] = None
        if clip > 0:
            self.sqnorm.zero_()
            lib.grad_sqnorm(st.flat_grad, self.sqnorm)
            sq = self.sqnorm
            self.launches += 1
        b1, b2 = g["betas"]
        lib.adamw_fused(self.master, self.exp_avg, self.exp_avg_sq, st.flat_grad, st.flat_param, float(g["lr"]), float(b1), float(b2),
                        float(g["eps"]), float(g["weight_decay"]), self.step_count, sqnorm=sq, max_norm=clip, prescale=1.0 / st.world)
        self.launches += 1
        st.generation += 1
        return loss

    def grad_norm(self) -> float:
        """Global gradient norm of the last step (after averaging over ranks, before clipping); host sync."""
        return float(self.sqnorm.sqrt().item()) / self.store.world

    def state_dict(self):
        d = super().state_dict()
        d["slb"] = dict(master=self.master, exp_avg=self.exp_avg, exp_avg_sq=self.exp_avg_sq, step=self.step_count)
        return d

    def load_state_dict(self, state_dict):
        slb = state_dict.pop("slb", None)
        super().load_state_dict(state_dict)
        if slb is not None:
            self.master.copy_(slb["master"]); self.exp_avg.copy_(slb["exp_avg"]); self.exp_avg_sq.copy_(slb["exp_avg_sq"])
            self.step_count = int(slb["step"])
            lib.load()
            self.store.flat_param.copy_(self.master)
