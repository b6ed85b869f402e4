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
        key_of = {id(p): k for k, p in store.params.items()}
        missing = [p for p in params if p.requires_grad and id(p) not in key_of]
        if missing:
            raise RuntimeError(f"{len(missing)} parameters handed to FusedAdamW are not in the flat parameter store")
        # (index in param_groups order, flat offset, numel, shape) of every trainable parameter.  Frozen parameters may be
        # passed too (the reference hands AdamW all of ``self.parameters()``, driving.py:718): they keep their index, so
        # that state_dict() lines up with a torch.optim.AdamW built the reference's way, and never get any state.
        self._slices = [(i, *store.offsets[key_of[id(p)]], p.shape) for i, p in enumerate(params) if id(p) in key_of]
        self.max_grad_norm = float(max_grad_norm)
        n = store.numel
        dev = store.flat_param.device
        self.master = store.flat_param.float()
        self.exp_avg = torch.zeros(n, device=dev, dtype=torch.float32)
        self.exp_avg_sq = torch.zeros(n, device=dev, dtype=torch.float32)
        self.sqnorm = torch.zeros(1, device=dev, dtype=torch.float32)
        self.step_count = 0
        self.launches = 0
        self._epoch = store.weights_epoch

    def check_attached(self) -> None:
        """The parameters must still live in the flat buffers this optimizer updates.  ``model.load_state_dict`` / ``.to()``
        rebuild the engines (and with them the flat store), so an optimizer created *before* them would silently update
        orphaned buffers: create it after loading the weights (the reference's order, train.py:104-111 then trainer.fit)."""
        flat = self.store.flat_param
        p = next(iter(self.store.params.values()))
        if not flat.data_ptr() <= p.data_ptr() < flat.data_ptr() + flat.numel() * flat.element_size():
            raise RuntimeError("FusedAdamW: the model's parameters no longer live in this optimizer's flat store (the model was re-loaded or "
                               "moved after the optimizer was created); build the optimizer after load_state_dict / .to()")

    def resync_master(self) -> None:
        """After ``load_state_dict`` on the model: take the bf16 parameters as the new fp32 master copy."""
        self.master.copy_(self.store.flat_param)
        self._epoch = self.store.weights_epoch

    def zero_grad(self, set_to_none: bool = True) -> None:
        self.store.zero_grad()

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        st = self.store
        self.check_attached()
        if st.weights_epoch != self._epoch:   # the model was re-loaded in place since the last step: its values are the truth
            self.resync_master()
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
        # normally one range = the whole buffer; parameters that received no gradient this step (a sub-network that did not
        # run backward, a head torch left without .grad) are skipped as torch.optim.AdamW skips ``p.grad is None``
        for a, b in st.update_ranges():
            lib.adamw_fused(self.master[a:b], self.exp_avg[a:b], self.exp_avg_sq[a:b], st.flat_grad[a:b], st.flat_param[a:b], float(g["lr"]),
                            float(b1), float(b2), float(g["eps"]), float(g["weight_decay"]), self.step_count, sqnorm=sq, max_norm=clip,
                            prescale=1.0 / st.world)
            self.launches += 1
        st.generation += 1
        return loss

    def grad_norm(self) -> float:
        """Global gradient norm of the last step (after averaging over ranks, before clipping); host sync."""
        return float(self.sqnorm.sqrt().item()) / self.store.world

    def state_dict(self):
        """``torch.optim.AdamW`` layout: ``state[i] = {step, exp_avg, exp_avg_sq}`` per parameter index (views into the flat
        fp32 moment buffers) + ``param_groups``, so Lightning's checkpoint connector — or the reference's own optimizer —
        can resume from it.  The fp32 master weights travel as an extra ``master`` entry per parameter, which
        ``torch.optim.Optimizer.load_state_dict`` carries along untouched."""
        groups = [{**{k: v for k, v in g.items() if k != "params"}, "params": list(range(len(g["params"])))} for g in self.param_groups]
        state = {}
        if self.step_count > 0:
            for i, o, n, shape in self._slices:
                state[i] = {"step": torch.tensor(float(self.step_count)), "exp_avg": self.exp_avg[o:o + n].view(shape),
                            "exp_avg_sq": self.exp_avg_sq[o:o + n].view(shape), "master": self.master[o:o + n].view(shape)}
        return {"state": state, "param_groups": groups}

    @torch.no_grad()
    def load_state_dict(self, state_dict) -> None:
        """Accepts its own ``state_dict()`` or one written by ``torch.optim.AdamW`` over the same parameter list (then the
        master weights are taken from the model's current parameters)."""
        groups = state_dict["param_groups"]
        if len(groups) != 1 or len(groups[0]["params"]) != len(self.param_groups[0]["params"]):
            raise ValueError("optimizer state does not match: expected one parameter group over the same parameter list")
        for k in ("lr", "betas", "eps", "weight_decay", "initial_lr", "max_lr", "min_lr", "base_momentum", "max_momentum"):
            if k in groups[0]:
                self.param_groups[0][k] = tuple(groups[0][k]) if k == "betas" else groups[0][k]
        state = state_dict["state"]
        if not state:
            self.step_count = 0
            self.exp_avg.zero_(); self.exp_avg_sq.zero_()
            self.resync_master()
            return
        steps = {int(float(state[i]["step"])) for i, *_ in self._slices if i in state}
        absent = [i for i, *_ in self._slices if i not in state]
        if absent or len(steps) != 1:
            raise ValueError(f"optimizer state does not match: {len(absent)} trainable parameters without state, step counts {sorted(steps)}")
        self.step_count = steps.pop()
        have_master = all("master" in state[i] for i, *_ in self._slices)
        for i, o, n, shape in self._slices:
            e = state[i]
            if tuple(e["exp_avg"].shape) != tuple(shape):
                raise ValueError(f"optimizer state does not match: parameter {i} has shape {tuple(shape)}, state {tuple(e['exp_avg'].shape)}")
            self.exp_avg[o:o + n].view(shape).copy_(e["exp_avg"])
            self.exp_avg_sq[o:o + n].view(shape).copy_(e["exp_avg_sq"])
            if have_master:
                self.master[o:o + n].view(shape).copy_(e["master"])
        if have_master:
            self.store.flat_param.copy_(self.master)
            self._epoch = self.store.weights_epoch
        else:
            self.resync_master()
