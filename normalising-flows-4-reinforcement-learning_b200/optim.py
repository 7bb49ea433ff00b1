"""The optimizer step that follows the flow's backward in every reference training loop
(il-and-cil/torch_nfbc_ant.py:150-155, 271-274), fused over the flat parameter arena.

``FlatAdamW`` is a ``torch.optim.Optimizer`` (so the reference's ``CosineLRSchedule``,
il-and-cil/utils/torch_utils.py:21-49, drives it through ``param_groups`` unchanged): parameters that live
in a ``RealNVP`` arena are updated by ONE library kernel per flow (nf_adamw_step) reading the flat gradient
arena; any other parameter (the conditioning encoder) goes through ``torch.optim.AdamW``'s own functional
path with the same hyper-parameters.
"""
from __future__ import annotations

import ctypes
import math
from typing import Iterable, Sequence

import torch

from . import _lib
from ._lib import lib
from .module import _on


class FlatAdamW(torch.optim.Optimizer):
    def __init__(self, params: Iterable, flows: Sequence = (), lr=1e-3, betas=(0.9, 0.999), eps=1e-8,
                 weight_decay=1e-2):
        params = list(params)
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        super().__init__(params, defaults)
        self._flows = list(flows)
        flow_ids = set()
        for f in self._flows:
            flow_ids.update(id(p) for p in f.parameters())
        # parameters outside the flow arenas keep their param-group structure (and per-group hyper-parameters)
        self._other_groups = []
        for gi, g in enumerate(self.param_groups):
            rest = [p for p in g["params"] if id(p) not in flow_ids]
            if rest:
                self._other_groups.append((gi, rest))
        self._other_opt = (torch.optim.AdamW([dict(params=rest) for _, rest in self._other_groups], **defaults)
                           if self._other_groups else None)
        # a flow's arena is updated by one kernel with one set of hyper-parameters: all of its parameters must sit in
        # the same param group (that group's lr / betas / eps / weight_decay apply)
        self._flow_group = {}
        for f in self._flows:
            owners = {gi for gi, g in enumerate(self.param_groups) for p in g["params"]
                      if any(p is q for q in f._tparams if q.requires_grad)}
            if len(owners) != 1:
                raise ValueError("FlatAdamW: the trainable parameters of a flow must all belong to exactly one "
                                 f"param group (found {len(owners)})")
            self._flow_group[id(f)] = owners.pop()
        self._fstate = {}     # id(flow) -> dict(step, exp_avg, exp_avg_sq, mask)

    def _flow_state(self, flow):
        st = self._fstate.get(id(flow))
        flat = flow._flat
        if st is None or st["exp_avg"].data_ptr() == 0 or st["exp_avg"].device != flat.device or \
                st["exp_avg"].numel() != flat.numel():
            mask = torch.zeros(flow.n_layers * _lib.NF_P_SLOTS, dtype=torch.uint8)
            st = dict(step=0, exp_avg=torch.zeros_like(flat), exp_avg_sq=torch.zeros_like(flat), mask=mask, mask_key=None)
            self._fstate[id(flow)] = st
        key = tuple(p.requires_grad for p, _ in flow._slots)
        if st["mask_key"] != key:
            st["mask"] = torch.tensor([1 if r else 0 for r in key], dtype=torch.uint8, device=flat.device)
            st["mask_key"] = key
        return st

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for flow in self._flows:
            group = self.param_groups[self._flow_group[id(flow)]]
            lr, (b1, b2), eps, wd = group["lr"], group["betas"], group["eps"], group["weight_decay"]
            if not flow._is_flat():
                flow._flatten()
            if all(p.grad is None for p in flow._tparams if p.requires_grad):
                continue          # torch.optim.AdamW skips parameters without a gradient (no decay, no moment update)
            g = flow._gflat
            views_ok = g is not None and all(
                (not p.requires_grad) or (p.grad is not None and p.grad.data_ptr() == v.data_ptr())
                for p, v in zip(flow._tparams, flow._gviews or ()))
            if not views_ok:    # gradients did not come from the flat-arena backward: gather them
                g = torch.zeros_like(flow._flat)
                for p, off in flow._trainable:
                    if p.grad is not None:
                        g[off:off + p.numel()].copy_(p.grad.reshape(-1))
            st = self._flow_state(flow)
            st["step"] += 1
            with _on(flow._flat.device):
                _lib.check(lib.nf_adamw_step(
                    ctypes.byref(flow._desc), flow._flat.data_ptr(), g.data_ptr(), st["exp_avg"].data_ptr(),
                    st["exp_avg_sq"].data_ptr(), st["mask"].data_ptr(), float(lr), float(b1), float(b2), float(eps),
                    float(wd), st["step"], flow._stream(flow._flat.device)), "nf_adamw_step")
            flow._packed_key = None           # parameters changed: rebuild the derived tables on the next call
            flow._bump_version()              # ... and make the change visible to the forward/backward version check
        if self._other_opt is not None:
            for og, (gi, _) in zip(self._other_opt.param_groups, self._other_groups):
                src = self.param_groups[gi]
                og.update(lr=src["lr"], betas=src["betas"], eps=src["eps"], weight_decay=src["weight_decay"])
            self._other_opt.step()
        return loss

    # ------------------------------------------------------------------ checkpointing
    def state_dict(self):
        """torch's format plus the flat moments of every flow and the delegate optimizer's state (the flow moments
        live outside ``self.state``: one tensor per arena, not one per parameter)."""
        sd = super().state_dict()
        sd["nf_flows"] = [dict(step=st["step"], exp_avg=st["exp_avg"].clone(), exp_avg_sq=st["exp_avg_sq"].clone())
                          if (st := self._fstate.get(id(f))) is not None else None for f in self._flows]
        sd["nf_other"] = self._other_opt.state_dict() if self._other_opt is not None else None
        return sd

    def load_state_dict(self, state_dict):
        state_dict = dict(state_dict)
        flows = state_dict.pop("nf_flows", None)
        other = state_dict.pop("nf_other", None)
        super().load_state_dict(state_dict)
        if flows is not None:
            if len(flows) != len(self._flows):
                raise ValueError("FlatAdamW.load_state_dict: number of flows differs")
            for f, saved in zip(self._flows, flows):
                if saved is None:
                    self._fstate.pop(id(f), None)
                    continue
                st = self._flow_state(f)
                st["step"] = int(saved["step"])
                st["exp_avg"].copy_(saved["exp_avg"])
                st["exp_avg_sq"].copy_(saved["exp_avg_sq"])
        if other is not None and self._other_opt is not None:
            self._other_opt.load_state_dict(other)


class CosineLRSchedule(torch.nn.Module):
    """Linear warm-up then cosine decay, the schedule of il-and-cil/utils/torch_utils.py:21-49 (same
    constructor, ``step()`` returns the new learning rate and writes it to ``optimizer.param_groups``)."""

    def __init__(self, optimizer, warmup_steps: int, total_steps: int, min_lr: float, max_lr: float):
        super().__init__()
        self.register_buffer("counter", torch.zeros(()))
        self.warmup_steps, self.total_steps = warmup_steps, total_steps
        self.optimizer, self.min_lr, self.max_lr = optimizer, min_lr, max_lr
        self._n = 0
        self.set_lr(min_lr)
        # resume: the reference reads the step count from the `counter` buffer (torch_utils.py:33-36), so a loaded
        # checkpoint continues the schedule instead of restarting the warm-up
        self.register_load_state_dict_post_hook(lambda module, keys: module._resync())

    def _resync(self):
        self._n = int(round(float(self.counter.item())))

    def set_lr(self, lr: float) -> float:
        if self.min_lr <= lr <= self.max_lr:
            for pg in self.optimizer.param_groups:
                pg["lr"] = lr
        return max(self.min_lr, min(self.max_lr, lr))

    def step(self) -> float:
        self._n += 1
        self.counter.fill_(self._n)      # kept for state_dict compatibility; no device sync (the reference calls .item())
        if self._n <= self.warmup_steps:
            return self.set_lr(self.min_lr + self._n / self.warmup_steps * (self.max_lr - self.min_lr))
        t = (self._n - self.warmup_steps) / (self.total_steps - self.warmup_steps)
        return self.set_lr(self.min_lr + 0.5 * (1 + math.cos(math.pi * t)) * (self.max_lr - self.min_lr))
