"""Conditioning encoder with the reference's ``GEncoder`` API.

``GEncoder(input_size, rep_size)`` (/root/reference/il-and-cil/utils/torch_fcnf.py:154-190) is the MLP every
reference caller runs immediately before the flow to produce its conditioning input (``torch_nfbc_ant.py:148,267``):
four times ``Linear(512) -> LayerNorm -> SiLU`` and a final ``Linear(rep_size)``.  This module keeps the attribute
names (``fc1..fc4``, ``norm1..norm4``, ``fc_out``) and therefore the ``state_dict`` keys, holds the parameters in one
flat arena laid out by the C library (``nf_encoder_param_layout``), and runs forward and backward through the C ABI
(``nf_encoder_forward`` / ``nf_encoder_backward``): tcgen05 GEMMs plus two row-wise kernels.  CUDA only; there is no
CPU fallback.
"""
from __future__ import annotations

import ctypes

import torch
from torch import nn

from . import _lib
from ._lib import lib
from .module import _accumulate_nodes, _dense, _on, _param_grads_wanted, _SlotGuard, _StashSlot


class _Encode(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model: "GEncoder", need_stash: bool, nodes, g, *params):
        out, stash, slot = model._run_forward(g, need_stash)
        ctx.model, ctx.stash, ctx.nodes = model, stash, nodes
        ctx.guard = _SlotGuard(slot) if slot is not None else None
        ctx.save_for_backward(g)
        return out

    @staticmethod
    def backward(ctx, dout):
        model = ctx.model
        if ctx.stash is None:
            raise RuntimeError("nf_b200: encoder backward without a stash (forward ran under no_grad, or backward twice)")
        (g,) = ctx.saved_tensors
        need_p = _param_grads_wanted(ctx.needs_input_grad[4:], ctx.nodes)
        dg, dflat = model._run_backward(g, ctx.stash, dout, ctx.needs_input_grad[3], need_p)
        ctx.stash = None
        if ctx.guard is not None:
            ctx.guard.release()      # the stash may be reused by the next forward
        if not need_p:
            return (None, None, None, dg) + (None,) * (len(ctx.needs_input_grad) - 4)
        views = model._grad_views(dflat)
        return (None, None, None, dg) + tuple(v if need else None for v, need in zip(views, ctx.needs_input_grad[4:]))


class GEncoder(nn.Module):
    """``GEncoder(input_size, rep_size)``; extras are keyword-only: ``hidden`` (512), ``n_hidden`` (4), ``ln_eps``
    (1e-5; the flax twin nf.py:188-212 uses 1e-6) and ``precision`` ("fp32" = tcgen05 3xTF32, "fp32_simt" = FFMA)."""

    def __init__(self, input_size, rep_size, *, hidden: int = 512, n_hidden: int = 4, ln_eps: float = 1e-5,
                 precision: str = "fp32"):
        super().__init__()
        if precision not in ("fp32", "tf32x3", "fp32_simt", "tf32"):
            raise ValueError("precision must be 'fp32', 'fp32_simt' or 'tf32'")
        self.input_size, self.rep_size = int(input_size), int(rep_size)
        self.hidden, self.n_hidden, self.ln_eps, self.precision = int(hidden), int(n_hidden), float(ln_eps), precision
        fan = self.input_size
        for i in range(1, self.n_hidden + 1):
            setattr(self, f"fc{i}", nn.Linear(fan, self.hidden))
            fan = self.hidden
        self.fc_out = nn.Linear(self.hidden, self.rep_size)
        for i in range(1, self.n_hidden + 1):
            setattr(self, f"norm{i}", nn.LayerNorm(self.hidden, eps=self.ln_eps))
        self._desc = _lib.NfEncoderDesc(self.input_size, self.hidden, self.rep_size, self.n_hidden, self.ln_eps,
                                        _lib.PREC[precision])
        offs = (ctypes.c_int64 * (4 * self.n_hidden + 2))()
        total = lib.nf_encoder_param_layout(ctypes.byref(self._desc), offs)
        if total < 0:
            raise _lib.NfError("nf_encoder_param_layout: " + lib.nf_last_error().decode(errors="replace"))
        self._offs, self._total = list(offs), int(total)
        self._flat = None
        self._ws = None
        self._stash_pool = {}
        self._dp_group, self._dp_world, self._dp_avg = None, 1, False
        self._flatten()
        self.register_load_state_dict_post_hook(lambda module, keys: module._flatten())

    # ------------------------------------------------------------------ parameter arena
    def _slot_params(self):
        out = []
        for i in range(self.n_hidden):
            fc, norm = getattr(self, f"fc{i + 1}"), getattr(self, f"norm{i + 1}")
            out += [(fc.weight, self._offs[4 * i]), (fc.bias, self._offs[4 * i + 1]),
                    (norm.weight, self._offs[4 * i + 2]), (norm.bias, self._offs[4 * i + 3])]
        out += [(self.fc_out.weight, self._offs[4 * self.n_hidden]), (self.fc_out.bias, self._offs[4 * self.n_hidden + 1])]
        return out

    def _flatten(self):
        sp = self._slot_params()
        dev, dtype = sp[0][0].device, sp[0][0].dtype
        flat = torch.zeros(self._total, dtype=dtype, device=dev)
        with torch.no_grad():
            for p, off in sp:
                view = flat[off:off + p.numel()].view(p.shape)
                view.copy_(p.detach())
                p.data = view
        self._flat, self._slots = flat, sp

    def _apply(self, fn, recurse=True):
        out = super()._apply(fn, recurse)
        self._flatten()
        self._ws = None
        self._stash_pool = {}
        return out

    def _is_flat(self):
        base = self._flat.data_ptr()
        return all(p.data_ptr() == base + 4 * off for p, off in (self._slots[0], self._slots[-1]))

    def _grad_views(self, dflat):
        return [dflat[off:off + p.numel()].view(p.shape) for p, off in self._slots]

    # ------------------------------------------------------------------ plumbing
    def _check(self, g):
        if not g.is_cuda:
            raise RuntimeError("nf_b200.GEncoder runs on CUDA only (sm_100a kernels, no CPU fallback)")
        if g.dim() != 2 or g.shape[1] != self.input_size:
            raise ValueError(f"g must be (B, {self.input_size}), got {tuple(g.shape)}")
        if g.dtype != torch.float32:
            raise TypeError("nf_b200 computes in float32: g must be float32")
        if not self._is_flat():
            self._flatten()
        if self._flat.device != g.device or self._flat.dtype != torch.float32:
            raise RuntimeError("module parameters must be float32 on the same device as the input")

    def _workspace(self, B, dev):
        need = lib.nf_encoder_workspace_bytes(ctypes.byref(self._desc), B)
        if self._ws is None or self._ws.numel() < need or self._ws.device != dev:
            self._ws = torch.empty(need, dtype=torch.uint8, device=dev)
        return self._ws

    def _stream(self, dev):
        return ctypes.c_void_p(torch._C._cuda_getCurrentRawStream(dev.index if dev.index is not None
                                                                  else torch.cuda.current_device()))

    def _run_forward(self, g, need_stash):
        dev, B = g.device, g.shape[0]
        g = _dense(g)
        with _on(dev):
            ws = self._workspace(B, dev)
            out = torch.empty((B, self.rep_size), dtype=torch.float32, device=dev)
            stash, slot = None, None
            if need_stash:
                stash, slot = self._acquire_stash(B, dev)
            _lib.check(lib.nf_encoder_forward(ctypes.byref(self._desc), self._flat.data_ptr(), g.data_ptr(), B,
                                              out.data_ptr(), stash.data_ptr() if stash is not None else None,
                                              ws.data_ptr(), ws.numel(), self._stream(dev)), "nf_encoder_forward")
        return out, stash, slot

    _MAX_STASH_SLOTS = 4

    def _acquire_stash(self, B, dev):
        """A free persistent stash for batch size B (the library's launch graphs are keyed on it), or a fresh
        unpooled one when all are in flight."""
        pool = self._stash_pool.setdefault((B, dev), [])
        for slot in pool:
            if not slot.busy:
                slot.busy = True
                return slot.tensor, slot
        t = torch.empty(lib.nf_encoder_stash_bytes(ctypes.byref(self._desc), B), dtype=torch.uint8, device=dev)
        if len(pool) < self._MAX_STASH_SLOTS:
            slot = _StashSlot(t)
            slot.busy = True
            pool.append(slot)
            return t, slot
        return t, None

    def _run_backward(self, g, stash, dout, need_dg, need_p=True):
        dev, B = g.device, g.shape[0]
        g, dout = _dense(g), _dense(dout)
        with _on(dev):
            ws = self._workspace(B, dev)
            dg = torch.empty_like(g) if need_dg else None
            dflat = torch.empty(self._total, dtype=torch.float32, device=dev)
            _lib.check(lib.nf_encoder_backward(ctypes.byref(self._desc), self._flat.data_ptr(),
                                               stash.data_ptr(), B, dout.data_ptr(),
                                               dg.data_ptr() if dg is not None else None, dflat.data_ptr(),
                                               ws.data_ptr(), ws.numel(), self._stream(dev)), "nf_encoder_backward")
            if self._dp_world > 1 and need_p:     # (a rank-local evaluation must not enter the collective)
                self._allreduce_grads(dflat)
        return dg, dflat

    # ------------------------------------------------------------------ data parallel
    def enable_data_parallel(self, process_group=None):
        """Batch-sharded data parallel training, like ``RealNVP.enable_data_parallel``: backward averages the flat
        gradient arena over the ranks of ``process_group`` (one all-reduce) before the views are handed to autograd."""
        import torch.distributed as dist
        self._dp_group = process_group
        self._dp_world = dist.get_world_size(process_group)
        self._dp_avg = dist.get_backend(process_group) == "nccl"
        if self._dp_world > 1:      # replicas start from rank 0's parameters
            if not self._is_flat():
                self._flatten()
            with torch.no_grad():
                dist.broadcast(self._flat, src=dist.get_global_rank(process_group, 0) if process_group is not None else 0,
                               group=process_group)
        return self

    def _allreduce_grads(self, dflat):
        import torch.distributed as dist
        if self._dp_avg:
            dist.all_reduce(dflat, op=dist.ReduceOp.AVG, group=self._dp_group)
        else:
            dist.all_reduce(dflat, op=dist.ReduceOp.SUM, group=self._dp_group)
            dflat.mul_(1.0 / self._dp_world)

    def forward(self, g: torch.Tensor) -> torch.Tensor:
        self._check(g)
        params = [p for p, _ in self._slots]
        # (grad mode is off inside Function.forward, so whether a backward can follow is decided here)
        need_stash = torch.is_grad_enabled() and (g.requires_grad or any(p.requires_grad for p in params))
        nodes = _accumulate_nodes([p for p in params if p.requires_grad]) if need_stash else []
        return _Encode.apply(self, need_stash, nodes, g, *params)
