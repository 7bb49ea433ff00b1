"""Causal-transformer flow with the call-site API of the reference's goal-conditioned BC script (SURVEY.md section 8
row f1).

``il-and-cil/torch_nfgcbc.py:27`` imports ``RealNVP`` from ``utils.torch_networks`` -- a file that is absent from the
reference tree (PARITY UNPINNED; upstream follows apple/ml-tarflow).  What the reference pins is the contract, and this
module keeps it:

    model = RealNVP(input_dims, channels, rep_dim, head_dim, blocks, layers_per_block, prior)      # torch_nfgcbc.py:119
    z, outputs, logdets = model(actions (B, A, 1), cond (B, 1, rep))                                # :136-142
    actions = model.reverse(prior_sample (B, A, 1), cond)                                           # torch_evaluations.py:79-82

The arithmetic is the TarFlow meta block as restated by the checker ``oracle/torch_tarflow.py`` (same sub-module names,
same ``state_dict`` keys, same-seed initialisation).  Parameters live in one flat arena laid out by the C library
(``nf_tarflow_param_layout``); forward, the autoregressive inverse and backward run through the C ABI
(``nf_tarflow_forward`` / ``_inverse`` / ``_backward``, csrc/tarflow.cu).  CUDA only; there is no CPU fallback.
"""
from __future__ import annotations

import ctypes

import torch
from torch import nn

from . import _lib
from ._lib import lib
from .module import _accumulate_nodes, _dense, _on, _param_grads_wanted, _SlotGuard, _StashSlot


class _Attention(nn.Module):
    def __init__(self, channels, head_dim, ln_eps):
        super().__init__()
        self.norm = nn.LayerNorm(channels, eps=ln_eps)
        self.qkv = nn.Linear(channels, 3 * channels)
        self.proj = nn.Linear(channels, channels)


class _MLP(nn.Module):
    def __init__(self, channels, expansion, ln_eps):
        super().__init__()
        self.norm = nn.LayerNorm(channels, eps=ln_eps)
        self.fc1 = nn.Linear(channels, channels * expansion)
        self.fc2 = nn.Linear(channels * expansion, channels)


class _AttentionBlock(nn.Module):
    def __init__(self, channels, head_dim, expansion, ln_eps):
        super().__init__()
        self.attn = _Attention(channels, head_dim, ln_eps)
        self.mlp = _MLP(channels, expansion, ln_eps)


class TransformerMetaBlock(nn.Module):
    """Parameter container of one block (construction order = the checker's, so equal seeds give equal weights)."""

    def __init__(self, A, channels, rep_dim, head_dim, layers, expansion, ln_eps):
        super().__init__()
        self.proj_in = nn.Linear(1, channels)
        self.cond_proj = nn.Linear(rep_dim, channels)
        self.pos_embed = nn.Parameter(torch.randn(A, channels) * 1e-2)
        self.layers = nn.ModuleList([_AttentionBlock(channels, head_dim, expansion, ln_eps) for _ in range(layers)])
        self.proj_out = nn.Linear(channels, 2)
        nn.init.zeros_(self.proj_out.weight)      # a fresh block is the identity
        nn.init.zeros_(self.proj_out.bias)


class _Flow(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model: "TransformerRealNVP", need_stash: bool, nodes, x, y, *params):
        z, ld, outs, stash, slot = model._run_forward(x, y, need_stash)
        ctx.model, ctx.stash, ctx.nodes, ctx.B = model, stash, nodes, x.shape[0]
        ctx.guard = _SlotGuard(slot) if slot is not None else None
        ctx.mark_non_differentiable(outs)
        return z, ld, outs

    @staticmethod
    def backward(ctx, dz, dld, _douts):
        model = ctx.model
        if ctx.stash is None:
            raise RuntimeError("nf_b200: transformer-flow backward without a stash (forward ran under no_grad, or backward twice)")
        need_p = _param_grads_wanted(ctx.needs_input_grad[5:], ctx.nodes)
        dx, dy, dflat = model._run_backward(ctx.stash, ctx.B, dz, dld, ctx.needs_input_grad[3], ctx.needs_input_grad[4],
                                            need_p)
        ctx.stash = None
        if ctx.guard is not None:
            ctx.guard.release()
        head = (None, None, None, dx, dy)
        if not need_p:
            return head + (None,) * (len(ctx.needs_input_grad) - 5)
        views = model._grad_views(dflat)
        return head + tuple(v if need else None for v, need in zip(views, ctx.needs_input_grad[5:]))


class TransformerRealNVP(nn.Module):
    """``RealNVP(input_dims, channels, rep_dim, head_dim, blocks, layers_per_block, prior)`` of torch_nfgcbc.py:119.
    Keyword-only extras: ``expansion`` (4), ``ln_eps`` (1e-5), ``logscale_clamp`` (0 = raw log-scale; c > 0 bounds it as
    c tanh(raw / c)), ``precision`` ("fp32" = tcgen05 3xTF32, "fp32_simt" = FFMA checker)."""

    def __init__(self, in_channels, channels, rep_dim, head_dim, blocks, layers_per_block, prior=None, *,
                 expansion: int = 4, ln_eps: float = 1e-5, logscale_clamp: float = 0.0, precision: str = "fp32"):
        super().__init__()
        if precision not in ("fp32", "tf32x3", "fp32_simt"):
            raise ValueError("precision must be 'fp32' or 'fp32_simt'")
        self.in_channels, self.channels, self.rep_dim = int(in_channels), int(channels), int(rep_dim)
        self.head_dim, self.n_blocks, self.n_layers = int(head_dim), int(blocks), int(layers_per_block)
        self.expansion, self.ln_eps, self.logscale_clamp = int(expansion), float(ln_eps), float(logscale_clamp)
        self.precision = precision
        self.prior = prior
        self.blocks = nn.ModuleList([
            TransformerMetaBlock(self.in_channels, self.channels, self.rep_dim, self.head_dim, self.n_layers,
                                 self.expansion, self.ln_eps) for _ in range(self.n_blocks)])
        self._desc = _lib.NfTarflowDesc(self.in_channels, self.channels, self.rep_dim, self.head_dim, self.n_blocks,
                                        self.n_layers, self.expansion, self.ln_eps, self.logscale_clamp,
                                        _lib.PREC[precision])
        offs = (ctypes.c_int64 * _lib.NF_TARFLOW_LAYOUT_SLOTS)()
        total = lib.nf_tarflow_param_layout(ctypes.byref(self._desc), offs)
        if total < 0:
            raise _lib.NfError("nf_tarflow_param_layout: " + lib.nf_last_error().decode(errors="replace"))
        self._offs, self._total = list(offs), int(total)
        self._flat, self._ws, self._stash_pool = None, None, {}
        self._dp_group, self._dp_world, self._dp_avg = None, 1, False
        self._flatten()
        self.register_load_state_dict_post_hook(lambda module, keys: module._flatten())

    # ------------------------------------------------------------------ parameter arena
    def _slot_params(self):
        (bs, pos, win, bin_, wc, bc, layers, ls, ln1w, ln1b, wqkv, bqkv, wproj, bproj, ln2w, ln2b, w1, b1, w2, b2, wout,
         bout) = self._offs
        out = []
        for i, blk in enumerate(self.blocks):
            o = i * bs
            out += [(blk.pos_embed, o + pos), (blk.proj_in.weight, o + win), (blk.proj_in.bias, o + bin_),
                    (blk.cond_proj.weight, o + wc), (blk.cond_proj.bias, o + bc)]
            for j, lay in enumerate(blk.layers):
                q = o + layers + j * ls
                out += [(lay.attn.norm.weight, q + ln1w), (lay.attn.norm.bias, q + ln1b), (lay.attn.qkv.weight, q + wqkv),
                        (lay.attn.qkv.bias, q + bqkv), (lay.attn.proj.weight, q + wproj), (lay.attn.proj.bias, q + bproj),
                        (lay.mlp.norm.weight, q + ln2w), (lay.mlp.norm.bias, q + ln2b), (lay.mlp.fc1.weight, q + w1),
                        (lay.mlp.fc1.bias, q + b1), (lay.mlp.fc2.weight, q + w2), (lay.mlp.fc2.bias, q + b2)]
            out += [(blk.proj_out.weight, o + wout), (blk.proj_out.bias, o + bout)]
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
        self._ws, self._stash_pool = None, {}
        return out

    def _is_flat(self):
        base = self._flat.data_ptr()
        return all(p.data_ptr() == base + 4 * off for p, off in (self._slots[0], self._slots[-1]))

    def _grad_views(self, dflat):
        return [dflat[off:off + p.numel()].view(p.shape) for p, off in self._slots]

    # ------------------------------------------------------------------ plumbing
    def _check(self, x, y):
        A, R = self.in_channels, self.rep_dim
        if not (x.is_cuda and y.is_cuda):
            raise RuntimeError("nf_b200.TransformerRealNVP runs on CUDA only (sm_100a kernels, no CPU fallback)")
        if x.dim() != 3 or x.shape[1] != A or x.shape[2] != 1:
            raise ValueError(f"actions must be (B, {A}, 1), got {tuple(x.shape)}")
        if not ((y.dim() == 3 and y.shape[1] == 1 and y.shape[2] == R) or (y.dim() == 2 and y.shape[1] == R)) \
                or y.shape[0] != x.shape[0]:
            raise ValueError(f"cond must be (B, 1, {R}), got {tuple(y.shape)}")
        if x.dtype != torch.float32 or y.dtype != torch.float32:
            raise TypeError("nf_b200 computes in float32: inputs must be float32")
        if not self._is_flat():
            self._flatten()
        if self._flat.device != x.device or self._flat.dtype != torch.float32 or y.device != x.device:
            raise RuntimeError("module parameters and inputs must be float32 on the same device")

    def _workspace(self, B, dev):
        need = lib.nf_tarflow_workspace_bytes(ctypes.byref(self._desc), B)
        if self._ws is None or self._ws.numel() < need or self._ws.device != dev:
            self._ws = torch.empty(need, dtype=torch.uint8, device=dev)
        return self._ws

    @staticmethod
    def _stream(dev):
        return ctypes.c_void_p(torch._C._cuda_getCurrentRawStream(dev.index if dev.index is not None
                                                                  else torch.cuda.current_device()))

    _MAX_STASH_SLOTS = 2

    def _acquire_stash(self, B, dev):
        pool = self._stash_pool.setdefault((B, dev), [])
        for slot in pool:
            if not slot.busy:
                slot.busy = True
                return slot.tensor, slot
        t = torch.empty(lib.nf_tarflow_stash_bytes(ctypes.byref(self._desc), B), dtype=torch.uint8, device=dev)
        if len(pool) < self._MAX_STASH_SLOTS:
            slot = _StashSlot(t)
            slot.busy = True
            pool.append(slot)
            return t, slot
        return t, None

    def _run_forward(self, x, y, need_stash):
        dev, B, A = x.device, x.shape[0], self.in_channels
        x2, y2 = _dense(x.reshape(B, A)), _dense(y.reshape(B, self.rep_dim))
        with _on(dev):
            ws = self._workspace(B, dev)
            z = torch.empty((B, A, 1), dtype=torch.float32, device=dev)
            ld = torch.empty((B,), dtype=torch.float32, device=dev)
            outs = torch.empty((self.n_blocks, B, A, 1), dtype=torch.float32, device=dev)
            stash, slot = (self._acquire_stash(B, dev) if need_stash else (None, None))
            _lib.check(lib.nf_tarflow_forward(ctypes.byref(self._desc), self._flat.data_ptr(), x2.data_ptr(),
                                              y2.data_ptr(), B, z.data_ptr(), ld.data_ptr(), outs.data_ptr(),
                                              stash.data_ptr() if stash is not None else None, ws.data_ptr(),
                                              ws.numel(), self._stream(dev)), "nf_tarflow_forward")
        return z, ld, outs, stash, slot

    def _run_backward(self, stash, B, dz, dld, need_dx, need_dy, need_p):
        dev, A = self._flat.device, self.in_channels
        dz = _dense(dz.reshape(B, A)) if dz is not None else None
        dld = _dense(dld) if dld is not None else None
        with _on(dev):
            ws = self._workspace(B, dev)
            dx = torch.empty((B, A, 1), dtype=torch.float32, device=dev) if need_dx else None
            dy = torch.empty((B, 1, self.rep_dim), dtype=torch.float32, device=dev) if need_dy else None
            dflat = torch.empty(self._total, dtype=torch.float32, device=dev) if need_p else None
            p = lambda t: t.data_ptr() if t is not None else None      # noqa: E731
            _lib.check(lib.nf_tarflow_backward(ctypes.byref(self._desc), self._flat.data_ptr(), stash.data_ptr(), B,
                                               p(dz), p(dld), p(dx), p(dy), p(dflat), ws.data_ptr(), ws.numel(),
                                               self._stream(dev)), "nf_tarflow_backward")
            if self._dp_world > 1 and need_p:     # (a rank-local evaluation must not enter the collective)
                self._allreduce_grads(dflat)
        return dx, dy, dflat

    # ------------------------------------------------------------------ data parallel
    def enable_data_parallel(self, process_group=None):
        """Batch-sharded data parallel training (rows are independent): parameters are broadcast from rank 0 and
        backward averages the flat gradient arena over the ranks (one NCCL all-reduce) before autograd sees the views."""
        import torch.distributed as dist
        self._dp_group = process_group
        self._dp_world = dist.get_world_size(process_group)
        self._dp_avg = dist.get_backend(process_group) == "nccl"
        if self._dp_world > 1:
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

    # ------------------------------------------------------------------ the reference's call sites
    def forward(self, x: torch.Tensor, cond: torch.Tensor):
        """(z (B, A, 1), outputs [blocks x (B, A, 1)], logdets (B,))   -- torch_nfgcbc.py:141"""
        self._check(x, cond)
        params = [p for p, _ in self._slots]
        need_stash = torch.is_grad_enabled() and (x.requires_grad or cond.requires_grad or
                                                  any(p.requires_grad for p in params))
        nodes = _accumulate_nodes([p for p in params if p.requires_grad]) if need_stash else []
        y = cond if cond.dim() == 3 else cond.unsqueeze(1)
        z, ld, outs = _Flow.apply(self, need_stash, nodes, x, y, *params)
        return z, list(outs.unbind(0)), ld

    @torch.no_grad()
    def reverse(self, z: torch.Tensor, cond: torch.Tensor, return_logdet: bool = False):
        """Sampling direction: the autoregressive loop over the A dimensions of every block, last block first
        (utils/torch_evaluations.py:79-82)."""
        self._check(z, cond)
        dev, B, A = z.device, z.shape[0], self.in_channels
        z2, y2 = _dense(z.reshape(B, A)), _dense(cond.reshape(B, self.rep_dim))
        with _on(dev):
            ws = self._workspace(B, dev)
            x = torch.empty((B, A, 1), dtype=torch.float32, device=dev)
            ld = torch.empty((B,), dtype=torch.float32, device=dev) if return_logdet else None
            _lib.check(lib.nf_tarflow_inverse(ctypes.byref(self._desc), self._flat.data_ptr(), z2.data_ptr(),
                                              y2.data_ptr(), B, x.data_ptr(), ld.data_ptr() if ld is not None else None,
                                              ws.data_ptr(), ws.numel(), self._stream(dev)), "nf_tarflow_inverse")
        return (x, ld) if return_logdet else x
