"""torch.nn.Module front of the B200-native coupling flow.

Mirrors the reference's operator interface for the hot path so the training scripts only
change their import (`from nf_b200 import RealNVP` instead of
`from utils.torch_fcnf import RealNVP`, /root/reference/il-and-cil/torch_nfbc_ant.py:30):

* ``RealNVP(in_channels, channels, cond_channels, n_layers, prior)``      torch_fcnf.py:119-132
* ``forward(x, y) -> (z, outputs, log_dets)``                               torch_fcnf.py:134-141
* ``reverse(x, y, return_sequence=False)``                                  torch_fcnf.py:143-152
* ``.prior``, ``.blocks[i].{l,t,s}``, identical ``state_dict`` keys          torch_fcnf.py:14-31, 61-93

All arithmetic runs in libnf_b200.so (hand-written sm_100a CUDA) through the C ABI of
include/nf_b200.h; there is no CPU or PyTorch-eager fallback -- CPU tensors raise.
"""
from __future__ import annotations

import ctypes
import os
import itertools
import operator
import math
from typing import List, Optional

import numpy as np
import torch
from torch import nn

from . import _lib
from ._lib import lib


_get_grad = operator.attrgetter("grad")
_get_requires_grad = operator.attrgetter("requires_grad")
_get_version = operator.attrgetter("_version")


class _NoSwitch:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


_NO_SWITCH = _NoSwitch()


def _on(dev):
    """``torch.cuda.device(dev)`` only when ``dev`` is not already the current device."""
    if dev.index is None or dev.index == torch.cuda.current_device():
        return _NO_SWITCH
    return torch.cuda.device(dev)


def _dense(t: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    """Contiguous and 16-byte aligned (the C ABI's contract); row slices of a tensor are contiguous
    but can start at any float, so those are copied."""
    if t is None:
        return None
    t = t.contiguous()
    if t.data_ptr() % 16 != 0:
        t = t.clone()
    return t


def count_parameters(model):
    """torch_fcnf.py:11."""
    return sum(p.numel() for p in model.parameters())


class InvertiblePLU(nn.Module):
    """Parameter container of the invertible linear layer W = P L (U + diag s)
    (torch_fcnf.py:14-31: same initialiser, same parameter names, P/P_inv frozen)."""

    def __init__(self, features: int):
        super().__init__()
        self.features = features
        w = torch.empty((features, features))
        nn.init.orthogonal_(w)
        P, L, U = torch.linalg.lu(w)
        self.s = nn.Parameter(torch.diag(U))
        self.U = nn.Parameter(U - torch.diag(self.s))
        self.L = nn.Parameter(L)
        self.P = nn.Parameter(P, requires_grad=False)
        self.P_inv = nn.Parameter(torch.linalg.inv(P), requires_grad=False)


def _conditioner(in_features: int, first_hidden: int, channels: int, out_features: int, ln_eps: float):
    """Linear-LeakyReLU-LayerNorm x2 + Linear with a zero-initialised head (torch_fcnf.py:74-93)."""
    net = nn.Sequential(
        nn.Linear(in_features, first_hidden), nn.LeakyReLU(), nn.LayerNorm(first_hidden, eps=ln_eps),
        nn.Linear(first_hidden, channels), nn.LeakyReLU(), nn.LayerNorm(channels, eps=ln_eps),
        nn.Linear(channels, out_features),
    )
    nn.init.zeros_(net[-1].weight)
    nn.init.zeros_(net[-1].bias)
    return net


class MetaBlock(nn.Module):
    """Parameter container of one flow block (torch_fcnf.py:60-93); the arithmetic of
    forward/reverse (:95-117) is done for the whole stack by RealNVP."""

    def __init__(self, in_channels: int, channels: int, cond_channels: int,
                 first_hidden: Optional[int] = None, ln_eps: float = 1e-5):
        super().__init__()
        k1 = int(np.ceil(in_channels / 2)) + cond_channels
        h1 = channels if first_hidden is None else first_hidden
        self.l = InvertiblePLU(in_channels)
        self.t = _conditioner(k1, h1, channels, in_channels // 2, ln_eps)
        self.s = _conditioner(k1, h1, channels, in_channels // 2, ln_eps)


class _StashSlot:
    """One persistent forward->backward stash buffer.  The library's launch graphs are keyed on the
    buffers the kernels touch, so training keeps reusing the same few stashes instead of allocating a
    fresh one per forward; a slot is busy from its forward until its backward ran (or its autograd
    graph was dropped)."""
    __slots__ = ("tensor", "busy", "__weakref__")

    def __init__(self, tensor):
        self.tensor = tensor
        self.busy = False


class _SlotGuard:
    """Releases the slot when the autograd context that owns it goes away without a backward."""

    def __init__(self, slot):
        self.slot = slot

    def release(self):
        if self.slot is not None:
            self.slot.busy = False
            self.slot = None

    def __del__(self):
        self.release()


def _accumulate_nodes(params):
    """AccumulateGrad nodes of `params` (call with grad mode on, outside Function.forward)."""
    nodes = []
    for p in params:
        fn = p.view_as(p).grad_fn
        if fn is not None:
            nodes.append(fn.next_functions[0][0])
    return nodes


def _param_grads_wanted(needs, nodes):
    """``ctx.needs_input_grad`` is fixed at forward time from ``requires_grad``; it does not know the ``inputs=`` of
    ``torch.autograd.grad``.  The reference's evaluation-time denoise step (torch_nfbc_ant.py:212-217) differentiates
    w.r.t. the action only while the parameters still require grad: ask the engine whether this backward pass will
    run the parameters' AccumulateGrad nodes, so that step neither pays for the weight-gradient kernels nor leaves
    ``p.grad`` behind (nor issues the data-parallel all-reduce on one rank only)."""
    if not any(needs):
        return False
    if not nodes:
        return True
    try:
        return any(torch._C._will_engine_execute_node(n) for n in nodes)
    except RuntimeError:
        return True      # autograd.grad with the parameters themselves as inputs: leaf queries are unsupported there


_LIB_COMM = {}     # (device index, group id) -> world size of the library's NCCL communicator


def _init_library_comm(process_group, device):
    """Open the library's own NCCL communicator on `device` for the ranks of `process_group` (once per process)."""
    import torch.distributed as dist
    key = device.index if device.index is not None else torch.cuda.current_device()
    world, rank = dist.get_world_size(process_group), dist.get_rank(process_group)
    if _LIB_COMM.get(key) == world and lib.nf_dp_world() == world:
        return
    with _on(device):
        ident = torch.zeros(128, dtype=torch.uint8, device=device)
        if rank == 0:
            buf = ctypes.create_string_buffer(128)
            _lib.check(lib.nf_dp_unique_id(buf), "nf_dp_unique_id")
            ident.copy_(torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8))
        dist.broadcast(ident, src=dist.get_global_rank(process_group, 0) if process_group is not None else 0,
                       group=process_group)
        raw = bytes(ident.cpu().numpy().tobytes())
        torch.cuda.synchronize(device)
        _lib.check(lib.nf_dp_init(ctypes.c_char_p(raw), rank, world), "nf_dp_init")
        _connect_peer_memory(process_group, device, rank, world)
    _LIB_COMM[key] = world


# Largest gradient arena (bytes) averaged by the library's peer-memory kernel instead of NCCL; 0 turns the path off.
_P2P_MAX_BYTES = int(os.environ.get("NF_B200_DP_P2P_MAX_BYTES", str(16 << 20)))


def _connect_peer_memory(process_group, device, rank, world):
    """All-reduce over NVLink peer memory (csrc/dp.cu): every rank allocates its exchange buffer, the 64-byte CUDA IPC
    handles travel through ``torch.distributed``, every rank maps its peers' buffers.  The path is switched on only when
    ALL ranks connected; otherwise (more than 8 ranks, ranks on different boxes, IPC refused) NCCL keeps the job."""
    import torch.distributed as dist
    if _P2P_MAX_BYTES <= 0 or world < 2 or world > 8:
        return False
    mine = torch.zeros(65, dtype=torch.uint8, device=device)
    buf = ctypes.create_string_buffer(64)
    if lib.nf_dp_p2p_alloc(_P2P_MAX_BYTES // 4, buf) == 0:
        mine[:64].copy_(torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8))
        mine[64] = 1
    everyone = torch.zeros(world, 65, dtype=torch.uint8, device=device)
    dist.all_gather_into_tensor(everyone, mine, group=process_group)
    everyone = everyone.cpu()
    ok = bool(everyone[:, 64].min().item())
    if ok:
        handles = everyone[:, :64].contiguous().numpy().tobytes()
        ok = lib.nf_dp_p2p_connect(ctypes.c_char_p(handles), rank, world) == 0
    agreed = torch.tensor([1 if ok else 0], dtype=torch.int32, device=device)
    dist.all_reduce(agreed, op=dist.ReduceOp.MIN, group=process_group)
    ok = bool(agreed.item())
    if ok:
        _lib.check(lib.nf_dp_p2p_enable(1), "nf_dp_p2p_enable")
    return ok


class _Flow(torch.autograd.Function):
    """autograd bridge: forward = nf_flow_forward, backward = nf_flow_backward.

    ``params`` is either every trainable parameter (``fast_param_grads=False``: gradients are returned
    to autograd, one AccumulateGrad node per parameter) or a single carrier parameter
    (``fast_param_grads=True``: backward writes the flat gradient arena and sets ``p.grad`` views itself)."""

    @staticmethod
    def forward(ctx, model: "RealNVP", fast: bool, nodes, x, y, *params):
        z, logdet, outs, stash, packed, slot = model._run_forward(x, y, True, want_outputs=model.return_outputs)
        if outs is None:
            outs = z.new_zeros(())
        ctx.model = model
        ctx.fast = fast
        ctx.nodes = nodes
        ctx.stash = stash
        ctx.guard = _SlotGuard(slot) if slot is not None else None
        ctx.packed = packed
        ctx.pver = model._pver_seen      # computed by _prepared inside _run_forward
        ctx.cver = model._quick_version()
        ctx.B = x.shape[0]
        ctx.has_y = y is not None
        ctx.save_for_backward(*([y] if y is not None else []))
        ctx.mark_non_differentiable(outs)
        return z, logdet, outs

    @staticmethod
    def backward(ctx, dz, dld, _douts):
        model = ctx.model
        if ctx.stash is None:
            raise RuntimeError("nf_b200: backward called twice on the same forward (the stash is freed after use)")
        # (an optimizer step between forward and backward touches every trainable parameter, the first one included: one
        # version read instead of a second walk over all of them)
        if ctx.cver != model._quick_version():
            raise RuntimeError("nf_b200: parameters were modified between forward and backward")
        y = ctx.saved_tensors[0] if ctx.has_y else None
        need_x = ctx.needs_input_grad[3]
        need_y = ctx.needs_input_grad[4]
        need_p = _param_grads_wanted(ctx.needs_input_grad[5:], ctx.nodes)
        nparams = len(ctx.needs_input_grad) - 5
        if ctx.fast:
            dx, dy = model._run_backward_fast(ctx.stash, ctx.packed, y, ctx.B, dz, dld, need_x, need_y, need_p)
            pg = (None,) * nparams
        else:
            dx, dy, dflat = model._run_backward(ctx.stash, ctx.packed, y, ctx.B, dz, dld, need_x, need_y, need_p)
            if need_p:
                views = model._grad_views(dflat)
                pg = tuple(v if need else None for v, need in zip(views, ctx.needs_input_grad[5:]))
            else:
                pg = (None,) * nparams
        ctx.stash = None
        if ctx.guard is not None:
            ctx.guard.release()      # the stash may be reused by the next forward
        return (None, None, None, dx, dy) + pg


class _FlowReverse(torch.autograd.Function):
    """autograd bridge of the sampling direction: forward = nf_flow_inverse; backward = nf_flow_forward on the result
    (for the stash) followed by nf_flow_inverse_backward."""

    @staticmethod
    def forward(ctx, model: "RealNVP", fast: bool, nodes, want_logdet: bool, z, y, *params):
        out, logdet, _ = model._run_inverse(z, y, want_logdet)
        ctx.model, ctx.fast, ctx.nodes = model, fast, nodes
        ctx.pver = model._pver_seen
        ctx.has_y = y is not None
        ctx.save_for_backward(out, *([y] if y is not None else []))
        if logdet is None:
            logdet = out.new_zeros(())
            ctx.mark_non_differentiable(logdet)
        return out, logdet

    @staticmethod
    def backward(ctx, dx, dld):
        model = ctx.model
        if ctx.pver != model._params_version():
            raise RuntimeError("nf_b200: parameters were modified between reverse() and its backward")
        saved = ctx.saved_tensors
        out, y = saved[0], (saved[1] if ctx.has_y else None)
        B = out.shape[0]
        if dld is not None and dld.dim() == 0:
            dld = None
        need_z, need_y = ctx.needs_input_grad[4], ctx.needs_input_grad[5]
        need_p = _param_grads_wanted(ctx.needs_input_grad[6:], ctx.nodes)
        nparams = len(ctx.needs_input_grad) - 6
        _, _, _, stash, packed, slot = model._run_forward(out, y, True, want_outputs=False)
        try:
            if ctx.fast:
                dz, dy = model._run_backward_fast(stash, packed, y, B, dx, dld, need_z, need_y, need_p, inverse=True)
                pg = (None,) * nparams
            else:
                dz, dy, dflat = model._run_backward(stash, packed, y, B, dx, dld, need_z, need_y, need_p, inverse=True)
                if need_p:
                    views = model._grad_views(dflat)
                    pg = tuple(v if need else None for v, need in zip(views, ctx.needs_input_grad[6:]))
                else:
                    pg = (None,) * nparams
        finally:
            if slot is not None:
                slot.busy = False
        return (None, None, None, None, dz, dy) + pg


class _FlowNLL(torch.autograd.Function):
    """loss = -mean(log N(z; 0, I) + log_dets) with the loss and its cotangents computed by the library
    (nf_nll_forward / nf_nll_cotangents) instead of ~8 elementwise torch kernels either side of the flow."""

    @staticmethod
    def forward(ctx, model: "RealNVP", nodes, x, y, *params):
        z, logdet, _, stash, packed, slot = model._run_forward(x, y, True, want_outputs=False)
        ctx.nodes = nodes
        loss = torch.empty((), dtype=torch.float32, device=x.device)
        _lib.check(lib.nf_nll_forward(z.data_ptr(), logdet.data_ptr(), x.shape[0], model.in_channels,
                                      loss.data_ptr(), model._stream(x.device)), "nf_nll_forward")
        ctx.model, ctx.stash, ctx.packed = model, stash, packed
        ctx.guard = _SlotGuard(slot) if slot is not None else None
        ctx.pver = model._pver_seen      # computed by _prepared inside _run_forward
        ctx.B = x.shape[0]
        ctx.has_y = y is not None
        ctx.save_for_backward(z, *([y] if y is not None else []))
        return loss

    @staticmethod
    def backward(ctx, g):
        model = ctx.model
        if ctx.stash is None:
            raise RuntimeError("nf_b200: backward called twice on the same forward (the stash is freed after use)")
        if ctx.pver != model._params_version():
            raise RuntimeError("nf_b200: parameters were modified between forward and backward")
        z = ctx.saved_tensors[0]
        y = ctx.saved_tensors[1] if ctx.has_y else None
        B, D = ctx.B, model.in_channels
        dz = torch.empty_like(z)
        dld = torch.empty((B,), dtype=torch.float32, device=z.device)
        g = _dense(g.reshape(1).to(torch.float32))
        _lib.check(lib.nf_nll_cotangents(z.data_ptr(), g.data_ptr(), B, D, dz.data_ptr(), dld.data_ptr(),
                                         model._stream(z.device)), "nf_nll_cotangents")
        need_x, need_y = ctx.needs_input_grad[2], ctx.needs_input_grad[3]
        need_p = _param_grads_wanted(ctx.needs_input_grad[4:], ctx.nodes)
        dx, dy = model._run_backward_fast(ctx.stash, ctx.packed, y, B, dz, dld, need_x, need_y, need_p)
        ctx.stash = None
        if ctx.guard is not None:
            ctx.guard.release()
        return (None, None, dx, dy) + (None,) * (len(ctx.needs_input_grad) - 4)


class RealNVP(nn.Module):
    """Drop-in for the reference ``RealNVP`` (torch_fcnf.py:119-152).

    Extra keyword arguments select the flax twin's semantics
    (/root/reference/unsupservised-gcrl/nf.py): ``first_hidden=channels+cond_channels`` (:94,101),
    ``ln_eps=1e-6`` (flax LayerNorm default), ``ln_fast_variance=True`` (flax computes the LayerNorm variance as
    ``max(0, E[x^2] - E[x]^2)``), ``return_outputs=False`` (``forward`` returns the 2-tuple ``(z, log_dets)`` of
    nf.py:174-179 and no per-block outputs are materialised), and ``precision``:
    ``"fp32"`` (tcgen05 3xTF32, fp32-parity, the default), ``"fp32_simt"`` (FFMA checker) or ``"tf32"`` -- the optional
    reduced-precision mode: one TF32 tensor-core pass on the raw fp32 operands (10-bit mantissa, fp32 accumulation),
    contract 2e-2 relative on z / log_prob for D <= 64 and n_layers <= 12.  A bf16 mode is not built.

    ``fast_param_grads`` (default on): ``loss.backward()`` writes all parameter gradients into one
    persistent flat arena with a single library call and points every ``p.grad`` at its view of that
    arena (accumulating if gradients are already present), instead of handing ~24 tensors per block
    to autograd one AccumulateGrad node at a time -- the same trade DistributedDataParallel's
    ``gradient_as_bucket_view`` makes.  Gradients w.r.t. ``x`` and ``y`` always flow through autograd.
    Turn it off for ``torch.autograd.grad(..., model.parameters())`` or parameter hooks.
    """

    def __init__(self, in_channels, channels, cond_channels, n_layers, prior, *,
                 first_hidden: Optional[int] = None, ln_eps: float = 1e-5, precision: str = "fp32",
                 fast_param_grads: bool = True, ln_fast_variance: bool = False, return_outputs: bool = True):
        super().__init__()
        self.fast_param_grads = bool(fast_param_grads)
        self.ln_fast_variance = bool(ln_fast_variance)
        self.return_outputs = bool(return_outputs)
        self.in_channels = int(in_channels)
        self.channels = int(channels)
        self.cond_channels = int(cond_channels)
        self.n_layers = int(n_layers)
        self.first_hidden = int(channels if first_hidden is None else first_hidden)
        self.ln_eps = float(ln_eps)
        if precision not in _lib.PREC:
            raise ValueError(f"precision must be one of {sorted(_lib.PREC)}")
        self.precision = precision
        self.blocks = nn.ModuleList(
            [MetaBlock(in_channels, channels, cond_channels, first_hidden, ln_eps) for _ in range(n_layers)])
        self.prior = prior

        self._desc = _lib.make_desc(self.in_channels, self.channels, self.first_hidden, self.cond_channels,
                                    self.n_layers, self.ln_eps, _lib.PREC[precision],
                                    _lib.NF_FLAG_LN_FAST_VARIANCE if self.ln_fast_variance else 0)
        self._layout = _lib.param_layout(self._desc)
        self._flat: Optional[torch.Tensor] = None
        self._packed: Optional[torch.Tensor] = None
        self._packed_key = None
        self._ws: Optional[torch.Tensor] = None
        self._gflat: Optional[torch.Tensor] = None
        self._gviews = None
        self._stash_pool = {}
        self._neg_b = None
        self._pver_seen = 0
        self._xver = 0
        self._dp_group = None
        self._dp_world = 1
        self._dp_avg = False
        self._dp_in_graph = False
        self._noise_offset = {}
        self._flatten()
        self.register_load_state_dict_post_hook(lambda module, keys: module._flatten())

    # ------------------------------------------------------------------ parameter arena
    def _slot_params(self):
        """[(param, offset_in_floats)] in arena order (include/nf_b200.h NF_P_*)."""
        lay = self._layout
        out = []
        for i, blk in enumerate(self.blocks):
            base = i * lay.block_stride
            plist = [blk.l.s, blk.l.U, blk.l.L, blk.l.P, blk.l.P_inv]
            for net in (blk.t, blk.s):
                plist += [net[0].weight, net[0].bias, net[2].weight, net[2].bias, net[3].weight, net[3].bias,
                          net[5].weight, net[5].bias, net[6].weight, net[6].bias]
            for slot, p in enumerate(plist):
                assert p.numel() == lay.slot_numel[slot], (slot, tuple(p.shape), lay.slot_numel[slot])
                out.append((p, base + lay.slot_off[slot]))
        return out

    def _flatten(self):
        """(Re)pack every parameter into one flat fp32 arena and make the nn.Parameters views of it.
        The arena layout is the one the C library computes (nf_param_layout)."""
        sp = self._slot_params()
        dev = sp[0][0].device
        dtype = sp[0][0].dtype
        flat = torch.zeros(self._layout.total, dtype=dtype, device=dev)
        with torch.no_grad():
            for p, off in sp:
                view = flat[off:off + p.numel()].view(p.shape)
                view.copy_(p.detach())
                p.data = view
        self._flat = flat
        self._slots = sp
        self._trainable = [(p, off) for p, off in sp if p.requires_grad]
        # one split_with_sizes call yields a 1-D view per arena segment (slots and the gaps between them)
        sizes, picks, pos = [], [], 0
        for p, off in sp:
            if off > pos:
                sizes.append(off - pos)
            if p.requires_grad:
                picks.append((len(sizes), tuple(p.shape) if p.dim() != 1 else None))
            sizes.append(p.numel())
            pos = off + p.numel()
        if pos < self._layout.total:
            sizes.append(self._layout.total - pos)
        self._split_sizes, self._split_picks = sizes, picks
        self._tparams = [p for p, _ in self._trainable]
        self._carrier_cache = [None, None]    # [carrier parameter, its AccumulateGrad nodes] (a list: nn.Module would
                                              # register a Parameter assigned as an attribute)
        self._ws_need = {}
        self._packed = None
        self._packed_key = None
        self._gflat = None
        self._gviews = None

    def _apply(self, fn, recurse=True):
        out = super()._apply(fn, recurse)
        self._flatten()
        self._ws = None
        self._stash_pool = {}
        return out

    def _is_flat(self):
        p0, off0 = self._slots[0]
        p1, off1 = self._slots[-1]
        base = self._flat.data_ptr()
        return (p0.data_ptr() == base + 4 * off0 and p1.data_ptr() == base + 4 * off1
                and p0.device == self._flat.device)

    def _bump_version(self):
        """Called by code that updates the arena outside torch's version counters (FlatAdamW)."""
        self._xver += 1

    def _quick_version(self):
        tp = self._tparams
        return self._xver + (tp[0]._version + tp[-1]._version if tp else 0)

    def _params_version(self):
        return self._xver + sum(map(_get_version, self._tparams))     # (C-level walk: ~150 parameters per forward)

    def _grad_carriers(self):
        """The trainable parameters handed to the autograd function: all of them, or -- ``fast_param_grads`` -- only the
        first one (the carrier whose AccumulateGrad node tells backward whether parameter gradients are wanted).  The
        carrier is cached and re-validated with one attribute read instead of a walk over ~150 parameters per forward;
        its AccumulateGrad node is cached too (holding it keeps it the node autograd uses)."""
        if not self.fast_param_grads:
            params = [p for p in self._tparams if p.requires_grad]
            return params, _accumulate_nodes(params)
        cache = self._carrier_cache
        c = cache[0]
        if c is None or not c.requires_grad:
            c = cache[0] = next((p for p in self._tparams if p.requires_grad), None)
            cache[1] = None
        if c is None:
            return [], []
        if cache[1] is None:
            cache[1] = _accumulate_nodes([c])
        return [c], cache[1]

    def _grad_views(self, dflat):
        chunks = dflat.split_with_sizes(self._split_sizes)
        return [chunks[i] if shape is None else chunks[i].view(shape) for i, shape in self._split_picks]

    # ------------------------------------------------------------------ plumbing
    def _check_inputs(self, x, y):
        """``y=None`` means "the conditioning input is identically zero" (the fast path of
        unsupservised-gcrl/train_auto_NF_rl.py:328-339, which evaluates the flow with ``y = zeros`` a million times
        per actor step): the y part of the first Linear is skipped instead of multiplying zeros."""
        if not (x.is_cuda and (y is None or y.is_cuda)):
            raise RuntimeError("nf_b200.RealNVP runs on CUDA only (sm_100a kernels, no CPU fallback); "
                               "move the module and its inputs to a B200")
        if x.dim() != 2 or x.shape[1] != self.in_channels:
            raise ValueError(f"x must be (B, {self.in_channels}), got {tuple(x.shape)}")
        if y is not None and (y.dim() != 2 or y.shape[1] != self.cond_channels or y.shape[0] != x.shape[0]):
            raise ValueError(f"y must be ({x.shape[0]}, {self.cond_channels}), got {tuple(y.shape)}")
        if x.dtype != torch.float32 or (y is not None and y.dtype != torch.float32):
            raise TypeError("nf_b200 computes in float32: x and y must be float32")
        if not self._is_flat():
            self._flatten()
        if self._flat.device != x.device or self._flat.dtype != torch.float32:
            raise RuntimeError("module parameters must be float32 on the same device as the inputs")

    def _stream(self, device):
        # the raw handle of torch's current stream on that device (what current_stream().cuda_stream returns,
        # without building a Stream object: this runs several times per step)
        return ctypes.c_void_p(torch._C._cuda_getCurrentRawStream(device.index if device.index is not None
                                                                  else torch.cuda.current_device()))

    def _workspace(self, B, device):
        need = self._ws_need.get(B)
        if need is None:
            need = self._ws_need[B] = lib.nf_flow_workspace_bytes(ctypes.byref(self._desc), B)
        if self._ws is None or self._ws.numel() < need or self._ws.device != device:
            self._ws = torch.empty(need, dtype=torch.uint8, device=device)
        return self._ws

    def _prepared(self, what, device):
        """Derived parameter tables (W, W^-1, folded LayerNorm affine, log|det|), rebuilt whenever a
        parameter changed (torch_fcnf.py:34-38, 45-54 redo this in every call).  The buffer is
        persistent, so the library's cached launch graphs keep seeing the same pointer."""
        self._pver_seen = pver = self._params_version()   # (the autograd functions reuse it: one walk over the parameters per forward)
        key = (pver, self._flat.data_ptr())
        have = self._packed_key
        if self._packed is not None and have is not None and have[0] == key and (have[1] & what) == what:
            return self._packed
        if self._packed is None or self._packed.device != device:
            nbytes = lib.nf_flow_packed_bytes(ctypes.byref(self._desc))
            self._packed = torch.empty(nbytes, dtype=torch.uint8, device=device)
        packed = self._packed
        ws = self._workspace(1, device)
        if have is not None and have[0] == key:
            what |= have[1]
        _lib.check(lib.nf_flow_prepare(ctypes.byref(self._desc), self._flat.data_ptr(), packed.data_ptr(), what,
                                       ws.data_ptr(), ws.numel(), self._stream(device)), "nf_flow_prepare")
        self._packed_key = (key, what)
        return packed

    def _run_forward(self, x, y, want_stash, want_outputs=True):
        dev = x.device
        B, D, n = x.shape[0], self.in_channels, self.n_layers
        x = _dense(x)
        y = _dense(y)
        yp = y.data_ptr() if y is not None else None
        with _on(dev):
            packed = self._prepared(_lib.NF_PREP_FWD, dev)
            ws = self._workspace(B, dev)
            z = torch.empty((B, D), dtype=torch.float32, device=dev)
            logdet = torch.empty((B,), dtype=torch.float32, device=dev)
            stash, slot = None, None
            if want_stash:
                stash, slot = self._acquire_stash(B, dev)
            # `outputs` are independent tensors as in the reference (torch_fcnf.py:134-141): the pooled stash is
            # reused by the next forward, so they are copied out by the library's trailing copy kernel (the one
            # that also delivers z and log_dets -- no extra launch)
            outs, outs_ptr = None, None
            if want_outputs:
                outs = torch.empty((n, B, D), dtype=torch.float32, device=dev)
                outs_ptr = outs.data_ptr()
            _lib.check(lib.nf_flow_forward(
                ctypes.byref(self._desc), self._flat.data_ptr(), packed.data_ptr(), x.data_ptr(), yp, B,
                z.data_ptr(), logdet.data_ptr(), outs_ptr, stash.data_ptr() if stash is not None else None,
                ws.data_ptr(), ws.numel(), self._stream(dev)), "nf_flow_forward")
        return z, logdet, outs, stash, packed, slot

    _MAX_STASH_SLOTS = 4

    def _acquire_stash(self, B, dev):
        """A free persistent stash for batch size B (or a fresh, unpooled one when all are in flight)."""
        pool = self._stash_pool.setdefault((B, dev), [])
        for slot in pool:
            if not slot.busy:
                slot.busy = True
                return slot.tensor, slot
        nbytes = lib.nf_flow_stash_bytes(ctypes.byref(self._desc), B)
        t = torch.empty(nbytes // 4, dtype=torch.float32, device=dev)
        if len(pool) < self._MAX_STASH_SLOTS:
            slot = _StashSlot(t)
            slot.busy = True
            pool.append(slot)
            return t, slot
        return t, None

    def _run_backward(self, stash, packed, y, B, dz, dld, need_x, need_y, need_p, out=None, inverse=False):
        """``inverse=True``: backward of the reverse direction (nf_flow_inverse_backward); dz / dld are then the
        cotangents of the reverse output and reverse log-det, and the returned dx is the gradient w.r.t. its input."""
        dev = stash.device
        D, R = self.in_channels, self.cond_channels
        need_y = need_y and y is not None
        with _on(dev):
            ws = self._workspace(B, dev)
            dz, dld, y = _dense(dz), _dense(dld), _dense(y)
            dx = torch.empty((B, D), dtype=torch.float32, device=dev) if need_x else None
            dy = torch.empty((B, R), dtype=torch.float32, device=dev) if need_y else None
            dflat = None
            if need_p:
                dflat = out if out is not None else torch.empty(self._layout.total, dtype=torch.float32, device=dev)
            ptr = lambda t: t.data_ptr() if t is not None else None
            if inverse:
                _lib.check(lib.nf_flow_inverse_backward(
                    ctypes.byref(self._desc), self._flat.data_ptr(), packed.data_ptr(), stash.data_ptr(), ptr(y),
                    B, ptr(dz), ptr(dld), ptr(dx), ptr(dy), ptr(dflat), ws.data_ptr(), ws.numel(), self._stream(dev)),
                    "nf_flow_inverse_backward")
                if need_p and self._dp_world > 1:
                    if getattr(self, "_dp_in_graph", False):
                        _lib.check(lib.nf_dp_allreduce(dflat.data_ptr(), dflat.numel(), 1, self._stream(dev)), "nf_dp_allreduce")
                    else:
                        self._allreduce_grads(dflat)
            elif need_p and self._dp_world > 1 and getattr(self, "_dp_in_graph", False):
                bucket = int(self.dp_bucket_blocks)
                if bucket <= 0:    # default: one bucket while the all-reduce is latency-bound, four for large arenas
                    bucket = self.n_layers if self._layout.total * 4 < (32 << 20) else max(1, -(-self.n_layers // 4))
                _lib.check(lib.nf_flow_backward_dp(
                    ctypes.byref(self._desc), self._flat.data_ptr(), packed.data_ptr(), stash.data_ptr(), ptr(y),
                    B, ptr(dz), ptr(dld), ptr(dx), ptr(dy), dflat.data_ptr(), bucket, ws.data_ptr(), ws.numel(),
                    self._stream(dev)), "nf_flow_backward_dp")
            elif need_p and self._dp_world > 1 and int(self.dp_bucket_blocks) > 0:
                self._backward_overlapped(stash, packed, y, B, dz, dld, dx, dy, dflat, ws, dev)
            else:
                _lib.check(lib.nf_flow_backward(
                    ctypes.byref(self._desc), self._flat.data_ptr(), packed.data_ptr(), stash.data_ptr(), ptr(y),
                    B, ptr(dz), ptr(dld), ptr(dx), ptr(dy), ptr(dflat), ws.data_ptr(), ws.numel(), self._stream(dev)),
                    "nf_flow_backward")
                if need_p and self._dp_world > 1:
                    self._allreduce_grads(dflat)
        return dx, dy, dflat

    def _run_backward_fast(self, stash, packed, y, B, dz, dld, need_x, need_y, need_p, inverse=False):
        """Backward with the parameter gradients written straight into the persistent flat arena."""
        dev = stash.device
        tp = self._tparams
        if not need_p:
            dx, dy, _ = self._run_backward(stash, packed, y, B, dz, dld, need_x, need_y, False, inverse=inverse)
            return dx, dy
        if self._gflat is None or self._gflat.device != dev:
            self._gflat = torch.zeros(self._layout.total, dtype=torch.float32, device=dev)
            self._gviews = self._grad_views(self._gflat)
        views = self._gviews
        # Which state are the .grad fields in?  fresh: every one is None (after zero_grad(set_to_none=True): write the arena
        # in place and hand out the views); ours: every one already is our view (accumulate with one flat add); anything
        # else goes parameter by parameter.  Both common states are recognised with C-level identity tests -- a Python
        # loop over ~150 parameters with two attribute reads each was 30 us of host time per step.
        grads = list(map(_get_grad, tp))
        n_none = sum(map(operator.is_, grads, itertools.repeat(None)))
        fresh = n_none == len(grads)
        ours = False
        if not fresh and n_none == 0:
            ours = sum(map(operator.is_, grads, views)) == len(grads)
            if not ours:      # same storage through another Python object?
                ours = all(g.data_ptr() == v.data_ptr() and g.shape == v.shape for g, v in zip(grads, views))
        all_trainable = sum(map(_get_requires_grad, tp)) == len(tp)
        if not all_trainable and not fresh:
            ours = False      # parameters frozen after construction: take the per-parameter path
        target = self._gflat if fresh else None
        dx, dy, dflat = self._run_backward(stash, packed, y, B, dz, dld, need_x, need_y, True, out=target, inverse=inverse)
        if fresh:
            for p, v in zip(tp, views):
                if all_trainable or p.requires_grad:
                    p.grad = v
        elif ours:
            self._gflat.add_(dflat)
        else:
            for p, v in zip(tp, self._grad_views(dflat)):
                if not p.requires_grad:
                    continue
                if p.grad is None:
                    p.grad = v.clone()
                else:
                    p.grad.add_(v)
        return dx, dy

    # ------------------------------------------------------------------ data parallel
    def enable_data_parallel(self, process_group=None, in_graph=None):
        """Batch-sharded data parallel training: gradients of the flat arena are averaged over the ranks of
        `process_group` inside backward.

        With the NCCL backend on CUDA the library opens its own communicator (``nf_dp_init``; the 128-byte unique id
        travels through ``torch.distributed``) and the averaging runs INSIDE the backward launch graph
        (``nf_flow_backward_dp``): blocks finish from the top down, every bucket of ``dp_bucket_blocks`` blocks is
        all-reduced on a side stream while the next bucket computes.  Other backends (gloo on CPU tests) use one
        ``torch.distributed.all_reduce`` after the backward call.  Replicas start from rank 0's parameters."""
        import torch.distributed as dist
        self._dp_group = process_group
        self._dp_world = dist.get_world_size(process_group)
        self._dp_avg = dist.get_backend(process_group) == "nccl"     # NCCL averages inside the reduction
        self._dp_in_graph = False
        if self._dp_world > 1:
            if not self._is_flat():
                self._flatten()
            want = self._dp_avg and self._flat.is_cuda if in_graph is None else bool(in_graph)
            if want:
                _init_library_comm(process_group, self._flat.device)
                self._dp_in_graph = True
            with torch.no_grad():
                dist.broadcast(self._flat, src=dist.get_global_rank(process_group, 0) if process_group is not None else 0,
                               group=process_group)
            self._packed_key = None
            self._bump_version()
        return self

    def _allreduce_grads(self, dflat):
        import torch.distributed as dist
        if self._dp_avg:
            dist.all_reduce(dflat, op=dist.ReduceOp.AVG, group=self._dp_group)
        else:
            dist.all_reduce(dflat, op=dist.ReduceOp.SUM, group=self._dp_group)
            dflat.mul_(1.0 / self._dp_world)

    # Blocks per gradient bucket; 0 = one bucket = one all-reduce of the whole arena at the end of backward.
    # Measured on 2 B200s (bench.py, C2a, 4096 rows per GPU, single-GPU step 0.72 ms): in-graph, one bucket 0.783 ms;
    # in-graph, 2 / 3 buckets 0.81 / 0.85 ms; torch.distributed.all_reduce after the backward call 0.80 ms; per-block
    # buckets through range calls 1.08 ms.  A 4 MB all-reduce is latency-bound (~0.06 ms), so the LAST bucket's
    # all-reduce is exposed at full price however the arena is cut, and every earlier one competes with the backward
    # kernels for SMs: overlap only pays for arenas large against the all-reduce latency (C5: 259 MB).
    dp_bucket_blocks = 0

    def _backward_overlapped(self, stash, packed, y, B, dz, dld, dx, dy, dflat, ws, dev):
        """Data-parallel backward: the pass runs from the top block down in ranges of ``dp_bucket_blocks`` blocks
        (nf_flow_backward_range) and the gradient slice of each finished range is all-reduced asynchronously
        (NCCL on its own stream, averaging in the reduction) while the next range computes; the arena is laid
        out by block, so a range's gradients are one contiguous bucket."""
        import torch.distributed as dist
        ptr = lambda t: t.data_ptr() if t is not None else None
        n, stride = self.n_layers, self._layout.block_stride
        avg = dist.get_backend(self._dp_group) == "nccl"
        handles = []
        end = n
        bucket = int(self.dp_bucket_blocks)
        while end > 0:
            begin = max(0, end - bucket)
            _lib.check(lib.nf_flow_backward_range(
                ctypes.byref(self._desc), self._flat.data_ptr(), packed.data_ptr(), stash.data_ptr(), ptr(y), B,
                ptr(dz), ptr(dld), ptr(dx), ptr(dy), dflat.data_ptr(), begin, end, ws.data_ptr(), ws.numel(),
                self._stream(dev)), "nf_flow_backward_range")
            seg = dflat[begin * stride:end * stride]
            handles.append(dist.all_reduce(seg, op=dist.ReduceOp.AVG if avg else dist.ReduceOp.SUM,
                                           group=self._dp_group, async_op=True))
            end = begin
        for h in handles:
            h.wait()
        if not avg:
            dflat.mul_(1.0 / self._dp_world)

    # ------------------------------------------------------------------ reference API
    def forward(self, x, y=None):
        """-> (z, outputs, log_dets): torch_fcnf.py:134-141.  ``outputs`` (list of n_layers (B, D)
        tensors, the block outputs logged at torch_nfbc_ant.py:286-287) are not differentiable.  ``y=None``: y == 0."""
        self._check_inputs(x, y)
        params, nodes = self._grad_carriers() if torch.is_grad_enabled() else ([], [])
        need_grad = torch.is_grad_enabled() and (x.requires_grad or (y is not None and y.requires_grad) or len(params) > 0)
        if need_grad:
            z, logdet, outs = _Flow.apply(self, self.fast_param_grads, nodes, x, y, *params)
        else:
            z, logdet, outs, _, _, _ = self._run_forward(x, y, False, want_outputs=self.return_outputs)
        if not self.return_outputs:
            return z, logdet                     # flax signature (nf.py:174-179)
        return z, list(outs.unbind(0)), logdet

    def _run_inverse(self, x, y, want_logdet, want_seq=False):
        dev = x.device
        B, D, n = x.shape[0], self.in_channels, self.n_layers
        x = _dense(x.detach())
        y = _dense(y.detach()) if y is not None else None
        with _on(dev):
            packed = self._prepared(_lib.NF_PREP_FWD | _lib.NF_PREP_INV, dev)
            ws = self._workspace(B, dev)
            out = torch.empty((B, D), dtype=torch.float32, device=dev)
            logdet = torch.empty((B,), dtype=torch.float32, device=dev) if want_logdet else None
            seq = torch.empty((n + 1, B, D), dtype=torch.float32, device=dev) if want_seq else None
            _lib.check(lib.nf_flow_inverse(
                ctypes.byref(self._desc), self._flat.data_ptr(), packed.data_ptr(), x.data_ptr(),
                y.data_ptr() if y is not None else None, B,
                out.data_ptr(), logdet.data_ptr() if logdet is not None else None,
                seq.data_ptr() if seq is not None else None, ws.data_ptr(), ws.numel(), self._stream(dev)),
                "nf_flow_inverse")
        return out, logdet, seq

    def reverse(self, x, y=None, return_sequence=False, return_logdet=False):
        """torch_fcnf.py:143-152 (sampling direction).  With ``return_logdet=True`` also returns the flax-style
        reverse log-det (nf.py:131-148).  ``y=None``: y == 0.

        Differentiable like the reference's (w.r.t. the input, ``y`` and every trainable parameter; both directions
        are differentiated by offline-rl/agents/vinf.py:50,68): when gradients are enabled and something requires
        them, the result carries an autograd node whose backward re-runs the forward direction on the result to obtain
        the activations (the intermediate values of reverse(z) are those of forward(x)) and then the reverse-direction
        backward (nf_flow_inverse_backward).  The sampling call itself costs the same either way; under
        ``torch.no_grad()`` (the reference's evaluation loops, torch_nfbc_ant.py:207-210) no node is created.
        ``return_sequence=True`` results are not differentiable."""
        self._check_inputs(x, y)
        params = [p for p in self._tparams if p.requires_grad] if torch.is_grad_enabled() else []
        need_grad = (torch.is_grad_enabled() and not return_sequence and x.shape[0] > 0
                     and (x.requires_grad or (y is not None and y.requires_grad) or len(params) > 0))
        if need_grad:
            carrier = params[:1] if self.fast_param_grads else params
            out, logdet = _FlowReverse.apply(self, self.fast_param_grads, _accumulate_nodes(carrier), return_logdet,
                                             x, y, *carrier)
            return (out, logdet) if return_logdet else out
        out, logdet, seq = self._run_inverse(x, y, return_logdet, return_sequence)
        res = list(seq.unbind(0)) if return_sequence else out
        return (res, logdet) if return_logdet else res

    # ------------------------------------------------------------------ extras
    def log_prob(self, x, y=None):
        """log N(z; 0, I) + log_dets -- what the callers assemble from ``prior.log_prob(z) + logdets``
        (torch_nfbc_ant.py:269) for the standard-normal prior, without the O(D^2) Mahalanobis path."""
        res = self.forward(x, y)
        z, logdet = res[0], res[-1]
        return -0.5 * (z * z).sum(dim=1) - 0.5 * self.in_channels * math.log(2.0 * math.pi) + logdet

    @torch.no_grad()
    def argmin_log_prob(self, x, y=None, return_index=False):
        """For candidates ``x (G, N, D)`` (conditioning ``y (G, N, R)`` or ``None`` = the masked, all-zero conditioning of
        the caller): the candidate with the LOWEST log-probability of every group, ``(G, D)`` -- the exploration-goal
        selection ``rg[argmin(prior.log_prob(z) + logdets)]`` of unsupservised-gcrl/train_auto_NF_rl.py:328-333, vmapped
        over the groups.  One flow forward over the G * N rows (``y = None`` takes the y == 0 fast path) and one reduction
        kernel (``nf_logprob_group_argmin``); ties go to the smaller index like ``jnp.argmin``."""
        if x.dim() != 3 or x.shape[2] != self.in_channels:
            raise ValueError(f"x must be (G, N, {self.in_channels}), got {tuple(x.shape)}")
        G, N, D = x.shape
        xf = _dense(x.reshape(G * N, D))
        yf = _dense(y.reshape(G * N, -1)) if y is not None else None
        res = self.forward(xf, yf)
        z, logdet = _dense(res[0]), _dense(res[-1])
        out = torch.empty((G, D), dtype=torch.float32, device=x.device)
        idx = torch.empty((G,), dtype=torch.int64, device=x.device)
        with _on(x.device):
            _lib.check(lib.nf_logprob_group_argmin(z.data_ptr(), logdet.data_ptr(), xf.data_ptr(), G, N, D, out.data_ptr(),
                                                   idx.data_ptr(), None, self._stream(x.device)), "nf_logprob_group_argmin")
        return (out, idx) if return_index else out

    def augment(self, x, noise_std, seed=0):
        """``x + noise_std * randn_like(x)`` -- the training-time noise augmentation of torch_nfbc_ant.py:264-265 -- as
        one library kernel (nf_noise_augment: counter-based Philox4x32-10 + Box-Muller).  The module keeps a running
        counter, so successive calls with the same ``seed`` draw fresh noise; the stream is reproducible from
        (seed, call order).  Not differentiated through (the reference adds the noise to the data, outside autograd)."""
        if not x.is_cuda or x.dtype != torch.float32:
            raise RuntimeError("nf_b200.RealNVP.augment needs a float32 CUDA tensor")
        xd = _dense(x.detach())
        out = torch.empty_like(xd)
        n = xd.numel()
        off = self._noise_offset.get(int(seed), 0)
        with _on(xd.device):
            _lib.check(lib.nf_noise_augment(xd.data_ptr(), n, float(noise_std), int(seed) & (2**64 - 1), off,
                                            out.data_ptr(), self._stream(xd.device)), "nf_noise_augment")
        self._noise_offset[int(seed)] = off + (n + 3) // 4
        return out

    def nll(self, x, y=None, noise_std=0.0, seed=0):
        """-mean(log N(z; 0, I) + log_dets): the training loss of torch_nfbc_ant.py:269-270 for the standard-normal
        prior, with the loss reduction and its cotangents fused into two library kernels.  Parameter gradients
        go to the flat gradient arena (``fast_param_grads`` semantics).  ``noise_std > 0`` first applies the
        reference's noise augmentation (:264-265) with the library's Philox kernel (``augment``)."""
        if noise_std:
            x = self.augment(x, noise_std, seed)
        self._check_inputs(x, y)
        params = [p for p in self._tparams if p.requires_grad] if torch.is_grad_enabled() else []
        if torch.is_grad_enabled() and (x.requires_grad or (y is not None and y.requires_grad) or params):
            return _FlowNLL.apply(self, _accumulate_nodes(params[:1]), x, y, *params[:1])
        return -self.log_prob(x, y).mean()

    def denoise(self, x, y, noise_std):
        """One denoising step ``x + noise_std**2 * d/dx [log N(z; 0, I) + log_dets]`` -- the evaluation-time step of
        torch_nfbc_ant.py:212-217 (clone, forward, ``autograd.grad(logprob.sum(), [action])``, in-place update) as one
        call without an autograd graph: forward with a stash, the cotangents (dz = -z, dlogdet = 1) from the library's
        NLL kernel, a backward that only produces dx (no parameter or y gradients), one axpy.  Returns a new tensor."""
        self._check_inputs(x, y)
        dev, B, D = x.device, x.shape[0], self.in_channels
        x, y = _dense(x.detach()), (_dense(y.detach()) if y is not None else None)
        if B == 0:
            return x.clone()
        z, logdet, _, stash, packed, slot = self._run_forward(x, y, True, want_outputs=False)
        try:
            with _on(dev):
                # nf_nll_cotangents yields dz = g z / B, dlogdet = -g / B: g = -B gives the gradient of logprob.sum()
                if self._neg_b is None or self._neg_b[0] != (B, dev):
                    self._neg_b = ((B, dev), torch.full((4,), -float(B), dtype=torch.float32, device=dev))
                g = self._neg_b[1]
                dz = torch.empty_like(z)
                dld = torch.empty((B,), dtype=torch.float32, device=dev)
                _lib.check(lib.nf_nll_cotangents(z.data_ptr(), g.data_ptr(), B, D, dz.data_ptr(), dld.data_ptr(),
                                                 self._stream(dev)), "nf_nll_cotangents")
            dx, _, _ = self._run_backward(stash, packed, y, B, dz, dld, True, False, False)
        finally:
            if slot is not None:
                slot.busy = False
        return torch.add(x, dx, alpha=float(noise_std) ** 2)

    def sample_denoised(self, y, noise_std, generator=None):
        """``reverse(prior sample, y)`` followed by one ``denoise`` step (torch_nfbc_ant.py:207-217)."""
        return self.denoise(self.sample(y, generator=generator), y, noise_std)

    def sample(self, y, generator=None):
        """reverse(prior sample, y) with a standard-normal prior (torch_nfbc_ant.py:209-210)."""
        eps = torch.randn((y.shape[0], self.in_channels), dtype=torch.float32, device=y.device, generator=generator)
        return self.reverse(eps, y)
