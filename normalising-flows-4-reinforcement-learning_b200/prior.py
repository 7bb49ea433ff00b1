"""Standard-normal prior with the interface the reference's scripts use on ``model.prior``.

The scripts build ``distributions.MultivariateNormal(torch.zeros(D), torch.eye(D))`` and call ``model.prior.log_prob(z)`` in
the loss line and ``model.prior.sample((n,))`` when sampling (il-and-cil/torch_nfbc_ant.py:146-147, 209, 269).  For an
identity covariance the Cholesky / Mahalanobis machinery of ``MultivariateNormal.log_prob`` (about ten small kernels forward,
as many backward) computes ``-0.5 |z|^2 - D/2 log 2 pi``.  ``StandardNormal(D)`` is that prior with ``log_prob`` as ONE library
kernel forward and one backward (``nf_stdnormal_logprob*``), so a script keeps its loss line

    loss = -(model.prior.log_prob(z) + logdets).mean()

and passes ``prior=nf_b200.StandardNormal(D)`` where it passed the MultivariateNormal.  CUDA float32 tensors go through the
kernel; anything else falls back to the closed form in torch (the prior is not part of the flow's hot path contract)."""
from __future__ import annotations

import ctypes
import math

import torch

from . import _lib
from ._lib import lib
from .module import _dense, _on


class _LogProb(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z):
        zd = _dense(z)
        B, D = zd.shape
        lp = torch.empty((B,), dtype=torch.float32, device=zd.device)
        with _on(zd.device):
            st = ctypes.c_void_p(torch._C._cuda_getCurrentRawStream(zd.device.index))
            _lib.check(lib.nf_stdnormal_logprob(zd.data_ptr(), B, D, lp.data_ptr(), st), "nf_stdnormal_logprob")
        ctx.save_for_backward(zd)
        return lp

    @staticmethod
    def backward(ctx, g):
        (zd,) = ctx.saved_tensors
        B, D = zd.shape
        gd = _dense(g)
        dz = torch.empty_like(zd)
        with _on(zd.device):
            st = ctypes.c_void_p(torch._C._cuda_getCurrentRawStream(zd.device.index))
            _lib.check(lib.nf_stdnormal_logprob_bwd(zd.data_ptr(), gd.data_ptr(), B, D, dz.data_ptr(), st),
                       "nf_stdnormal_logprob_bwd")
        return dz


class StandardNormal:
    """``N(0, I_D)`` with ``log_prob(z (..., D)) -> (...)`` and ``sample(sample_shape) -> sample_shape + (D,)``."""

    def __init__(self, dim: int, device=None):
        self.dim = int(dim)
        self.device = torch.device(device) if device is not None else None
        self.event_shape = torch.Size([self.dim])

    def to(self, device):
        self.device = torch.device(device)
        return self

    def log_prob(self, z: torch.Tensor) -> torch.Tensor:
        if z.shape[-1] != self.dim:
            raise ValueError(f"the last dimension of z must be {self.dim}, got {tuple(z.shape)}")
        if z.is_cuda and z.dtype == torch.float32 and z.numel() > 0:
            return _LogProb.apply(z.reshape(-1, self.dim)).reshape(z.shape[:-1])
        return -0.5 * (z * z).sum(-1) - 0.5 * self.dim * math.log(2.0 * math.pi)

    def sample(self, sample_shape=torch.Size()) -> torch.Tensor:
        return torch.randn(tuple(sample_shape) + (self.dim,), device=self.device)

    rsample = sample
