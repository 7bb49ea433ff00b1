"""Error anatomy of the 3xTF32 tcgen05 GEMM: rms and signed mean error vs float64 for several engines
and reduction lengths, random-sign and all-positive operands (the latter exposes accumulate truncation)."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nf_b200 import _lib
lib = _lib.lib
dev = torch.device("cuda:0")
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
torch.backends.cuda.matmul.allow_tf32 = False
M, N = 1024, 256
for positive in (False, True, 'neg', 'mixed'):
    for K in (256, 1024, 2048):
        g = torch.Generator(device="cpu").manual_seed(K)
        A = torch.randn(M, K, generator=g); W = torch.randn(N, K, generator=g) / K ** 0.5
        if positive is True: A, W = A.abs(), W.abs()
        elif positive == 'neg': A, W = A.abs(), -W.abs()
        elif positive == 'mixed': A = A.abs()
        A, W = A.to(dev), W.to(dev)
        ref = A.double() @ W.double().T
        scale = ref.abs().mean().item()
        out = []
        for eng in (0, 1, 4):
            C = torch.empty(M, N, device=dev)
            rc = lib.nf_linear(A.data_ptr(), K, W.data_ptr(), K, None, C.data_ptr(), N, M, N, K, eng, st)
            assert rc == 0, lib.nf_last_error()
            e = C.double() - ref
            out.append(f"eng{eng}: rms {e.pow(2).mean().sqrt().item()/scale:.2e} mean {e.mean().item()/scale:+.2e} max {e.abs().max().item()/ref.abs().max().item():.2e}")
        e = (A @ W.T).double() - ref
        out.append(f"cublas32: rms {e.pow(2).mean().sqrt().item()/scale:.2e} mean {e.mean().item()/scale:+.2e} max {e.abs().max().item()/ref.abs().max().item():.2e}")
        print(f"positive={positive} K={K}\n   " + "\n   ".join(out), flush=True)
