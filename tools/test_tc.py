"""tcgen05 GEMM (nf_linear) against a float64 matmul, several shapes/engines."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nf_b200 import _lib
lib = _lib.lib
dev = torch.device("cuda:0")
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

def run(M, N, K, engine, pad=0, bias=True, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    A = torch.randn(M, K + pad, generator=g).to(dev)
    W = (torch.randn(N, K + pad, generator=g) / K ** 0.5).to(dev)
    b = torch.randn(N, generator=g).to(dev) if bias else None
    C = torch.full((M, N), float("nan"), device=dev)
    rc = lib.nf_linear(A.data_ptr(), K + pad, W.data_ptr(), K + pad, b.data_ptr() if bias else None, C.data_ptr(), N,
                       M, N, K, engine, st)
    if rc != 0:
        print(f"M={M} N={N} K={K} engine={engine}: rc={rc} {lib.nf_last_error().decode()}"); return
    torch.cuda.synchronize()
    ref = A[:, :K].double() @ W[:, :K].double().T + (b.double() if bias else 0)
    err = ((C.double() - ref).abs().max() / ref.abs().max()).item()
    print(f"M={M} N={N} K={K} pad={pad} engine={engine}: rel err {err:.3e} nan={int(torch.isnan(C).sum())}", flush=True)
    return C

if __name__ == "__main__":
    for eng in (0, 1, 3, 4, 5, 6):
        run(128, 256, 32, eng)
    for eng in (1, 3, 4, 5):
        run(128, 256, 256, eng)
        run(4096, 512, 260, eng)
    for eng in (1, 5):
        run(257, 200, 36, eng)
        run(1000, 64, 512, eng, pad=4)
        run(300, 136, 100, eng)
        run(128, 16, 8, eng)
        run(8192, 1024, 1024, eng)
    # does the tensor core truncate raw fp32 operands?  engine 4 (raw as hi, lo from truncation) vs 1
    c1 = run(512, 256, 512, 1, seed=3); c4 = run(512, 256, 512, 4, seed=3)
    print("raw-hi vs rna split max diff", (c1 - c4).abs().max().item())
    # timing
    M, N, K = 16384, 512, 512
    A = torch.randn(M, K, device=dev); W = torch.randn(N, K, device=dev); C = torch.empty(M, N, device=dev)
    for eng in (0, 1, 3, 4, 5, 6):
        for _ in range(3): lib.nf_linear(A.data_ptr(), K, W.data_ptr(), K, None, C.data_ptr(), N, M, N, K, eng, st)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): lib.nf_linear(A.data_ptr(), K, W.data_ptr(), K, None, C.data_ptr(), N, M, N, K, eng, st)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        print(f"engine {eng}: {ms*1e3:.1f} us  {2*M*N*K/ms/1e9:.1f} TFLOP/s (algorithmic)")
    torch.backends.cuda.matmul.allow_tf32 = False
    for _ in range(3): torch.matmul(A, W.T)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): torch.matmul(A, W.T)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"cuBLAS fp32: {ms*1e3:.1f} us  {2*M*N*K/ms/1e9:.1f} TFLOP/s")
