import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nf_b200 import _lib
lib = _lib.lib
dev = torch.device("cuda:0")
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

def run(B, n_out, n_in, engine, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    dY = torch.randn(B, n_out, generator=g).to(dev)
    X = torch.randn(B, n_in, generator=g).to(dev)
    base = torch.randn(n_out, n_in, generator=g).to(dev)
    dW = base.clone()
    rc = lib.nf_linear_wgrad(dY.data_ptr(), n_out, X.data_ptr(), n_in, dW.data_ptr(), n_in, B, n_out, n_in, engine, st)
    if rc != 0:
        print(f"B={B} {n_out}x{n_in} engine={engine}: rc={rc} {lib.nf_last_error().decode()}"); return
    torch.cuda.synchronize()
    ref = base.double() + dY.double().T @ X.double()
    err = ((dW.double() - ref).abs().max() / ref.abs().max()).item()
    print(f"B={B} n_out={n_out} n_in={n_in} engine={engine}: rel err {err:.3e}", flush=True)

if __name__ == "__main__":
    for eng in (0, 1, 3, 5):
        run(32, 128, 256, eng)
    for eng in (1, 5):
        run(256, 128, 256, eng); run(4096, 512, 512, eng); run(1000, 200, 260, eng); run(8192, 1024, 1024, eng)
        run(300, 64, 36, eng); run(5000, 512, 1024, eng)
    B, n_out, n_in = 16384, 512, 512
    dY = torch.randn(B, n_out, device=dev); X = torch.randn(B, n_in, device=dev); dW = torch.zeros(n_out, n_in, device=dev)
    for eng in (0, 1, 3, 5):
        for _ in range(3): lib.nf_linear_wgrad(dY.data_ptr(), n_out, X.data_ptr(), n_in, dW.data_ptr(), n_in, B, n_out, n_in, eng, st)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): lib.nf_linear_wgrad(dY.data_ptr(), n_out, X.data_ptr(), n_in, dW.data_ptr(), n_in, B, n_out, n_in, eng, st)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        print(f"wgrad engine {eng}: {ms*1e3:.1f} us  {2*B*n_out*n_in/ms/1e9:.1f} TFLOP/s (algorithmic)")
