"""In-kernel timeline of the tcgen05 GEMM variants (clock64 per phase, see nf_debug_tc_timeline)."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nf_b200 import _lib
lib = _lib.lib
lib.nf_debug_tc_timeline.argtypes = [ctypes.c_void_p]
dev = torch.device("cuda:0")
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
lib.nf_graphs_enable(0)
NAMES = ["prologue", "first tile", "mainloop(split)", "mma drain", "tmem->regs (1st chunk if plain)", "exchange (1st chunk stores if plain)", "finish+store (rest if plain)"]

def run(M, N, K, engine=4, reps=3):
    A = torch.randn(M, K, device=dev); W = torch.randn(N, K, device=dev) / K ** 0.5; C = torch.empty(M, N, device=dev)
    ctas = ((M + 127) // 128) * ((N + 127) // 128) * 4
    buf = torch.zeros(ctas * 16, dtype=torch.int64, device=dev)
    for _ in range(reps):
        lib.nf_linear(A.data_ptr(), K, W.data_ptr(), K, None, C.data_ptr(), N, M, N, K, engine, st)
    torch.cuda.synchronize()
    lib.nf_debug_tc_timeline(buf.data_ptr())
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    lib.nf_linear(A.data_ptr(), K, W.data_ptr(), K, None, C.data_ptr(), N, M, N, K, engine, st)
    e1.record(); torch.cuda.synchronize()
    lib.nf_debug_tc_timeline(None)
    t = buf.view(-1, 16).cpu()
    t = t[t[:, 0] != 0]
    d = (t[:, 1:8] - t[:, :7]).double()
    if (t[:, 8] != 0).any():
        names = ['mbar_wait(split)', 'syncwarp', 'fence + elect', '6 MMA issue', '2 commits + syncwarp', 'rest of loop (to next top)']
        for i, nm in enumerate(names):
            v = (t[:, 9 + i] - t[:, 8 + i]).double()
            print(f'    MMA warp stage 8: {nm:36s} {v.mean():8.0f}  (min {v.min():6.0f} max {v.max():6.0f})')
    # plain epilogue: slots 5, 6 unused -> fold
    print(f"engine={engine} M={M} N={N} K={K}: {e0.elapsed_time(e1)*1e3:.1f} us event, {len(t)} CTAs; mean cycles per phase:")
    tot = (t[:, 7] - t[:, 0]).double()
    for i, n in enumerate(NAMES):
        v = d[:, i]
        if (t[:, i + 1] == 0).all() or (t[:, i] == 0).all(): continue
        print(f"    {n:40s} {v.mean():9.0f}  (min {v.min():7.0f} max {v.max():7.0f})")
    print(f"    {'epilogue total':18s} {(t[:, 7] - t[:, 4]).double().mean():9.0f}")
    print(f"    {'total in CTA':18s} {tot.mean():9.0f} cycles = {tot.mean()/1.965e3:.2f} us at 1965 MHz")

if __name__ == "__main__":
    if len(sys.argv) > 1:
        for spec in sys.argv[1:]:
            parts = [int(v) for v in spec.split("x")]
            M, N, K = parts[:3]
            run(M, N, K, engine=parts[3] if len(parts) > 3 else 4)
    else:
        run(8192, 1024, 1024)
