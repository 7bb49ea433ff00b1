"""Achieved HBM bandwidth of the standalone affine / log-det kernels of the C ABI (SURVEY section 8d:
4*(2D + 2dt + 2) algorithmic bytes per row forward / inverse; backward reads dz (D), z_t, s (dt each), dlogdet and writes dxp
(D), ds, dt: 4*(2D + 4dt + 1)) and of the row tail at a streaming batch size."""
import sys, os, ctypes, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nf_b200 import _lib
lib = _lib.lib
dev = torch.device("cuda:0")
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
peak = 6549.8
try:
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
for D, B in ((8, 1 << 23), (32, 1 << 22), (400, 1 << 18)):
    dt = D // 2
    xp = torch.randn(B, D, device=dev); s = 0.1 * torch.randn(B, dt, device=dev); t = torch.randn(B, dt, device=dev)
    z = torch.empty(B, D, device=dev); ld = torch.zeros(B, device=dev); so = torch.empty(B, dt, device=dev)
    dz = torch.randn(B, D, device=dev); dld = torch.randn(B, device=dev)
    dxp = torch.empty(B, D, device=dev); ds = torch.empty(B, dt, device=dev); dtt = torch.empty(B, dt, device=dev)
    def fwd(): lib.nf_affine_logdet_fwd(xp.data_ptr(), s.data_ptr(), t.data_ptr(), None, None, 0.0, B, D, z.data_ptr(), ld.data_ptr(), None, st)
    def inv(): lib.nf_affine_logdet_inv(z.data_ptr(), s.data_ptr(), t.data_ptr(), None, None, 0.0, B, D, xp.data_ptr(), ld.data_ptr(), st)
    def bwd(): lib.nf_affine_logdet_bwd(dz.data_ptr(), z.data_ptr(), s.data_ptr(), dld.data_ptr(), B, D, dxp.data_ptr(), ds.data_ptr(), dtt.data_ptr(), st)
    for name, fn, nbytes in (("fwd", fwd, 4 * (2 * D + 2 * dt + 2) * B), ("inv", inv, 4 * (2 * D + 2 * dt + 2) * B),
                             ("bwd", bwd, 4 * (2 * D + 4 * dt + 1) * B)):
        for _ in range(3): fn()
        ms = []
        for _ in range(5):
            flush.zero_()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        m = sorted(ms)[len(ms) // 2]
        print(f"affine {name} D={D} B={B}: {m*1e3:.1f} us  {nbytes/m/1e6:.0f} GB/s = {nbytes/m/1e6/peak*100:.0f}% of measured HBM peak ({peak:.0f} GB/s)", flush=True)
