"""C4 denoise-gradient discrepancy at full size: worst rows and which precision mode shows it."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import tests.test_fullsize_gpu as T
from nf_b200 import _lib
cfg, sd, _, y = T.setup("C4", 8192)
B = 8192
c64, c32 = T.Checker(cfg, sd, torch.float64), T.Checker(cfg, sd, torch.float32)
gen = torch.Generator(device=T.dev()).manual_seed(7)
eps = torch.randn((B, cfg["D"]), device=T.dev(), generator=gen)
a64 = c64.reverse(eps, y)
act = a64.float().contiguous()
z64, ld64, minabs = c64.forward(act, y, want_minabs=True)
dz = -z64; dld = torch.ones(B, device=T.dev(), dtype=torch.float64)
dx64, _, _ = c64.backward(act, y, dz, dld)
dx32, _, _ = c32.backward(act, y, dz, dld)
scale = float(dx64.abs().max())
print("act absmax", float(act.abs().max()), "z64 absmax", float(z64.abs().max()), "dx64 absmax", scale, "ld64 range", float(ld64.min()), float(ld64.max()))
def rep(tag, g, z=None):
    err = (g.double() - dx64).abs().amax(1) / scale
    o = torch.argsort(err, descending=True)[:6]
    print(tag, [(f"{float(err[i]):.1e}", f"minabs {float(minabs[i]):.1e}", f"|act| {float(act[i].abs().max()):.1f}", f"|dx| {float(dx64[i].abs().max()):.1f}") for i in o])
    for tmargin in (3e-5, 1e-4, 3e-4):
        s = minabs > tmargin
        print("    rows minabs>%.0e: %d, max err %.2e" % (tmargin, int(s.sum()), float(err[s].max()) if int(s.sum()) else -1))
    if z is not None: print("    z err", T.rel(z, z64))
rep("ref32", dx32)
for prec in ("fp32", "fp32_simt"):
    model = T.ours_of(cfg, sd, precision=prec)
    a = act.clone().requires_grad_(True)
    z, _, ld = model(a, y)
    (g,) = torch.autograd.grad((-0.5 * (z * z).sum(1) + ld).sum(), [a])
    rep(prec, g, z.detach())
    # same cotangents as the checker (not ours' own z)
    a = act.clone().requires_grad_(True)
    z, _, ld = model(a, y)
    (g2,) = torch.autograd.grad((z * dz.float()).sum() + (ld * dld.float()).sum(), [a])
    rep(prec + " (checker cotangents)", g2)
