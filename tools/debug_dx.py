"""Where does the input-gradient error at full size come from?  Per-row dx error of a config against the float64
checker, with the distance of the row's nearest LeakyReLU pre-activation from the kink."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import tests.test_fullsize_gpu as T
from nf_b200 import _lib

name, B = sys.argv[1], int(sys.argv[2])
cfg, sd, x, y = T.setup(name, B)
c64 = T.Checker(cfg, sd, torch.float64)
c32 = T.Checker(cfg, sd, torch.float32)
z64, ld64, minabs = c64.forward(x, y, want_minabs=True)
dz = -z64; dld = torch.ones(B, device=x.device, dtype=torch.float64)
dx64, _, _ = c64.backward(x, y, dz, dld)
dx32, _, _ = c32.backward(x, y, dz, dld)
scale = float(dx64.abs().max())
def report(tag, dx):
    err = (dx.double() - dx64).abs().amax(dim=1) / scale
    order = torch.argsort(err, descending=True)[:8]
    print(tag, "max row err %.3e" % float(err.max()), "rows>1e-5:", int((err > 1e-5).sum()),
          "| worst rows (err, minabs):", [(f"{float(err[i]):.1e}", f"{float(minabs[i]):.1e}") for i in order])
    for t in (1e-6, 3e-6, 1e-5, 2e-5, 3e-5, 1e-4):
        safe = minabs > t
        if int(safe.sum()):
            print("    max err over rows with minabs > %.0e: %.3e (%d rows)" % (t, float(err[safe].max()), int(safe.sum())))
report("ref32", dx32)
for prec in ("fp32", "fp32_simt"):
    for fl in (-1,):
        _lib.set_option("fused_ln", fl)
        model = T.ours_of(cfg, sd, precision=prec)
        a = x.clone().requires_grad_(True)
        z, _, ld = model(a, y)
        (g,) = torch.autograd.grad((-0.5 * (z * z).sum(1) + ld).sum(), [a])
        report(f"{prec} fused_ln={fl}", g)
        print("    z err %.2e" % T.rel(z.detach(), z64))
