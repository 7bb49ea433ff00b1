import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import nf_b200
from nf_b200 import _lib
from oracle import synth, flow_oracle as fo

def err(a, b):
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-30))

D, C, R, n, B = 8, 512, 256, 1, 64
dev = torch.device("cuda:0")
sd = synth.make_state_dict(D, C, R, n, seed=0)
x, y = synth.make_inputs(B, D, R); dz, dld = synth.make_cotangents(B, D)
m = nf_b200.RealNVP(D, C, R, n, None)
m.load_state_dict({k: torch.tensor(v) for k, v in sd.items()}); m = m.to(dev)
xt = torch.tensor(x, device=dev); yt = torch.tensor(y, device=dev)
z, ld, outs, stash, packed = m._run_forward(xt, yt, True)
torch.cuda.synchronize()
st = stash.cpu().numpy()
dc, dt, H1 = 4, 4, C
mw = C // 32
# stash layout
off = 0
xs = st[off:off + (n + 1) * B * D].reshape(n + 1, B, D); off += (n + 1) * B * D
ss = st[off:off + n * B * dt].reshape(n, B, dt); off += n * B * dt
nets = []
for net in range(2):
    d = {}
    d["xh1"] = st[off:off + B * H1].reshape(B, H1); off += B * H1
    d["xh2"] = st[off:off + B * C].reshape(B, C); off += B * C
    d["r1"] = st[off:off + B]; off += B
    d["r2"] = st[off:off + B]; off += B
    d["m1"] = st[off:off + B * mw].view(np.uint32).reshape(B, mw); off += B * mw
    d["m2"] = st[off:off + B * mw].view(np.uint32).reshape(B, mw); off += B * mw
    nets.append(d)
sd64 = fo.cast_state_dict(sd, np.float64)
zz, ldd, cache = fo.block_forward(sd64, 0, x.astype(np.float64), y.astype(np.float64), keep=True)
for net, key in ((0, "ct"), (1, "cs")):
    c = cache[key]
    print("net", net, "xh1", err(nets[net]["xh1"], c["xh1"]), "xh2", err(nets[net]["xh2"], c["xh2"]),
          "r1", err(nets[net]["r1"], c["r1"][:, 0]), "r2", err(nets[net]["r2"], c["r2"][:, 0]))
    for nm, u in (("m1", c["u1"]), ("m2", c["u2"])):
        bits = ((nets[net][nm][:, :, None] >> np.arange(32)[None, None, :]) & 1).reshape(B, -1).astype(bool)
        print("   ", nm, "mismatch", int((bits != (u > 0)).sum()))
# backward direct
dzt = torch.tensor(dz, device=dev); dldt = torch.tensor(dld, device=dev)
dx, dy, dflat = m._run_backward(stash, packed, yt, B, dzt, dldt, True, True, True)
torch.cuda.synchronize()
ws = m._ws.view(torch.float32).cpu().numpy()
def pad32(v): return (v + 31) // 32 * 32
o = 0
d0 = ws[o:o + B * D]; o += pad32(B * D); d1 = ws[o:o + B * D]; o += pad32(B * D)
dxp = ws[o:o + B * D].reshape(B, D); o += pad32(B * D)
dout = ws[o:o + 2 * B * dt].reshape(2, B, dt); o += pad32(2 * B * dt)
gA = ws[o:o + 2 * B * C].reshape(2, B, C); o += pad32(2 * B * C)
gB = ws[o:o + 2 * B * C].reshape(2, B, C); o += pad32(2 * B * C)
# oracle intermediates
e = np.exp(-cache["s"]); dzt_ = dz[:, dc:].astype(np.float64)
dt_ = -dzt_ * e; ds_ = -dzt_ * cache["zt"] - dld[:, None]
print("dout t", err(dout[0], dt_), "dout s", err(dout[1], ds_))
for net, key, dout_ref in ((0, "ct", dt_), (1, "cs", ds_)):
    p = f"blocks.0.{'t' if net == 0 else 's'}."
    c = cache[key]
    dh2 = dout_ref @ sd64[p + "6.weight"]
    du2 = fo._ln_lrelu_backward(dh2, sd64[p + "5.weight"], c["xh2"], c["r2"], c["u2"])
    dh1 = du2 @ sd64[p + "3.weight"]
    du1 = fo._ln_lrelu_backward(dh1, sd64[p + "2.weight"], c["xh1"], c["r1"], c["u1"])
    print("net", net, "du2", err(gA[net], du2), "du1", err(gB[net], du1))
    bad = np.argwhere(np.abs(gA[net] - du2) > 1e-4 * np.abs(du2).max())
    print("   bad du2 entries", len(bad), bad[:10].tolist())
