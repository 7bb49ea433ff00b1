"""Where does the host time of one synchronous training step go?  (bench.py's e2e leg)"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import nf_b200
from oracle import synth

cfg = dict(synth.CONFIGS["C2a"]); eps = cfg.pop("ln_eps"); H1 = cfg.pop("H1")
B = 4096
dev = torch.device("cuda:0")
sd = synth.make_state_dict(cfg["D"], cfg["C"], cfg["R"], cfg["n"], seed=0, H1=H1)
m = nf_b200.RealNVP(cfg["D"], cfg["C"], cfg["R"], cfg["n"], None, first_hidden=H1, ln_eps=eps, precision="fp32")
m.load_state_dict({k: torch.tensor(v) for k, v in sd.items()}); m = m.to(dev)
xh, yh = synth.make_inputs(B, cfg["D"], cfg["R"])
xp, yp = torch.tensor(xh).pin_memory(), torch.tensor(yh).pin_memory()
plist = [p for p in m.parameters() if p.requires_grad]
D = cfg["D"]

def step(sync_points):
    T = {}
    t0 = time.perf_counter()
    x = xp.to(dev, non_blocking=True); y = yp.to(dev, non_blocking=True)
    if sync_points: torch.cuda.synchronize()
    T["h2d"] = time.perf_counter() - t0; t0 = time.perf_counter()
    m._packed_key = None
    for p in plist: p.grad = None
    z, _, ld = m(x, y)
    if sync_points: torch.cuda.synchronize()
    T["fwd"] = time.perf_counter() - t0; t0 = time.perf_counter()
    loss = -((-0.5 * (z * z).sum(1) - 0.5 * D * 1.8378770664) + ld).mean()
    if sync_points: torch.cuda.synchronize()
    T["loss"] = time.perf_counter() - t0; t0 = time.perf_counter()
    loss.backward()
    if sync_points: torch.cuda.synchronize()
    T["bwd"] = time.perf_counter() - t0; t0 = time.perf_counter()
    v = loss.item()
    T["item"] = time.perf_counter() - t0
    return T

for sp in (True, False):
    for _ in range(3): step(sp)
    acc = {}
    for _ in range(10):
        for k, v in step(sp).items(): acc[k] = acc.get(k, 0) + v * 100
    print("sync_points" if sp else "async      ", {k: f"{v:.3f} ms" for k, v in acc.items()}, "total %.3f ms" % sum(acc.values()), flush=True)
