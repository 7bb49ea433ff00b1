"""cProfile of the host side of one training step (C2a), to see where the Python / launch time goes."""
import sys, os, cProfile, pstats, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import nf_b200
from oracle import synth
cfg = dict(synth.CONFIGS["C2a"]); eps = cfg.pop("ln_eps"); H1 = cfg.pop("H1")
dev = torch.device("cuda:0")
sd = synth.make_state_dict(cfg["D"], cfg["C"], cfg["R"], cfg["n"], seed=0, H1=H1)
m = nf_b200.RealNVP(cfg["D"], cfg["C"], cfg["R"], cfg["n"], None, first_hidden=H1, ln_eps=eps, precision="fp32")
m.load_state_dict({k: torch.tensor(v) for k, v in sd.items()}); m = m.to(dev)
x, y = synth.make_inputs(4096, cfg["D"], cfg["R"])
x = torch.tensor(x, device=dev); y = torch.tensor(y, device=dev)
plist = [p for p in m.parameters() if p.requires_grad]
D = cfg["D"]
def step():
    m._packed_key = None
    for p in plist: p.grad = None
    yy = y.detach().requires_grad_(True)
    z, _, ld = m(x, yy)
    loss = -((-0.5 * (z * z).sum(1) - 0.5 * D * 1.8378770664) + ld).mean()
    loss.backward()
for _ in range(10): step()
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
for _ in range(200): step()
pr.disable(); torch.cuda.synchronize()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats(sys.argv[1] if len(sys.argv) > 1 else "cumulative").print_stats(30); print(s.getvalue()[:6000])
