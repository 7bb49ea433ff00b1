"""Host time spent inside each C-ABI call of one asynchronous C2a training step (perf_counter around the ctypes call,
no synchronisation): how much of the step's host time is the library (graph launches) and how much is Python / autograd."""
import sys, os, time, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import nf_b200
from nf_b200 import _lib
from oracle import synth

acc = collections.defaultdict(float); cnt = collections.Counter()
def wrap(name):
    fn = getattr(_lib.lib, name)
    def w(*a):
        t = time.perf_counter(); r = fn(*a); acc[name] += time.perf_counter() - t; cnt[name] += 1; return r
    setattr(_lib.lib, name, w)
for n in ("nf_flow_prepare", "nf_flow_forward", "nf_flow_backward", "nf_flow_inverse"):
    wrap(n)

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
    return loss
for _ in range(10): step()
torch.cuda.synchronize(); acc.clear(); cnt.clear()
N = 50
tot = 0.0
for _ in range(N):
    t = time.perf_counter(); l = step(); tot += time.perf_counter() - t
    l.item()                      # drain, like the e2e leg
print(f"host time per step (issue only): {tot / N * 1e3:.3f} ms")
for k in acc: print(f"  {k:18s} {acc[k] / N * 1e6:7.1f} us per step ({cnt[k] // N} calls)")
print(f"  python / autograd / torch ops: {(tot - sum(acc.values())) / N * 1e6:7.1f} us per step")
