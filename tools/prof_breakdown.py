import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import nf_b200
from nf_b200 import _lib
from oracle import synth

def run(name, B, precision="fp32", mode="train"):
    cfg = dict(synth.CONFIGS[name]); eps = cfg.pop("ln_eps"); H1 = cfg.pop("H1")
    dev = torch.device("cuda:0")
    sd = synth.make_state_dict(cfg["D"], cfg["C"], cfg["R"], cfg["n"], seed=0, H1=H1)
    m = nf_b200.RealNVP(cfg["D"], cfg["C"], cfg["R"], cfg["n"], None, first_hidden=H1, ln_eps=eps, precision=precision)
    m.load_state_dict({k: torch.tensor(v) for k, v in sd.items()}); m = m.to(dev)
    x, y = synth.make_inputs(B, cfg["D"], cfg["R"])
    x = torch.tensor(x, device=dev); y = torch.tensor(y, device=dev)
    def train():
        m._packed_key = None
        for p in m.parameters(): p.grad = None
        loss = -m.log_prob(x, y).mean(); loss.backward()
    def rev():
        with torch.no_grad(): m.reverse(x, y)
    fn = train if mode == "train" else rev
    for _ in range(3): fn()
    torch.cuda.synchronize()
    _lib.lib.nf_prof_enable(1); _lib.prof_collect()
    for _ in range(5): fn()
    torch.cuda.synchronize()
    prof = _lib.prof_collect(); _lib.lib.nf_prof_enable(0)
    tot = sum(v[0] for v in prof.values())
    print(f"--- {name} B={B} {precision} {mode}: total profiled {tot/5:.3f} ms/step")
    for k, (ms, work, n) in prof.items():
        if n: print(f"   {k:10s} {ms/5*1e3:8.1f} us/step  {n//5:4d} launches  avg {ms/n*1e3:7.1f} us   {work/ms/1e9 if ms else 0:8.1f} G(work)/s")

if __name__ == "__main__":
    for a in sys.argv[1:]:
        name, B, mode = a.split(":")
        run(name, int(B), mode=mode)
