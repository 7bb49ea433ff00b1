"""Quick device timing of forward / forward+backward / reverse for the BASELINE configs."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import nf_b200
from nf_b200 import _lib
from oracle import synth

def run(name, B, precision, iters=10):
    cfg = dict(synth.CONFIGS[name]); eps = cfg.pop("ln_eps"); H1 = cfg.pop("H1")
    dev = torch.device("cuda:0")
    sd = synth.make_state_dict(cfg["D"], cfg["C"], cfg["R"], cfg["n"], seed=0, H1=H1)
    m = nf_b200.RealNVP(cfg["D"], cfg["C"], cfg["R"], cfg["n"], None, first_hidden=H1, ln_eps=eps, precision=precision)
    m.load_state_dict({k: torch.tensor(v) for k, v in sd.items()}); m = m.to(dev)
    x, y = synth.make_inputs(B, cfg["D"], cfg["R"])
    x = torch.tensor(x, device=dev); y = torch.tensor(y, device=dev)
    plist = [p for p in m.parameters() if p.requires_grad]
    def train():
        for p in plist: p.grad = None
        loss = -m.log_prob(x, y).mean(); loss.backward()
    def fwd():
        with torch.no_grad(): m(x, y)
    def rev():
        with torch.no_grad(): m.reverse(x, y)
    out = {}
    for nm, fn in (("train", train), ("fwd", fwd), ("rev", rev)):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        _lib.lib.nf_launch_count(1)
        t0 = time.perf_counter()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters): fn()
        e1.record(); torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) / iters * 1e3
        ms = e0.elapsed_time(e1) / iters
        out[nm] = (ms, wall, _lib.lib.nf_launch_count(1) // iters)
    F = synth.flops_fwd_per_sample(**synth.CONFIGS[name])
    print(f"{name} B={B} {precision}: " + " | ".join(
        f"{k} {v[0]:.3f} ms (wall {v[1]:.3f}, {v[2]} launches) {B / v[0] * 1e3:.3e}/s {F * (3 if k == 'train' else 1) * B / v[0] / 1e9:.1f} TF/s"
        for k, v in out.items()), flush=True)

if __name__ == "__main__":
    precs = sys.argv[1].split(",") if len(sys.argv) > 1 else ["fp32_simt"]
    specs = sys.argv[2:] or ["C2a:4096", "C3:16384", "C4:8192", "C1a:256", "C5:8192"]
    for prec in precs:
        for spec in specs:
            name, B = spec.split(":")
            run(name, int(B), prec, iters=3 if name == "C5" else 10)
