"""Upper bounds for removing work from the training step (EXPERIMENT: option "whatif" skips kernels, gradients are
wrong while it is set): device time of a training step with column reductions / weight-gradient GEMMs / the dy GEMM
skipped.  python tools/whatif.py [C2a:4096 C3:16384 ...]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import nf_b200
from nf_b200 import _lib
from oracle import synth

def model_of(name, B):
    cfg = dict(synth.CONFIGS[name]); eps = cfg.pop("ln_eps"); H1 = cfg.pop("H1")
    dev = torch.device("cuda:0")
    sd = synth.make_state_dict(cfg["D"], cfg["C"], cfg["R"], cfg["n"], seed=0, H1=H1)
    m = nf_b200.RealNVP(cfg["D"], cfg["C"], cfg["R"], cfg["n"], None, first_hidden=H1, ln_eps=eps)
    m.load_state_dict({k: torch.tensor(v) for k, v in sd.items()}); m = m.to(dev)
    x, y = synth.make_inputs(B, cfg["D"], cfg["R"])
    return m, torch.tensor(x, device=dev), torch.tensor(y, device=dev)

def time_train(m, x, y, iters=30):
    plist = [p for p in m.parameters() if p.requires_grad]
    def train():
        m._packed_key = None
        for p in plist: p.grad = None
        loss = -m.log_prob(x, y.detach().requires_grad_(True)).mean(); loss.backward()
    for _ in range(5): train()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): train()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

for spec in (sys.argv[1:] or ["C2a:4096", "C3:16384", "C1a:256"]):
    name, B = spec.split(":"); B = int(B)
    m, x, y = model_of(name, B)
    res = {}
    for w, label in ((0, "full"), (1, "-colreduce"), (2, "-wgrad"), (4, "-dy"), (3, "-colreduce-wgrad"), (7, "-all three")):
        _lib.set_option("whatif", w)
        res[label] = time_train(m, x, y)
    _lib.set_option("whatif", 0)
    print(name, B, " | ".join(f"{k} {v:.3f} ms" for k, v in res.items()), flush=True)
