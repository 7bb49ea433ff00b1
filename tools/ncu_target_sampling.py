"""One or two sampling passes (model.reverse) of a config: the target of ncu captures of the inverse direction.
python tools/ncu_target_sampling.py [C4] [8192] [passes]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import nf_b200
from oracle import synth
name = sys.argv[1] if len(sys.argv) > 1 else "C4"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
passes = int(sys.argv[3]) if len(sys.argv) > 3 else 2
cfg = dict(synth.CONFIGS[name]); eps = cfg.pop("ln_eps"); H1 = cfg.pop("H1")
dev = torch.device("cuda:0")
sd = synth.make_state_dict(cfg["D"], cfg["C"], cfg["R"], cfg["n"], seed=0, H1=H1)
m = nf_b200.RealNVP(cfg["D"], cfg["C"], cfg["R"], cfg["n"], None, first_hidden=H1, ln_eps=eps, precision="fp32")
m.load_state_dict({k: torch.tensor(v) for k, v in sd.items()}); m = m.to(dev)
x, y = synth.make_inputs(B, cfg["D"], cfg["R"])
x = torch.tensor(x, device=dev); y = torch.tensor(y, device=dev)
with torch.no_grad():
    for _ in range(passes):
        out = m.reverse(x, y)
torch.cuda.synchronize()
print("ok", float(out.abs().mean()))
