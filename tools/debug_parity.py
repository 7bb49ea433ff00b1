import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import nf_b200
from oracle import synth, flow_oracle as fo

def err(a, b):
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-30))

def run(D, C, R, n, B, precision="fp32_simt", H1=None, verbose=False):
    dev = torch.device("cuda:0")
    sd = synth.make_state_dict(D, C, R, n, seed=0, H1=H1)
    x, y = synth.make_inputs(B, D, R); dz, dld = synth.make_cotangents(B, D)
    m = nf_b200.RealNVP(D, C, R, n, None, first_hidden=H1, precision=precision)
    m.load_state_dict({k: torch.tensor(v) for k, v in sd.items()}); m = m.to(dev)
    xt = torch.tensor(x, device=dev, requires_grad=True); yt = torch.tensor(y, device=dev, requires_grad=True)
    z, outs, ld = m(xt, yt)
    ((z * torch.tensor(dz, device=dev)).sum() + (ld * torch.tensor(dld, device=dev)).sum()).backward()
    sd64 = fo.cast_state_dict(sd, np.float64)
    ref = fo.flow_backward(sd64, n, *(a.astype(np.float64) for a in (x, y, dz, dld)))
    sd32 = fo.cast_state_dict(sd, np.float32)
    r32 = fo.flow_backward(sd32, n, x, y, dz, dld)
    e = dict(z=err(z.detach().cpu().numpy(), ref["z"]), ld=err(ld.detach().cpu().numpy(), ref["log_dets"]),
             dx=err(xt.grad.cpu().numpy(), ref["dx"]), dy=err(yt.grad.cpu().numpy(), ref["dy"]))
    e32 = dict(dx=err(r32["dx"], ref["dx"]), dy=err(r32["dy"], ref["dy"]))
    ge = {k: err(p.grad.cpu().numpy(), ref["grads"][k]) for k, p in m.named_parameters() if p.requires_grad}
    worst = sorted(ge.items(), key=lambda kv: -kv[1])[:4]
    print(f"D={D} C={C} R={R} n={n} B={B} {precision}:", {k: f"{v:.1e}" for k, v in e.items()},
          "oracle32:", {k: f"{v:.1e}" for k, v in e32.items()}, "worst grads:", [(k, f"{v:.1e}") for k, v in worst], flush=True)

if __name__ == "__main__":
    prec = sys.argv[1] if len(sys.argv) > 1 else "fp32_simt"
    run(8, 512, 256, 12, 64, prec)
    run(8, 512, 256, 1, 64, prec)
    run(8, 512, 256, 3, 64, prec)
    run(8, 128, 256, 12, 64, prec)
    run(8, 512, 16, 12, 64, prec)
    run(9, 512, 256, 12, 64, prec)
    run(8, 512, 256, 12, 300, prec)
