"""The "stock PyTorch on the same B200" bar: oracle/torch_flow.py (eager nn.Linear /
LayerNorm / cat / exp + autograd, the op mix of the reference module) timed on cuda:0 next to nf_b200 on the
same configs, weights and inputs.  TF32 matmuls are left at torch's default (off): both sides are fp32-parity.

  python tools/eager_bar.py [C2a,C3,C4,C1a,C5]
"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import nf_b200
from oracle import synth
from oracle.torch_flow import TorchEncoder, TorchFlow

BATCH = {"C2a": 4096, "C3": 16384, "C4": 8192, "C1a": 256, "C1b": 256, "C1c": 256, "C5": 8192}


def timed(fn, iters, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def run(name, iters):
    cfg = dict(synth.CONFIGS[name]); eps = cfg.pop("ln_eps"); H1 = cfg.pop("H1")
    B, D = BATCH[name], cfg["D"]
    dev = torch.device("cuda:0")
    sd = {k: torch.tensor(v) for k, v in synth.make_state_dict(D, cfg["C"], cfg["R"], cfg["n"], seed=0, H1=H1).items()}
    ours = nf_b200.RealNVP(D, cfg["C"], cfg["R"], cfg["n"], None, first_hidden=H1, ln_eps=eps, precision="fp32")
    ours.load_state_dict(sd); ours = ours.to(dev)
    eager = TorchFlow(D, cfg["C"], cfg["R"], cfg["n"], first_hidden=H1, ln_eps=eps)
    eager.load_state_dict(sd); eager = eager.to(dev)
    x, y = synth.make_inputs(B, D, cfg["R"])
    x = torch.tensor(x, device=dev); y = torch.tensor(y, device=dev)

    def legs(m):
        plist = [p for p in m.parameters() if p.requires_grad]

        def train():
            for p in plist:
                p.grad = None
            yy = y.detach().requires_grad_(True)
            z, _, ld = m(x, yy)
            (-((-0.5 * (z * z).sum(1) - 0.5 * D * 1.8378770664093453) + ld).mean()).backward()

        def sample():
            with torch.no_grad():
                m.reverse(x, y)
        return train, sample

    row = {}
    for tag, m in (("eager", eager), ("nf_b200", ours)):
        train, sample = legs(m)
        row[tag] = (timed(train, iters), timed(sample, iters))
    (et, es), (ot, os_) = row["eager"], row["nf_b200"]
    print(f"{name} B={B}: train  eager {et:.3f} ms  nf_b200 {ot:.3f} ms  x{et / ot:.2f} | "
          f"sample eager {es:.3f} ms  nf_b200 {os_:.3f} ms  x{es / os_:.2f}", flush=True)


def run_encoder(name, B, iters=30):
    """encoder forward + backward (parameters and input gradients off: observations need none), eager vs nf_b200"""
    cfg = synth.ENCODER_CONFIGS[name]
    dev = torch.device("cuda:0")
    sd = {k: torch.tensor(v) for k, v in synth.make_encoder_state_dict(cfg["input_size"], cfg["rep_size"]).items()}
    ours = nf_b200.GEncoder(cfg["input_size"], cfg["rep_size"]); ours.load_state_dict(sd); ours = ours.to(dev)
    eager = TorchEncoder(cfg["input_size"], cfg["rep_size"]); eager.load_state_dict(sd); eager = eager.to(dev)
    g, dout = synth.make_encoder_inputs(B, cfg["input_size"], cfg["rep_size"])
    g = torch.tensor(g, device=dev); dout = torch.tensor(dout, device=dev)
    row = {}
    for tag, m in (("eager", eager), ("nf_b200", ours)):
        plist = list(m.parameters())

        def train():
            for p in plist:
                p.grad = None
            m(g).backward(dout)

        def infer():
            with torch.no_grad():
                m(g)
        row[tag] = (timed(train, iters), timed(infer, iters))
    (et, ei), (ot, oi) = row["eager"], row["nf_b200"]
    print(f"encoder {name} B={B}: fwd+bwd eager {et:.3f} ms  nf_b200 {ot:.3f} ms  x{et / ot:.2f} | "
          f"fwd eager {ei:.3f} ms  nf_b200 {oi:.3f} ms  x{ei / oi:.2f}", flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "encoder":
        for nm, b in (("gcbc", 4096), ("kitchen", 8192), ("ant", 256)):
            if len(sys.argv) < 3 or sys.argv[2] == nm:
                run_encoder(nm, b)
        sys.exit(0)
    names = sys.argv[1].split(",") if len(sys.argv) > 1 else ["C2a", "C3", "C4", "C1a", "C5"]
    for nm in names:
        run(nm, 5 if nm == "C5" else 30)
