"""Two training steps of the conditioning encoder (forward + backward): the target of the encoder's ncu launch list."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import nf_b200
from oracle import synth
name = sys.argv[1] if len(sys.argv) > 1 else "kitchen"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
cfg = synth.ENCODER_CONFIGS[name]
dev = torch.device("cuda:0")
sd = {k: torch.tensor(v) for k, v in synth.make_encoder_state_dict(cfg["input_size"], cfg["rep_size"]).items()}
m = nf_b200.GEncoder(cfg["input_size"], cfg["rep_size"]); m.load_state_dict(sd); m = m.to(dev)
g, dout = synth.make_encoder_inputs(B, cfg["input_size"], cfg["rep_size"])
g = torch.tensor(g, device=dev); dout = torch.tensor(dout, device=dev)
for _ in range(2):
    for p in m.parameters():
        p.grad = None
    m(g).backward(dout)
torch.cuda.synchronize()
print("ok")
