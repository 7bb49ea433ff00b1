"""One training step + one autoregressive sampling pass of the causal-transformer flow (bench.py's C2t shape): ncu target.
python tools/ncu_target_tarflow.py [B]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
bench.TARFLOW["batch"] = B
dev = torch.device("cuda:0")
r = bench.TarflowRunner(dev, 0, 1, torch.empty(1, device=dev))
loss = r.train_step(r.x, r.y)
out = r.sample_step(r.x, r.y)
torch.cuda.synchronize()
print("ok", float(loss), float(out.abs().mean()))
