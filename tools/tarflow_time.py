"""Device timing of the causal-transformer flow (bench.py's C2t shape): train step, forward, autoregressive inverse.
python tools/tarflow_time.py [B]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from nf_b200 import _lib

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
bench.TARFLOW["batch"] = B
dev = torch.device("cuda:0")
r = bench.TarflowRunner(dev, 0, 1, torch.empty(1, device=dev))

def timeit(fn, iters=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    _lib.lib.nf_launch_count(1)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters, _lib.lib.nf_launch_count(1) // iters

def fwd():
    with torch.no_grad(): r.model(r.x, r.y)
F = bench.tarflow_flops_fwd(bench.TARFLOW)
res = {"train": timeit(lambda: r.train_step(r.x, r.y)), "fwd": timeit(fwd), "rev": timeit(lambda: r.sample_step(r.x, r.y))}
print(f"C2t B={B}: " + " | ".join(f"{k} {v[0]:.3f} ms ({v[1]} launches) {B / v[0] * 1e3:.3e}/s "
                                  f"{F * (3 if k == 'train' else 1) * B / v[0] / 1e9:.1f} TF/s" for k, v in res.items()), flush=True)
