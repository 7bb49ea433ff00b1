import sys, os, ctypes, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nf_b200 import _lib
lib = _lib.lib
dev = torch.device("cuda:0")
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
B, D = 256, 8
xp = torch.randn(B, D, device=dev); s = torch.randn(B, D // 2, device=dev); t = torch.randn(B, D // 2, device=dev)
z = torch.empty(B, D, device=dev); ld = torch.zeros(B, device=dev)
args = (xp.data_ptr(), s.data_ptr(), t.data_ptr(), None, None, 0.0, B, D, z.data_ptr(), ld.data_ptr(), None, st)
for _ in range(10): lib.nf_affine_logdet_fwd(*args)
torch.cuda.synchronize()
N = 2000
t0 = time.perf_counter()
for _ in range(N): lib.nf_affine_logdet_fwd(*args)
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"nf_affine_logdet_fwd: enqueue {1e6*(t1-t0)/N:.2f} us/call, incl. drain {1e6*(t2-t0)/N:.2f} us/call")
A = torch.randn(256, 64, device=dev); W = torch.randn(64, 64, device=dev); C = torch.empty(256, 64, device=dev)
for eng in (0, 1):
    a = (A.data_ptr(), 64, W.data_ptr(), 64, None, C.data_ptr(), 64, 256, 64, 64, eng, st)
    for _ in range(10): lib.nf_linear(*a)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(N): lib.nf_linear(*a)
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"nf_linear engine {eng}: enqueue {1e6*(t1-t0)/N:.2f} us/call, incl. drain {1e6*(t2-t0)/N:.2f} us/call")
t0 = time.perf_counter()
for _ in range(N): z.add_(1.0)
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"torch add_: enqueue {1e6*(t1-t0)/N:.2f} us/call, incl. drain {1e6*(t2-t0)/N:.2f} us/call")
# whole module, tiny batch: CPU enqueue time vs device time
import nf_b200
from oracle import synth
cfg = dict(synth.CONFIGS["C2a"]); cfg.pop("ln_eps"); cfg.pop("H1")
m = nf_b200.RealNVP(cfg["D"], cfg["C"], cfg["R"], cfg["n"], None, precision="fp32").to(dev)
x = torch.randn(4096, 8, device=dev); y = torch.randn(4096, 64, device=dev)
def train():
    for p in m.parameters(): p.grad = None
    loss = -m.log_prob(x, y).mean(); loss.backward()
for _ in range(5): train()
torch.cuda.synchronize()
import cProfile, pstats
t0 = time.perf_counter()
for _ in range(20): train()
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"train C2a: enqueue {1e3*(t1-t0)/20:.3f} ms, incl drain {1e3*(t2-t0)/20:.3f} ms")
pr = cProfile.Profile(); pr.enable()
for _ in range(20): train()
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
