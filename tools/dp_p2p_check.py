"""Peer-memory all-reduce of the C ABI (csrc/dp.cu, nf_dp_p2p_*) against NCCL on the same buffers: equality within the
summation-order noise, bit-identical results on every rank, ragged / unaligned sizes, replay from a captured CUDA graph,
and device-timed latency of both paths per arena size.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/dp_p2p_check.py
"""
import ctypes, os, sys
os.environ.setdefault("NF_B200_DP_P2P_MIN_BYTES", "0")      # exercise the kernel at every size, whatever the world size
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import nf_b200
from nf_b200 import _lib
from nf_b200.module import _init_library_comm

lib = _lib.lib
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
_init_library_comm(None, dev)
active = lib.nf_dp_p2p_status()
if rank == 0:
    print(f"world={world} peer-memory path status after connect: {active}", flush=True)
assert active == 1, "peer-memory all-reduce did not connect"
st = torch.cuda.current_stream().cuda_stream


def ours(t, avg=True):
    _lib.check(lib.nf_dp_allreduce(t.data_ptr(), t.numel(), 1 if avg else 0, st), "nf_dp_allreduce")


def nccl(t, avg=True):
    _lib.check(lib.nf_dp_p2p_enable(0), "off")
    ours(t, avg)
    _lib.check(lib.nf_dp_p2p_enable(1), "on")


base = torch.empty((4 << 20) + 64, device=dev)
for n, off in ((1, 0), (3, 0), (5, 4), (1024, 0), (1027, 8), (65536, 0), (1048576, 0), (1048577, 0), (1050001, 4),
               (4194304, 0)):
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    src = torch.randn(n, device=dev, generator=g)
    base.fill_(-7.0)
    a = base[off:off + n]
    b = src.clone()
    for avg in (True, False):
        a.copy_(src); b.copy_(src)
        ours(a, avg); nccl(b, avg)
        torch.cuda.synchronize()
        err = (a - b).abs().max().item()
        gathered = [torch.empty_like(a) for _ in range(world)]
        dist.all_gather(gathered, a.contiguous())
        same = all(torch.equal(gathered[0], t) for t in gathered)
        if rank == 0:
            print(f"n={n} offset={off} avg={avg}: |peer - nccl| max {err:.2e}, identical on all ranks: {same}", flush=True)
        untouched = bool((base[:off] == -7.0).all()) and bool((base[off + n:] == -7.0).all())
        assert err < 1e-5 and same and untouched, (err, same, untouched)
assert lib.nf_dp_p2p_status() == 1

# replay from a captured graph: the call counter lives in device memory
n = 1 << 20
a = torch.zeros(n, device=dev)
side = torch.cuda.Stream()
with torch.cuda.stream(side):
    graph = torch.cuda.CUDAGraph()
    a.fill_(float(rank + 1))
    torch.cuda.synchronize()
    with torch.cuda.graph(graph, stream=side):
        _lib.check(lib.nf_dp_allreduce(a.data_ptr(), n, 1, side.cuda_stream), "nf_dp_allreduce (capture)")
    for it in range(5):
        a.fill_(float(rank + 1 + it))
        graph.replay()
        side.synchronize()
        want = sum(r + 1 + it for r in range(world)) / world
        assert torch.all(a == want), (it, a[:4], want)
if rank == 0:
    print("graph replay: 5 replays correct", flush=True)

# latency, device-timed; the two paths alternate so that clocks and neighbours are shared
for n in (1 << 16, 1 << 18, 1 << 20, 1 << 21, 1 << 22):
    a = torch.randn(n, device=dev)
    res = {}
    for name, fn in (("peer", ours), ("nccl", nccl), ("peer", ours), ("nccl", nccl)):
        for _ in range(5):
            fn(a)
        dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            fn(a)
        e1.record(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / 50 * 1e3], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        res.setdefault(name, []).append(t.item())
    ours(a); torch.cuda.synchronize()
    ns = (ctypes.c_uint64 * 6)()
    _lib.check(lib.nf_dp_p2p_timeline(ns), "nf_dp_p2p_timeline")
    d = [(ns[i + 1] - ns[i]) / 1e3 for i in range(5)]
    if rank == 0:
        print(f"all-reduce of {n * 4 / 2**20:.2f} MB over {world} ranks: peer-memory kernel "
              f"{min(res['peer']):.1f} us, NCCL {min(res['nccl']):.1f} us | CTA 0 phases (us): scatter {d[0]:.1f}, "
              f"flags {d[1]:.1f}, reduce {d[2]:.1f}, flags {d[3]:.1f}, gather {d[4]:.1f}", flush=True)
assert lib.nf_dp_p2p_status() == 1
dist.barrier()
lib.nf_dp_shutdown()
dist.destroy_process_group()
