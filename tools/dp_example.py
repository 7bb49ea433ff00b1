"""Two-or-more-rank data-parallel training step of encoder + flow over NCCL (torchrun), checking that both modules'
gradient arenas are averaged across ranks: every rank ends with identical gradients that equal the gradient of the
full batch computed on one rank.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/dp_example.py
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import nf_b200
from oracle import synth

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
B, D, R = 512, 8, 64
fsd = {k: torch.tensor(v) for k, v in synth.make_state_dict(D, 128, R, 3, seed=1).items()}
esd = {k: torch.tensor(v) for k, v in synth.make_encoder_state_dict(58, R, seed=2, hidden=128, n_hidden=2).items()}


def build(dp):
    flow = nf_b200.RealNVP(D, 128, R, 3, None, precision="fp32"); flow.load_state_dict(fsd); flow = flow.to(dev)
    enc = nf_b200.GEncoder(58, R, hidden=128, n_hidden=2); enc.load_state_dict(esd); enc = enc.to(dev)
    if dp:
        flow.enable_data_parallel(); enc.enable_data_parallel()
    return flow, enc


x, _ = synth.make_inputs(B * world, D, R, seed=5)
obs, _ = synth.make_encoder_inputs(B * world, 58, R, seed=6)
x, obs = torch.tensor(x, device=dev), torch.tensor(obs, device=dev)
flow, enc = build(True)
assert flow._dp_in_graph, "NCCL backend: the all-reduce runs inside the backward launch graph"
sl = slice(rank * B, (rank + 1) * B)
flow1, enc1 = build(False)
flow1.nll(x, enc1(obs)).backward()                   # the whole batch on this rank, no communication
for bucket in (0, 1, 3):                             # default (three buckets), one block per bucket, one bucket
    flow.dp_bucket_blocks = bucket
    for it in range(4):                              # eager, capture, replay, replay
        for p in list(flow.parameters()) + list(enc.parameters()):
            p.grad = None
        flow.nll(x[sl], enc(obs[sl])).backward()     # local shard; backward all-reduces both arenas
        worst = 0.0
        for (k, p), (_, q) in list(zip(flow.named_parameters(), flow1.named_parameters())) + \
                list(zip(enc.named_parameters(), enc1.named_parameters())):
            if p.grad is None:
                continue
            err = (p.grad - q.grad).abs().max().item() / max(q.grad.abs().max().item(), 1e-12)
            worst = max(worst, err)
        t = torch.tensor([worst], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            print(f"world={world} bucket_blocks={bucket} iteration {it}: max relative difference between DP-averaged "
                  f"and full-batch gradients: {t.item():.2e}", flush=True)
        assert t.item() < 5e-5
# a rank-local evaluation (denoise step: input gradient only) must not enter the collective
if rank == 0:
    a = x[:64].clone().requires_grad_(True)
    z, _, ld = flow(a, enc(obs[:64]).detach())
    (g,) = torch.autograd.grad((-0.5 * (z * z).sum(1) + ld).sum(), [a])
    torch.cuda.synchronize()
    print("rank-local input gradient ok", tuple(g.shape), flush=True)
if world <= 8:
    # the flow's arena went through the library's peer-memory kernel (csrc/dp.cu), and no peer ever timed out
    from nf_b200 import _lib
    status = _lib.lib.nf_dp_p2p_status()
    if rank == 0:
        print("peer-memory all-reduce status:", status, flush=True)
    assert status == 1 or int(os.environ.get("NF_B200_DP_P2P_MAX_BYTES", "1")) == 0, status
dist.barrier()
dist.destroy_process_group()
