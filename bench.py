"""Benchmark of the flow hot path (BASELINE.json metric), one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config C2a] [--batch B]

A step = one pass of the hot path over one batch of synthetic input:
    z, outputs, log_dets = model(x, y); loss = -(log N(z) + log_dets).mean(); loss.backward()
(torch_nfbc_ant.py:267-272 without the optimizer; y requires grad like the encoder output it is in the
reference, so dy is computed; data-parallel gradient all-reduce included when N > 1), and -- reported under "sampling" -- model.reverse(randn(B, D), y) (torch_nfbc_ant.py:209-210).
Workload at N = 1: BASELINE.json configs[1] (goal-conditioned NF policy, obs+goal conditioning,
action 8, batch 4096): D=8, C=256, R=64, n=6 ("C2a" in SURVEY.md section 8d); every rank runs the same
per-GPU batch when N > 1 (weak scaling).  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np

METRIC = "flow train samples/sec (log_prob fwd+bwd) & sampled actions/sec, 1/2/4/8 B200"
DEFAULT_BATCH = {"C1a": 256, "C1b": 256, "C1c": 256, "C2a": 4096, "C3": 16384, "C3t": 16384, "C4": 8192, "C5": 8192}


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(hbm_gbs=p["hbm_gbs"], bf16_tflops=p["bf16_tflops"], bf16_sustained=p.get("bf16_tflops_sustained"),
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_sustained=1400.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.samples, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            f = [t.strip() for t in s.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm))


def config_of(name):
    from oracle import synth
    cfg = dict(synth.CONFIGS[name])
    return cfg


# --------------------------------------------------------------------------------------------------
# CPU arm: the reference's op sequence in stock PyTorch (oracle/torch_flow.py) on the host cores
# --------------------------------------------------------------------------------------------------
LOG2PI = float(np.log(2 * np.pi))


def cpu_step_fn(cfgname, B, seed=0):
    """The reference's CPU path: oracle/torch_flow.py, the stock-PyTorch restatement of the reference module (the same
    nn.Linear / LayerNorm / cat / exp ops + autograd; pinned to the reference-generated fixtures by
    tests/test_oracle_golden.py), on all host threads.  (The numpy oracle is the parity checker, not the baseline: it is
    ~6x slower than the reference's torch path on the same cores.)"""
    import torch
    from oracle import synth
    from oracle.torch_flow import TorchFlow
    torch.set_num_threads(os.cpu_count())      # torchrun exports OMP_NUM_THREADS=1
    cfg = config_of(cfgname)
    sd = synth.make_state_dict(cfg["D"], cfg["C"], cfg["R"], cfg["n"], seed=seed, H1=cfg["H1"])
    model = TorchFlow(cfg["D"], cfg["C"], cfg["R"], cfg["n"], first_hidden=cfg["H1"], ln_eps=cfg["ln_eps"])
    model.load_state_dict({k: torch.tensor(v) for k, v in sd.items()})
    x, y = synth.make_inputs(B, cfg["D"], cfg["R"], seed=1234)
    x, y = torch.tensor(x), torch.tensor(y)
    D = cfg["D"]
    plist = [p for p in model.parameters() if p.requires_grad]

    def train():
        for p in plist:
            p.grad = None
        yy = y.detach().requires_grad_(True)
        z, _, ld = model(x, yy)
        (-((-0.5 * (z * z).sum(1) - 0.5 * D * LOG2PI) + ld).mean()).backward()

    def sample():
        with torch.no_grad():
            model.reverse(x, y)

    return train, sample


def time_cpu(fn, min_seconds, max_iters):
    fn()
    t0 = time.perf_counter()
    it = 0
    while True:
        fn()
        it += 1
        el = time.perf_counter() - t0
        if el >= min_seconds or it >= max_iters:
            return el / it, it


def run_reference_arm(args):
    """--impl reference: the reference's CPU path (stock-PyTorch restatement of the reference module, all host threads)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfgname = args.config
    B_full = args.batch or DEFAULT_BATCH[cfgname]
    B = min(B_full, 4096)
    train, sample = cpu_step_fn(cfgname, B)
    for _ in range(args.warmup):
        train()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        train()
    per = (time.perf_counter() - t0) / args.steps
    ts, _ = time_cpu(sample, 2.0, 10)
    cores = os.cpu_count()
    value = B / per
    out = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": per * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{cfgname} coupling flow, batch {B} per step on the host CPU", **config_of(cfgname)},
        "sampling": {"value": B / ts, "unit": "actions/s"},
        "cpu_baseline": {"value": value, "unit": "samples/s", "cores": cores, "kind": "port",
                         "sample": f"{args.steps} steps of forward+backward at batch {B} (oracle/torch_flow.py: the "
                                   f"reference's op sequence in stock PyTorch on the host, {cores} threads)"},
        "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(out), flush=True)


# --------------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    import nf_b200
    from nf_b200 import _lib
    from oracle import synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback for the product path)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    cfgname = args.config
    cfg = config_of(cfgname)
    B = args.batch or DEFAULT_BATCH[cfgname]
    D, R = cfg["D"], cfg["R"]
    sd = synth.make_state_dict(D, cfg["C"], R, cfg["n"], seed=0, H1=cfg["H1"])
    model = nf_b200.RealNVP(D, cfg["C"], R, cfg["n"], None, first_hidden=cfg["H1"], ln_eps=cfg["ln_eps"],
                            precision=args.precision)
    model.load_state_dict({k: torch.tensor(v) for k, v in sd.items()})
    model = model.to(dev)
    if world > 1:
        model.enable_data_parallel()
        if os.environ.get("NF_B200_DP_BUCKET_BLOCKS"):     # experiment knob: overlapped bucketed all-reduce
            model.dp_bucket_blocks = int(os.environ["NF_B200_DP_BUCKET_BLOCKS"])
    xh, yh = synth.make_inputs(B, D, R, seed=1234 + rank)
    x_pin, y_pin = torch.tensor(xh).pin_memory(), torch.tensor(yh).pin_memory()
    x, y = x_pin.to(dev), y_pin.to(dev)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)   # > 126 MB L2

    plist = [p for p in model.parameters() if p.requires_grad]

    def train_step(xd, yd):
        model._packed_key = None          # parameters change every real training step: rebuild the tables
        for p in plist:                   # optimizer.zero_grad(set_to_none=True)
            p.grad = None
        yd = yd.detach().requires_grad_(True)   # y is the encoder output in the reference (torch_nfbc_ant.py:267): dy is part of the step
        z, _, ld = model(xd, yd)
        loss = -((-0.5 * (z * z).sum(1) - 0.5 * D * LOG2PI) + ld).mean()
        loss.backward()
        return loss

    def train_step_nll(xd, yd):
        # the same step with the library's fused loss (RealNVP.nll: SURVEY section 8 row f3) instead of the ~8 torch
        # kernels + autograd nodes of the reference's loss expression
        model._packed_key = None
        for p in plist:
            p.grad = None
        loss = model.nll(xd, yd.detach().requires_grad_(True))
        loss.backward()
        return loss

    def sample_step(xd, yd):
        with torch.no_grad():
            return model.reverse(xd, yd)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup, e2e=False):
        def one():
            if e2e:
                xd = x_pin.to(dev, non_blocking=True)   # H2D of this step's inputs from pinned memory
                yd = y_pin.to(dev, non_blocking=True)
                out = fn(xd, yd)
                _ = float(out.reshape(-1)[0].item()) if out.dim() else float(out.item())   # D2H of the result
            else:
                fn(x, y)
        for _ in range(warmup):                         # same code path as the timed steps
            flush.zero_()
            one()
        barrier()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        wall0 = time.perf_counter()
        for a, b in ev:
            flush.zero_()                               # evict L2 between timed iterations (outside the events)
            a.record()
            one()
            b.record()
        barrier()
        wall = time.perf_counter() - wall0
        per = [a.elapsed_time(b) for a, b in ev]
        ms = sum(per)
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)    # max over ranks
        step_stats.append((float(np.median(per)), float(max(per))))
        if os.environ.get("BENCH_DEBUG"):
            print("per-step ms:", " ".join(f"{v:.3f}" for v in per), file=sys.stderr, flush=True)
        return t.item() / steps, wall

    step_stats = []
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    _lib.lib.nf_launch_count(1)
    gstats = [_lib.graph_stats()]
    ms_train, _ = timed(train_step, args.steps, args.warmup)
    launches = _lib.lib.nf_launch_count(1) // (args.steps + args.warmup)
    gstats.append(_lib.graph_stats())
    ms_sample, _ = timed(sample_step, args.steps, args.warmup)
    gstats.append(_lib.graph_stats())
    ms_e2e, _ = timed(train_step, args.steps, args.warmup, e2e=True)
    gstats.append(_lib.graph_stats())
    ms_e2e_sample, _ = timed(sample_step, args.steps, args.warmup, e2e=True)
    gstats.append(_lib.graph_stats())
    ms_e2e_nll, _ = timed(train_step_nll, args.steps, args.warmup, e2e=True)   # informational, not the headline
    leg_graphs = {name: dict(zip(("replays", "captures", "eager_calls"), (b[i] - a[i] for i in range(3))))
                  for name, a, b in zip(("train", "sample", "e2e_train", "e2e_sample"), gstats[:-1], gstats[1:])}
    # the four timed legs above are short (tens of ms at the default K): keep the same kernels running until
    # nvidia-smi (100 ms period) has seen them under load; these extra steps are not part of any reported time
    # (a FIXED count derived from the max-reduced step time: with world > 1 every train_step contains the gradient
    # all-reduce, so all ranks must run the same number of steps -- a wall-clock loop deadlocks the process group)
    for _ in range(int(min(4000, max(10, 1500.0 / max(ms_train, 1e-3))))):
        train_step(x, y)
    torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None
    replays, captures, eager = _lib.graph_stats()

    # ---- roofline pass: per-kernel-class CUDA-event timing inside the same step (profiling on) ----
    _lib.lib.nf_prof_enable(1)
    _lib.prof_collect()
    for _ in range(3):
        flush.zero_()
        train_step(x, y)
    torch.cuda.synchronize()
    prof = _lib.prof_collect()
    _lib.lib.nf_prof_enable(0)

    if rank != 0:
        # wait for rank 0 (CPU baseline, JSON line) and leave the process group together with it
        dist.barrier()
        dist.destroy_process_group()
        return

    peaks = load_peaks()
    F_fwd = synth.flops_fwd_per_sample(**cfg)
    total_B = B * world
    value = total_B / (ms_train * 1e-3)
    tot_ms = sum(v[0] for v in prof.values()) or 1.0
    dom = max(("tc_gemm", "tc_wgrad", "simt_gemm"), key=lambda k: prof[k][0])
    d_ms, d_work, d_n = prof[dom]
    achieved = d_work / (d_ms * 1e-3) / 1e12 if d_ms > 0 else 0.0
    breakdown = {k: {"ms_per_step": v[0] / 3, "launches_per_step": v[2] // 3,
                     "share": v[0] / tot_ms} for k, v in prof.items() if v[2]}
    passes = 3 if args.precision in ("fp32", "tf32x3") else 1
    roofline = {
        "bound": "tensor", "kernel": {"tc_gemm": "k_tc_gemm<BN,STAGES,false> (tcgen05 Linear fwd/dgrad)",
                                      "tc_wgrad": "k_tc_gemm<BN,STAGES,true> (tcgen05 weight gradient)",
                                      "simt_gemm": "k_gemm (FFMA)"}[dom],
        "achieved": achieved, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
        "frac": achieved / peaks["bf16_tflops"],
        # dram__bytes_read.sum + dram__bytes_write.sum of one launch of the class's heaviest member, the fused
        # layer-2 Linear (M=4096, N=256, K=256, both nets) -- ncu --set full, profiles/r1_ncu_full_forward_block_C2a.txt
        # (cold cache under ncu; algorithmic bytes: 8.4 MB in + 0.5 MB weights + 8.4 MB out)
        "traffic": 9.52e6 if cfgname == "C2a" and B == 4096 else None, "traffic_unit": "bytes/launch",
        "peak_source": peaks["source"] + ", dense bf16 burst",
        "note": (f"algorithmic FLOPs (2MNK per launch, split-precision passes NOT counted) / mean launch duration over "
                 f"{d_n} launches; the kernel issues {passes} kind::tf32 MMAs per product at half the bf16 rate, so the "
                 f"tensor-pipe-equivalent rate is {achieved * passes * 2:.0f} TFLOP/s bf16-equivalent"),
        "share_of_step": d_ms / tot_ms,
        "whole_step": {"achieved": 3 * F_fwd * B / (ms_train * 1e-3) / 1e12, "unit": "TFLOP/s",
                       "frac": 3 * F_fwd * B / (ms_train * 1e-3) / 1e12 / peaks["bf16_tflops"]},
        "breakdown": breakdown,
    }

    # ---- CPU baseline: the reference's op sequence in stock PyTorch on this box's host cores, bounded sample,
    # at N = 1 only (the other ranks of a multi-GPU run would just wait for it) -----------------------------------
    cpu_baseline = None
    if world == 1:
        Bc = min(B, 4096)
        ctrain, csample = cpu_step_fn(cfgname, Bc)
        per, iters = time_cpu(ctrain, 10.0, 200)
        cores = os.cpu_count()
        cpu_baseline = {"value": Bc / per, "unit": "samples/s", "cores": cores, "kind": "port",
                        "sample": f"{iters} steps of forward+backward at batch {Bc} of the same config "
                                  f"(oracle/torch_flow.py: the reference's op sequence in stock PyTorch on the host, "
                                  f"{cores} threads, {per * iters:.1f} s)"}

    bytes_in = x_pin.numel() * 4 + y_pin.numel() * 4
    out = {
        "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_train, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32 (tcgen05 3xTF32 split, fp32 accumulate)" if passes == 3 else args.precision,
        "data": "synthetic",
        "config": {"workload": f"{cfgname}: coupling flow D={D} C={cfg['C']} R={R} n={cfg['n']} H1={cfg['H1']}, "
                               f"batch {B} per GPU, log_prob forward+backward",
                   "global_batch": total_batch_str(total_B), "parallelism": f"dp{world}",
                   "l2": "flushed between timed iterations (256 MB write)", "precision": args.precision},
        "sampling": {"value": total_B / (ms_sample * 1e-3), "unit": "actions/s", "ms_per_step": ms_sample},
        "e2e": {"value": total_B / (ms_e2e * 1e-3), "unit": "samples/s", "h2d_bytes_per_step": bytes_in,
                "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e, "ms_median_max": list(step_stats[2]),
                "sampling": {"value": total_B / (ms_e2e_sample * 1e-3), "unit": "actions/s",
                             "d2h_bytes_per_step": 4, "ms_median_max": list(step_stats[3])},
                "with_fused_loss": {"value": total_B / (ms_e2e_nll * 1e-3), "unit": "samples/s",
                                    "ms_per_step": ms_e2e_nll,
                                    "note": "same end-to-end step with loss = model.nll(x, y) (library NLL kernels) "
                                            "instead of the reference's torch-side loss expression"}},
        "gpu_launches": int(launches * args.steps),
        "gpu_launches_per_step": int(launches),
        "launch_graphs": {"replays": replays, "captures": captures, "eager_calls": eager, "per_leg": leg_graphs,
                          "note": "the library replays its kernel sequence as a cached CUDA graph; gpu_launches "
                                  "counts the library's kernels inside the replayed graphs"},
        "train_ms_median_max": list(step_stats[0]),
        "clocks": clocks,
        "roofline": roofline,
        "cpu_baseline": cpu_baseline,
    }
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def total_batch_str(n):
    return int(n)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C2a")
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--precision", default="fp32", choices=["fp32", "fp32_simt"])
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
