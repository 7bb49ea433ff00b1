"""Benchmark of the flow hot path (BASELINE.json metric), one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--configs C1a,C2a,...]

A step = one pass of the hot path over one batch of synthetic input:
    z, outputs, log_dets = model(x, y); loss = -(model.prior.log_prob(z) + log_dets).mean(); loss.backward()
(torch_nfbc_ant.py:267-272 without the optimizer; y requires grad like the encoder output it is in the reference, so
dy is computed; the data-parallel gradient all-reduce is included when N > 1) and -- reported under "sampling" --
model.reverse(randn(B, D), y) (torch_nfbc_ant.py:209-210).

Headline (`value`, `e2e`, `roofline`): BASELINE.json configs[1] (goal-conditioned NF policy, obs+goal conditioning,
action 8, batch 4096): D=8, C=256, R=64, n=6 ("C2a" in SURVEY.md section 8d), 4096 rows per GPU (weak scaling),
EXACTLY --steps timed steps.  `per_config` carries every BASELINE configuration at its BASELINE batch in the same
JSON line (train / sampling / e2e / roofline per config; >= 100 timed steps for the small ones):
    C1a, C1c  ant NF-BC flow (D = 8 / 400), batch 256 per GPU                                    weak
    C2a       as the headline, 100 steps                                                         weak
    C3        latent-actor flow of nf.py (H1 = C + R, eps 1e-6), 16384 per GPU, DP all-reduce    weak
    C4        kitchen NF-BC sampling, 65536 parallel envs sharded over the GPUs (no collective)  strong
    C5        stress flow (24 blocks, width 1024, D = 32), global batch 65536 sharded            strong
    C2t       causal-transformer flow (torch_nfgcbc.py:119 shape: 8 tokens, width 256, 4 x 4 layers), batch 4096   weak
Before anything is timed, one forward of every config is compared with the float64 checker (the reference's op
sequence in stock PyTorch, oracle/torch_flow.py, on the GPU) and the run aborts on a mismatch.
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import math
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
HEADLINE = "C2a"
# name -> (batch, "per_gpu" | "global", legs, timed steps: "K" = --steps, int = at least that many)
PLAN = {
    "C1a": dict(batch=256, shard="per_gpu", train=True, steps=100),
    "C1c": dict(batch=256, shard="per_gpu", train=True, steps=100),
    "C2a": dict(batch=4096, shard="per_gpu", train=True, steps=100),
    "C3": dict(batch=16384, shard="per_gpu", train=True, steps=100),
    "C4": dict(batch=65536, shard="global", train=False, steps="K"),
    "C5": dict(batch=65536, shard="global", train=True, steps="K"),
}
CPU_SAMPLE_ROWS = {"C1a": 256, "C1c": 256, "C2a": 4096, "C3": 4096, "C4": 2048, "C5": 1024}
LOG2PI = float(np.log(2 * np.pi))


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(hbm_gbs=p["hbm_gbs"], bf16_tflops=p["bf16_tflops"], bf16_sustained=p.get("bf16_tflops_sustained"),
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_sustained=1400.0, source="fallback (B200_PROFILING.md)")


def load_traffic():
    """dram bytes per launch of the dominant kernels, parsed from committed `ncu --set full` captures by
    tools/ncu_traffic.py (profiles/ncu_traffic.json: {"<config>:<kernel class>": {"bytes": ..., "source": ...}})."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(path):
        return {}
    with open(path) as f:
        return json.load(f)


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.samples, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", os.environ.get("NF_B200_SMI_MS", "100")],
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
        # median over the samples taken under load (an idle GPU between legs parks at a low clock)
        busy = [v for v in sm if mx and v >= 0.6 * max(mx)] or sm
        return dict(sm_mhz=float(np.median(busy)) if busy else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm))


def config_of(name):
    from oracle import synth
    return dict(synth.CONFIGS[name])


def batch_per_gpu(name, world):
    p = PLAN[name]
    return p["batch"] if p["shard"] == "per_gpu" else max(1, p["batch"] // world)


def config_dict(name, world, precision="fp32"):
    """The `config` object BOTH arms print for the headline (identical dicts; the reference arm describes its bounded
    CPU sample in cpu_baseline.sample)."""
    cfg = config_of(name)
    B = batch_per_gpu(name, world)
    return {"workload": f"{name}: coupling flow D={cfg['D']} C={cfg['C']} R={cfg['R']} n={cfg['n']} H1={cfg['H1']}, "
                        f"batch {B} per GPU, log_prob forward+backward",
            **cfg, "batch_per_gpu": B, "global_batch": B * world, "parallelism": f"dp{world}",
            "l2": "flushed between timed iterations (256 MB write)", "precision": precision}


# --------------------------------------------------------------------------------------------------
# CPU arm: the reference's op sequence in stock PyTorch (oracle/torch_flow.py) on the host cores
# --------------------------------------------------------------------------------------------------
def reference_module():
    """The UNMODIFIED reference module il-and-cil/utils/torch_fcnf.py, compiled to bytecode by __graft_entry__.build() into
    the git-ignored oracle/_ref/ (build outputs only; it travels with the gpurun snapshot -- /root/reference does not exist
    on the GPU box).  None when absent."""
    path = os.path.join(ROOT, "oracle", "_ref", "torch_fcnf.bytecode")
    if not os.path.exists(path):
        return None
    import importlib.machinery
    import importlib.util
    try:
        loader = importlib.machinery.SourcelessFileLoader("nf_reference_torch_fcnf", path)
        spec = importlib.util.spec_from_loader("nf_reference_torch_fcnf", loader)
        mod = importlib.util.module_from_spec(spec)
        loader.exec_module(mod)
        return mod
    except Exception:       # bytecode of another interpreter version: fall back to the port
        return None


def cpu_kind(cfgname):
    """"reference" when the CPU arm of this configuration runs the staged reference file itself (torch widths and eps only:
    the flax twin's H1 = C + R / eps 1e-6 network exists upstream only in JAX), else "port" (oracle/torch_flow.py)."""
    cfg = config_of(cfgname)
    return "reference" if (cfg["H1"] == cfg["C"] and cfg["ln_eps"] == 1e-5 and reference_module() is not None) else "port"


def cpu_what(cfgname):
    if cpu_kind(cfgname) == "reference":
        return ("the unmodified reference module il-and-cil/utils/torch_fcnf.py compiled into oracle/_ref/ by build(), with the "
                "scripts' MultivariateNormal prior and loss line")
    return ("oracle/torch_flow.py, the reference's op sequence restated in stock PyTorch: the staged reference file is "
            "absent or this configuration uses the flax twin's widths")


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
    D = cfg["D"]
    use_ref = cpu_kind(cfgname) == "reference"
    if use_ref:
        # the reference's own module and its own loss line: model.prior is the full-covariance MultivariateNormal the
        # scripts build (torch_nfbc_ant.py:146-147, 267-270)
        ref = reference_module()
        prior = torch.distributions.MultivariateNormal(torch.zeros(D), torch.eye(D))
        model = ref.RealNVP(D, cfg["C"], cfg["R"], cfg["n"], prior)
    else:
        model = TorchFlow(cfg["D"], cfg["C"], cfg["R"], cfg["n"], first_hidden=cfg["H1"], ln_eps=cfg["ln_eps"])
    model.load_state_dict({k: torch.tensor(v) for k, v in sd.items()})
    x, y = synth.make_inputs(B, cfg["D"], cfg["R"], seed=1234)
    x, y = torch.tensor(x), torch.tensor(y)
    plist = [p for p in model.parameters() if p.requires_grad]

    def train():
        for p in plist:
            p.grad = None
        yy = y.detach().requires_grad_(True)
        z, _, ld = model(x, yy)
        if use_ref:
            (-(model.prior.log_prob(z) + ld).mean()).backward()          # torch_nfbc_ant.py:269-271
        else:
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
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    B = min(batch_per_gpu(HEADLINE, world), 4096)
    train, sample = cpu_step_fn(HEADLINE, B)
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
        "config": config_dict(HEADLINE, world),
        "sampling": {"value": B / ts, "unit": "actions/s"},
        "cpu_baseline": {"value": value, "unit": "samples/s", "cores": cores, "kind": cpu_kind(HEADLINE),
                         "sample": f"{args.steps} steps of forward+backward at batch {B} on the host CPU, {cores} threads ("
                                   + cpu_what(HEADLINE) + ")"},
        "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(out), flush=True)


# --------------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------------
KERNEL_NAMES = {"tc_gemm": "k_tc_gemm<BN,STAGES,false> (tcgen05 Linear fwd/dgrad)",
                "tc_wgrad": "k_tc_gemm<BN,STAGES,true> (tcgen05 weight gradient)",
                "simt_gemm": "k_gemm (FFMA)"}


class Runner:
    """One BASELINE configuration on this rank's GPU."""

    def __init__(self, name, precision, dev, rank, world, flush):
        import torch
        import nf_b200
        from oracle import synth
        self.torch, self.name, self.dev, self.rank, self.world, self.flush = torch, name, dev, rank, world, flush
        self.cfg = cfg = config_of(name)
        self.B = B = batch_per_gpu(name, world)
        self.D, self.R = cfg["D"], cfg["R"]
        self.sd = synth.make_state_dict(cfg["D"], cfg["C"], cfg["R"], cfg["n"], seed=0, H1=cfg["H1"])
        # the prior the scripts pass to the constructor (torch_nfbc_ant.py:146-147), as the library's one-kernel N(0, I)
        model = nf_b200.RealNVP(cfg["D"], cfg["C"], cfg["R"], cfg["n"], nf_b200.StandardNormal(cfg["D"], dev),
                                first_hidden=cfg["H1"], ln_eps=cfg["ln_eps"], precision=precision)
        model.load_state_dict({k: torch.tensor(v) for k, v in self.sd.items()})
        self.model = model.to(dev)
        if world > 1 and PLAN[name]["train"]:
            ig = os.environ.get("NF_B200_DP_IN_GRAPH")          # experiment knob: 0 = torch.distributed after backward
            self.model.enable_data_parallel(in_graph=None if ig is None else bool(int(ig)))
            if os.environ.get("NF_B200_DP_BUCKET_BLOCKS"):     # experiment knob: bucketed all-reduce
                self.model.dp_bucket_blocks = int(os.environ["NF_B200_DP_BUCKET_BLOCKS"])
        xh, yh = synth.make_inputs(B, cfg["D"], cfg["R"], seed=1234 + rank)
        self.x_pin, self.y_pin = torch.tensor(xh).pin_memory(), torch.tensor(yh).pin_memory()
        self.x, self.y = self.x_pin.to(dev), self.y_pin.to(dev)
        self.plist = [p for p in self.model.parameters() if p.requires_grad]
        self.step_stats = {}

    # ---- the steps ----
    def train_step(self, xd, yd):
        model, D = self.model, self.D
        model._packed_key = None          # parameters change every real training step: rebuild the tables
        yd = yd.detach().requires_grad_(True)   # y is the encoder output in the reference (torch_nfbc_ant.py:267)
        z, _, ld = model(xd, yd)
        loss = -(model.prior.log_prob(z) + ld).mean()          # the reference's loss line, torch_nfbc_ant.py:269-270
        for p in self.plist:              # optimizer.zero_grad(set_to_none=True): between the loss and backward(),
            p.grad = None                 # where the reference has it (torch_nfbc_ant.py:272)
        loss.backward()
        return loss

    def train_step_nll(self, xd, yd):
        # the same step with the library's fused loss (RealNVP.nll: SURVEY section 8 row f3)
        model = self.model
        model._packed_key = None
        loss = model.nll(xd, yd.detach().requires_grad_(True))
        for p in self.plist:
            p.grad = None
        loss.backward()
        return loss

    def sample_step(self, xd, yd):
        with self.torch.no_grad():
            return self.model.reverse(xd, yd)

    # ---- parity gate ----
    def parity_check(self, rows=2048, gate=True):
        """One forward of this configuration against the float64 checker (oracle/torch_flow.py on the GPU) on the
        first `rows` rows of the step's inputs; tolerance of BASELINE.md section 3 (gate=False: report only)."""
        torch = self.torch
        from oracle.torch_flow import TorchFlow
        cfg, n = self.cfg, min(rows, self.B)
        x, y = self.x[:n].contiguous(), self.y[:n].contiguous()
        with torch.no_grad():
            z, _, ld = self.model(x, y)
        res = {}
        ref = {}
        for dt in (torch.float64, torch.float32):
            m = TorchFlow(cfg["D"], cfg["C"], cfg["R"], cfg["n"], first_hidden=cfg["H1"], ln_eps=cfg["ln_eps"])
            m.load_state_dict({k: torch.tensor(v) for k, v in self.sd.items()})
            m = m.to(self.dev).to(dt)
            with torch.no_grad():
                zz, _, ll = m(x.to(dt), y.to(dt))
            ref[dt] = (zz.double(), ll.double())
            del m

        def rel(a, b):
            return float((a.double() - b).abs().max()) / max(float(b.abs().max()), 1e-30)
        for name, ours, i in (("z", z, 0), ("log_dets", ld, 1)):
            e, noise = rel(ours, ref[torch.float64][i]), rel(ref[torch.float32][i], ref[torch.float64][i])
            res[name] = {"err": e, "ref32_err": noise, "tol": max(1e-5, 2 * noise)}
            if gate and not e <= max(1e-5, 2 * noise):
                raise SystemExit(f"bench.py: parity check failed for {self.name} ({name}: {e:.3e} > {max(1e-5, 2 * noise):.3e})")
        res["rows"] = n
        return res

    # ---- timing ----
    def barrier(self):
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
        self.torch.cuda.synchronize()

    def timed(self, tag, fn, steps, warmup, e2e=False):
        torch = self.torch

        def one():
            if e2e:
                xd = self.x_pin.to(self.dev, non_blocking=True)   # H2D of this step's inputs from pinned memory
                yd = self.y_pin.to(self.dev, non_blocking=True)
                out = fn(xd, yd)
                _ = float(out.reshape(-1)[0].item()) if out.dim() else float(out.item())   # D2H of the result
            else:
                fn(self.x, self.y)
        for _ in range(warmup):                         # same code path as the timed steps
            self.flush.zero_()
            one()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        self.barrier()
        # (no work between the barrier and the first timed step: a pause there -- a gc.collect() was tried -- lets the
        # idle GPU fall back and the first step then takes twice as long)
        for a, b in ev:
            self.flush.zero_()                          # evict L2 between timed iterations (outside the events)
            a.record()
            one()
            b.record()
        self.barrier()
        per = [a.elapsed_time(b) for a, b in ev]
        t = torch.tensor([sum(per)], dtype=torch.float64, device=self.dev)
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(t, op=dist.ReduceOp.MAX)    # max over ranks
        self.step_stats[tag] = (float(np.median(per)), float(max(per)))
        return t.item() / steps

    def roofline(self, ms_step, peaks, traffic, train):
        """Per-kernel-class CUDA-event timing inside the same step (profiling on: eager launches)."""
        torch = self.torch
        from nf_b200 import _lib
        from oracle import synth
        _lib.lib.nf_prof_enable(1)
        _lib.prof_collect()
        reps = 3
        for _ in range(reps):
            self.flush.zero_()
            (self.train_step if train else self.sample_step)(self.x, self.y)
        torch.cuda.synchronize()
        prof = _lib.prof_collect()
        _lib.lib.nf_prof_enable(0)
        tot_ms = sum(v[0] for v in prof.values()) or 1.0
        dom = max(("tc_gemm", "tc_wgrad", "simt_gemm"), key=lambda k: prof[k][0])
        d_ms, d_work, d_n = prof[dom]
        achieved = d_work / (d_ms * 1e-3) / 1e12 if d_ms > 0 else 0.0
        F_fwd = synth.flops_fwd_per_sample(**self.cfg)
        step_flops = (3 if train else 1) * F_fwd * self.B
        tr = traffic.get(f"{self.name}:{dom}")
        return {
            "bound": "tensor", "kernel": KERNEL_NAMES[dom], "achieved": achieved, "peak": peaks["bf16_tflops"],
            "unit": "TFLOP/s", "frac": achieved / peaks["bf16_tflops"],
            "traffic": tr["bytes"] if tr else None, "traffic_unit": "bytes/launch",
            "traffic_source": tr["source"] if tr else None,
            "peak_source": peaks["source"] + ", dense bf16 burst",
            "launches_timed": int(d_n), "share_of_step": d_ms / tot_ms,
            "whole_step": {"achieved": step_flops / (ms_step * 1e-3) / 1e12, "unit": "TFLOP/s",
                           "frac": step_flops / (ms_step * 1e-3) / 1e12 / peaks["bf16_tflops"],
                           "what": "train step (3 F_fwd B)" if train else "sampling pass (F_fwd B)"},
            "breakdown": {k: {"ms_per_step": v[0] / reps, "launches_per_step": v[2] // reps, "share": v[0] / tot_ms}
                          for k, v in prof.items() if v[2]},
        }

    def free(self):
        self.model = None
        self.x = self.y = self.x_pin = self.y_pin = None
        self.torch.cuda.empty_cache()


# --------------------------------------------------------------------------------------------------
# causal-transformer flow: the transformer shape of BASELINE configs[1]
# --------------------------------------------------------------------------------------------------
TARFLOW = dict(A=8, C=256, R=64, hd=64, nb=4, L=4, E=4, batch=4096)     # torch_nfgcbc.py:74-82 defaults, batch of configs[1]


def tarflow_flops_fwd(c):
    A, C, R, H, E = c["A"], c["C"], c["R"], c["C"] // c["hd"], c["E"]
    layer = A * (2 * C * 3 * C + 2 * C * C + 2 * 2 * C * E * C) + H * 2 * 2 * c["hd"] * A * (A + 1) // 2
    return c["nb"] * (2 * R * C + c["L"] * layer + A * 2 * 2 * C)


def tarflow_models(c, seed=0):
    import torch
    from oracle import torch_tarflow as tt
    torch.manual_seed(seed)
    ref = tt.RealNVP(c["A"], c["C"], c["R"], c["hd"], c["nb"], c["L"], None, expansion=c["E"])
    tt.perturb_output_heads(ref, seed=seed + 1)
    return ref


class TarflowRunner(Runner):
    """The transformer flow on this rank's GPU; same timing / e2e machinery as the coupling-flow configurations."""

    def __init__(self, dev, rank, world, flush):
        import torch
        import nf_b200
        c = TARFLOW
        self.torch, self.name, self.dev, self.rank, self.world, self.flush = torch, "C2t", dev, rank, world, flush
        self.c, self.B, self.D, self.R = c, c["batch"], c["A"], c["R"]
        self.ref = tarflow_models(c)
        model = nf_b200.TransformerRealNVP(c["A"], c["C"], c["R"], c["hd"], c["nb"], c["L"],
                                           nf_b200.StandardNormal(c["A"], dev), expansion=c["E"])
        model.load_state_dict(self.ref.state_dict())
        self.model = model.to(dev)
        if world > 1:
            self.model.enable_data_parallel()
        g = torch.Generator().manual_seed(1234 + rank)
        B = self.B
        xh = torch.rand(B, c["A"], 1, generator=g) * 2 - 1 + 0.1 * torch.randn(B, c["A"], 1, generator=g)
        yh = torch.randn(B, 1, c["R"], generator=g)
        self.x_pin, self.y_pin = xh.pin_memory(), yh.pin_memory()
        self.x, self.y = self.x_pin.to(dev), self.y_pin.to(dev)
        self.plist = [p for p in self.model.parameters() if p.requires_grad]
        self.step_stats = {}

    def config(self):
        c = self.c
        return {"workload": f"C2t: causal-transformer flow (torch_nfgcbc.py:119 shape) A={c['A']} channels={c['C']} "
                            f"rep={c['R']} head_dim={c['hd']} blocks={c['nb']} layers_per_block={c['L']}, batch {self.B} per GPU, "
                            f"log_prob forward+backward; sampling = the autoregressive inverse",
                **{k: v for k, v in c.items() if k != "batch"}, "batch_per_gpu": self.B, "global_batch": self.B * self.world,
                "parallelism": f"dp{self.world}", "l2": "flushed between timed iterations (256 MB write)",
                "precision": "fp32", "parity": "unpinned upstream (source absent); checker oracle/torch_tarflow.py"}

    def train_step(self, xd, yd):
        model, D = self.model, self.D
        yd = yd.detach().requires_grad_(True)       # the encoder output (torch_nfgcbc.py:141)
        z, _, ld = model(xd, yd)
        loss = -(model.prior.log_prob(z.squeeze(-1)) + ld).mean()                   # torch_nfgcbc.py:142
        for p in self.plist:                        # optimizer.zero_grad() sits between loss and backward()
            p.grad = None
        loss.backward()
        return loss

    def parity_check(self, rows=512, gate=True):
        import copy
        torch = self.torch
        n = min(rows, self.B)
        x, y = self.x[:n].contiguous(), self.y[:n].contiguous()
        with torch.no_grad():
            z, _, ld = self.model(x, y)
        ref = {}
        for dt in (torch.float64, torch.float32):
            m = copy.deepcopy(self.ref).to(self.dev).to(dt)
            with torch.no_grad():
                zz, _, ll = m(x.to(dt), y.to(dt))
            ref[dt] = (zz.double(), ll.double())

        def rel(a, b):
            return float((a.double() - b).abs().max()) / max(float(b.abs().max()), 1e-30)
        res = {"rows": n}
        for name, ours, i in (("z", z, 0), ("log_dets", ld, 1)):
            e, noise = rel(ours, ref[torch.float64][i]), rel(ref[torch.float32][i], ref[torch.float64][i])
            res[name] = {"err": e, "ref32_err": noise, "tol": max(1e-5, 2 * noise)}
            if gate and not e <= max(1e-5, 2 * noise):
                raise SystemExit(f"bench.py: parity check failed for C2t ({name}: {e:.3e})")
        return res

    def roofline(self, ms_step, peaks, traffic, train):
        torch = self.torch
        from nf_b200 import _lib
        _lib.lib.nf_prof_enable(1)
        _lib.prof_collect()
        reps = 3
        for _ in range(reps):
            self.flush.zero_()
            (self.train_step if train else self.sample_step)(self.x, self.y)
        torch.cuda.synchronize()
        prof = _lib.prof_collect()
        _lib.lib.nf_prof_enable(0)
        tot_ms = sum(v[0] for v in prof.values()) or 1.0
        dom = max(("tc_gemm", "tc_wgrad", "simt_gemm"), key=lambda k: prof[k][0])
        d_ms, d_work, d_n = prof[dom]
        achieved = d_work / (d_ms * 1e-3) / 1e12 if d_ms > 0 else 0.0
        step_flops = (3 if train else 1) * tarflow_flops_fwd(self.c) * self.B
        return {"bound": "tensor", "kernel": KERNEL_NAMES[dom], "achieved": achieved, "peak": peaks["bf16_tflops"],
                "unit": "TFLOP/s", "frac": achieved / peaks["bf16_tflops"], "traffic": None,
                "launches_timed": int(d_n),
                "share_of_step": d_ms / tot_ms,
                "breakdown": {k: {"ms_per_step": v[0] / reps, "launches_per_step": v[2] // reps, "share": v[0] / tot_ms}
                              for k, v in prof.items() if v[2]},
                "whole_step": {"achieved": step_flops / (ms_step * 1e-3) / 1e12, "unit": "TFLOP/s",
                               "frac": step_flops / (ms_step * 1e-3) / 1e12 / peaks["bf16_tflops"],
                               "what": "train step (3 F_fwd B)" if train else "sampling pass (F_fwd B)"}}


def tarflow_cpu_baseline(Bc=256):
    """The checker (stock PyTorch, all host threads) as the CPU baseline of the transformer flow."""
    import torch
    torch.set_num_threads(os.cpu_count())
    c = TARFLOW
    m = tarflow_models(c)
    g = torch.Generator().manual_seed(1234)
    x = torch.rand(Bc, c["A"], 1, generator=g) * 2 - 1
    y = torch.randn(Bc, 1, c["R"], generator=g)
    plist = list(m.parameters())

    def train():
        for p in plist:
            p.grad = None
        z, _, ld = m(x, y.detach().requires_grad_(True))
        (-((-0.5 * (z * z).sum((1, 2)) - 0.5 * c["A"] * LOG2PI) + ld).mean()).backward()

    def sample():
        with torch.no_grad():
            m.reverse(x, y)
    per, iters = time_cpu(train, 2.5, 20)
    per_s, iters_s = time_cpu(sample, 2.5, 10)
    cores = os.cpu_count()
    return {"value": Bc / per, "unit": "samples/s", "cores": cores, "kind": "port",
            "sampling": {"value": Bc / per_s, "unit": "actions/s"},
            "sample": f"{iters} train steps and {iters_s} reverse passes at batch {Bc} (oracle/torch_tarflow.py on the host, "
                      f"{cores} threads; its reverse recomputes the full causal pass per dimension, there is no K/V cache)"}


def run_ours(args):
    import torch
    import torch.distributed as dist
    import nf_b200                                        # noqa: F401
    from nf_b200 import _lib

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback for the product path)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)   # > 126 MB L2
    peaks, traffic = load_peaks(), load_traffic()
    K, W = args.steps, args.warmup
    passes = 3 if args.precision in ("fp32", "tf32x3") else 1

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()

    # ------------------------------------------------------------------ headline: C2a, exactly K timed steps
    r = Runner(HEADLINE, args.precision, dev, rank, world, flush)
    parity = {HEADLINE: r.parity_check()}
    _lib.lib.nf_launch_count(1)
    gstats = [_lib.graph_stats()]
    ms_train = r.timed("train", r.train_step, K, W)
    launches = _lib.lib.nf_launch_count(1) // (K + W)
    gstats.append(_lib.graph_stats())
    ms_sample = r.timed("sample", r.sample_step, K, W)
    gstats.append(_lib.graph_stats())
    ms_e2e = r.timed("e2e_train", r.train_step, K, W, e2e=True)
    gstats.append(_lib.graph_stats())
    ms_e2e_sample = r.timed("e2e_sample", r.sample_step, K, W, e2e=True)
    gstats.append(_lib.graph_stats())
    ms_e2e_nll = r.timed("e2e_nll", r.train_step_nll, K, W, e2e=True)   # informational, not the headline
    leg_graphs = {name: dict(zip(("replays", "captures", "eager_calls"), (b[i] - a[i] for i in range(3))))
                  for name, a, b in zip(("train", "sample", "e2e_train", "e2e_sample"), gstats[:-1], gstats[1:])}
    # The timed legs above are short (tens of ms at the default K): keep the same kernels running until nvidia-smi
    # (100 ms period) has seen them under load; these extra steps are not part of any reported time.  A FIXED count
    # derived from the max-reduced step time: with world > 1 every train_step contains the gradient all-reduce, so
    # all ranks must run the same number of steps.
    for _ in range(int(min(4000, max(10, 1500.0 / max(ms_train, 1e-3))))):
        r.train_step(r.x, r.y)
    torch.cuda.synchronize()
    head_roof = r.roofline(ms_train, peaks, traffic, True)
    head_stats = dict(r.step_stats)
    B_head = r.B
    bytes_in = r.x_pin.numel() * 4 + r.y_pin.numel() * 4
    r.free()

    # ------------------------------------------------------------------ every BASELINE configuration
    per_config = {}
    names = [c for c in args.configs.split(",") if c]
    for name in names:
        plan = PLAN[name]
        steps = K if plan["steps"] == "K" else max(K, int(plan["steps"]))
        r = Runner(name, args.precision, dev, rank, world, flush)
        parity[name] = r.parity_check()
        entry = {"config": config_dict(name, world, args.precision), "scaling": "weak" if plan["shard"] == "per_gpu" else "strong",
                 "steps": steps, "warmup": W}
        total_B = r.B * world
        if plan["train"]:
            ms = r.timed("train", r.train_step, steps, W)
            entry["train"] = {"value": total_B / (ms * 1e-3), "unit": "samples/s", "ms_per_step": ms,
                              "ms_median_max": list(r.step_stats["train"])}
        ms_s = r.timed("sample", r.sample_step, steps, W)
        entry["sampling"] = {"value": total_B / (ms_s * 1e-3), "unit": "actions/s", "ms_per_step": ms_s,
                             "ms_median_max": list(r.step_stats["sample"])}
        e2e = {"h2d_bytes_per_step": r.x_pin.numel() * 4 + r.y_pin.numel() * 4, "d2h_bytes_per_step": 4}
        if plan["train"]:
            ms_e = r.timed("e2e_train", r.train_step, steps, W, e2e=True)
            e2e["train"] = {"value": total_B / (ms_e * 1e-3), "unit": "samples/s", "ms_per_step": ms_e}
        ms_es = r.timed("e2e_sample", r.sample_step, steps, W, e2e=True)
        e2e["sampling"] = {"value": total_B / (ms_es * 1e-3), "unit": "actions/s", "ms_per_step": ms_es}
        entry["e2e"] = e2e
        entry["roofline"] = r.roofline(ms if plan["train"] else ms_s, peaks, traffic, plan["train"])
        per_config[name] = entry
        r.free()
        if name in args.reduced.split(","):
            # the optional reduced-precision mode (one raw TF32 tensor-core pass; 2e-2 contract for D <= 64, n <= 12):
            # a second dtype of the same workload, reported beside the fp32 numbers
            r = Runner(name, "tf32", dev, rank, world, flush)
            red = {"dtype": "tf32 (one kind::tf32 pass on the raw fp32 operands, fp32 accumulate)",
                   "error_vs_float64": r.parity_check(gate=False)}
            if plan["train"]:
                ms_r = r.timed("train", r.train_step, K, W)
                red["train"] = {"value": total_B / (ms_r * 1e-3), "unit": "samples/s", "ms_per_step": ms_r}
            ms_rs = r.timed("sample", r.sample_step, K, W)
            red["sampling"] = {"value": total_B / (ms_rs * 1e-3), "unit": "actions/s", "ms_per_step": ms_rs}
            F = r.roofline(ms_r if plan["train"] else ms_rs, peaks, traffic, plan["train"])
            red["roofline"] = {k: F[k] for k in ("kernel", "achieved", "frac", "share_of_step", "whole_step")}
            entry["tf32"] = red
            r.free()

    if args.tarflow:
        r = TarflowRunner(dev, rank, world, flush)
        parity["C2t"] = r.parity_check()
        steps = max(K, 50)
        entry = {"config": r.config(), "scaling": "weak", "steps": steps, "warmup": W}
        total_B = r.B * world
        _lib.lib.nf_launch_count(1)
        ms = r.timed("train", r.train_step, steps, W)
        entry["gpu_launches_per_step"] = int(_lib.lib.nf_launch_count(1) // (steps + W))
        entry["train"] = {"value": total_B / (ms * 1e-3), "unit": "samples/s", "ms_per_step": ms,
                          "ms_median_max": list(r.step_stats["train"])}
        ms_s = r.timed("sample", r.sample_step, steps, W)
        entry["sampling"] = {"value": total_B / (ms_s * 1e-3), "unit": "actions/s", "ms_per_step": ms_s,
                             "ms_median_max": list(r.step_stats["sample"])}
        ms_e = r.timed("e2e_train", r.train_step, steps, W, e2e=True)
        ms_es = r.timed("e2e_sample", r.sample_step, steps, W, e2e=True)
        entry["e2e"] = {"h2d_bytes_per_step": r.x_pin.numel() * 4 + r.y_pin.numel() * 4, "d2h_bytes_per_step": 4,
                        "train": {"value": total_B / (ms_e * 1e-3), "unit": "samples/s", "ms_per_step": ms_e},
                        "sampling": {"value": total_B / (ms_es * 1e-3), "unit": "actions/s", "ms_per_step": ms_es}}
        entry["roofline"] = r.roofline(ms, peaks, traffic, True)
        # the evaluation loop's batch (20 parallel environments, utils/torch_evaluations.py:79-82): the autoregressive inverse
        # runs as one persistent cooperative kernel per block there (csrc/tarflow_ar.cu)
        xe, ye = r.x[:20].contiguous(), r.y[:20].contiguous()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(20)]
        for _ in range(W):
            r.sample_step(xe, ye)
        for a, b in ev:
            flush.zero_()
            a.record()
            r.sample_step(xe, ye)
            b.record()
        torch.cuda.synchronize()
        ms_eval = sum(a.elapsed_time(b) for a, b in ev) / len(ev)
        entry["sampling_eval_batch"] = {"batch": 20, "ms_per_call": ms_eval, "value": 20 / (ms_eval * 1e-3), "unit": "actions/s",
                                        "note": "per rank, not aggregated; persistent autoregressive-inverse kernel"}
        per_config["C2t"] = entry
        r.free()

    clocks = sampler.stop() if rank == 0 else None
    replays, captures, eager = _lib.graph_stats()
    p2p_status = _lib.lib.nf_dp_p2p_status() if world > 1 else 0     # 1: peer-memory all-reduce in use, 0: NCCL only
    if p2p_status < 0:
        raise SystemExit("bench.py: the peer-memory all-reduce timed out waiting for a rank (results are invalid)")

    if rank != 0:
        # wait for rank 0 (CPU baseline, JSON line) and leave the process group together with it
        dist.barrier()
        dist.destroy_process_group()
        return

    # ---- CPU baseline: the reference's op sequence in stock PyTorch on this box's host cores, bounded samples,
    # at N = 1 only (the other ranks of a multi-GPU run would just wait for it) -----------------------------------
    cpu_baseline = None
    if world == 1:
        cores = os.cpu_count()
        Bc = min(B_head, 4096)
        ctrain, csample = cpu_step_fn(HEADLINE, Bc)
        per, iters = time_cpu(ctrain, 10.0, 200)
        cpu_baseline = {"value": Bc / per, "unit": "samples/s", "cores": cores, "kind": cpu_kind(HEADLINE),
                        "sample": f"{iters} steps of forward+backward at batch {Bc} of the same config on the host, "
                                  f"{cores} threads, {per * iters:.1f} s ({cpu_what(HEADLINE)})"}
        for name in names:
            if name == HEADLINE:
                per_config[name]["cpu_baseline"] = dict(cpu_baseline)
                continue
            Bc = min(batch_per_gpu(name, 1), CPU_SAMPLE_ROWS[name])
            ctrain, csample = cpu_step_fn(name, Bc)
            fn, unit = (ctrain, "samples/s") if PLAN[name]["train"] else (csample, "actions/s")
            per, iters = time_cpu(fn, 2.5, 20)
            per_config[name]["cpu_baseline"] = {
                "value": Bc / per, "unit": unit, "cores": cores, "kind": cpu_kind(name),
                "sample": f"{iters} {'train steps' if PLAN[name]['train'] else 'reverse passes'} at batch {Bc} "
                          f"on the host, {cores} threads, {per * iters:.1f} s ({cpu_what(name)})"}

        if "C2t" in per_config:
            per_config["C2t"]["cpu_baseline"] = tarflow_cpu_baseline()

    total_B = B_head * world
    head_roof["note"] = (f"algorithmic FLOPs (2MNK per launch, split-precision passes NOT counted) / mean launch duration "
                         f"over {head_roof['launches_timed']} launches; the kernel issues {passes} kind::tf32 MMAs per "
                         f"product at half the bf16 rate, so the tensor-pipe-equivalent rate is "
                         f"{head_roof['achieved'] * passes * 2:.0f} TFLOP/s bf16-equivalent")
    out = {
        "metric": METRIC, "value": total_B / (ms_train * 1e-3), "unit": "samples/s", "n_gpus": world, "steps": K,
        "warmup": W, "ms_per_step": ms_train, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32 (tcgen05 3xTF32 split, fp32 accumulate)" if passes == 3 else args.precision,
        "data": "synthetic",
        "config": config_dict(HEADLINE, world, args.precision),
        "sampling": {"value": total_B / (ms_sample * 1e-3), "unit": "actions/s", "ms_per_step": ms_sample},
        "e2e": {"value": total_B / (ms_e2e * 1e-3), "unit": "samples/s", "h2d_bytes_per_step": bytes_in,
                "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e, "ms_median_max": list(head_stats["e2e_train"]),
                "sampling": {"value": total_B / (ms_e2e_sample * 1e-3), "unit": "actions/s",
                             "d2h_bytes_per_step": 4, "ms_median_max": list(head_stats["e2e_sample"])},
                "with_fused_loss": {"value": total_B / (ms_e2e_nll * 1e-3), "unit": "samples/s",
                                    "ms_per_step": ms_e2e_nll,
                                    "note": "same end-to-end step with loss = model.nll(x, y) (library NLL kernels) "
                                            "instead of the reference's torch-side loss expression"}},
        "gpu_launches": int(launches * K),
        "gpu_launches_per_step": int(launches),
        "launch_graphs": {"replays": replays, "captures": captures, "eager_calls": eager, "per_leg": leg_graphs,
                          "note": "the library replays its kernel sequence as a cached CUDA graph; gpu_launches "
                                  "counts the library's kernels inside the replayed graphs"},
        "allreduce": ("none (one GPU)" if world == 1 else
                      "peer-memory kernel over NVLink for arenas <= %d MB (csrc/dp.cu), NCCL above" % (nf_b200.module._P2P_MAX_BYTES >> 20)
                      if p2p_status == 1 else "NCCL"),
        "train_ms_median_max": list(head_stats["train"]),
        "clocks": clocks,
        "roofline": head_roof,
        "cpu_baseline": cpu_baseline,
        "parity_check": {"checker": "oracle/torch_flow.py in float64 on the GPU, first rows of each config's inputs; "
                                    "err = max|a-ref|/max|ref| <= max(1e-5, 2 err(ref32, ref64))", **parity},
        "per_config": per_config,
    }
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--configs", default="C1a,C1c,C2a,C3,C4,C5",
                    help="comma-separated configurations of the per_config table (empty: headline only)")
    ap.add_argument("--precision", default="fp32", choices=["fp32", "fp32_simt"])
    ap.add_argument("--reduced", default="C4,C5",
                    help="configurations additionally timed in the optional reduced-precision mode (precision='tf32'), "
                         "reported under per_config[...]['tf32'] -- never the headline (the reference is fp32)")
    ap.add_argument("--tarflow", type=int, default=1,
                    help="1: also time the causal-transformer flow (per_config['C2t'], SURVEY section 8 row f1)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
