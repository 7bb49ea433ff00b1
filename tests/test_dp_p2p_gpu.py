"""Peer-memory gradient all-reduce (csrc/dp.cu, nf_dp_p2p_*) on a box with at least two GPUs: the 2-rank run of
tools/dp_p2p_check.py compares it with NCCL on the same buffers (ragged sizes, offsets, sum and average, identical bits
on every rank, replay from a captured graph) and tools/dp_example.py checks the data-parallel training step that uses it
(shard gradients averaged == full-batch gradients).  Skipped on single-GPU boxes; the CPU suite covers the host logic
with gloo (tests/test_dp_gloo.py)."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2,
                                 reason="needs two GPUs of one box")]


def _torchrun(script, port):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tools", script)]
    env = dict(os.environ, NF_B200_DP_P2P_MIN_BYTES="0")     # small arenas through the peer-memory kernel too
    return subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=300, env=env)


def test_peer_memory_allreduce_matches_nccl():
    r = _torchrun("dp_p2p_check.py", 29531)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "peer-memory path status after connect: 1" in r.stdout
    assert "graph replay: 5 replays correct" in r.stdout


def test_data_parallel_step_over_peer_memory():
    r = _torchrun("dp_example.py", 29532)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
