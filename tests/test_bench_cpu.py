"""bench.py's reference arm runs without a GPU (the driver launches it on the GPU box's host cores): one short run
here checks the JSON contract of that line."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "1"], capture_output=True, text=True, timeout=600, env=env)
    assert res.returncode == 0, res.stderr[-2000:]
    line = [l for l in res.stdout.splitlines() if l.startswith("{")][-1]
    d = json.loads(line)
    assert d["impl"] == "reference" and d["unit"] == "samples/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 1 and d["n_gpus"] == 1
    # "reference" when __graft_entry__.build() has compiled the unmodified reference module into oracle/_ref/ (the build
    # container, and the GPU box through the snapshot), "port" (oracle/torch_flow.py) otherwise
    staged = os.path.exists(os.path.join(ROOT, "oracle", "_ref", "torch_fcnf.bytecode"))
    assert d["cpu_baseline"]["kind"] == ("reference" if staged else "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "C2a" in d["config"]["workload"]
    # both arms print the SAME config object for the headline (the driver compares them)
    sys.path.insert(0, ROOT)
    import bench
    assert d["config"] == bench.config_dict(bench.HEADLINE, 1, "fp32")
    assert d["config"]["batch_per_gpu"] == 4096 and d["config"]["D"] == 8 and d["config"]["n"] == 6


def test_per_config_plan_matches_baseline_configs():
    """BASELINE.json configs: batch 256 (ant), 4096 (GCBC), 16384 per GPU (latent actor), 65536 envs sharded
    (kitchen sampling), 65536 global (stress)."""
    sys.path.insert(0, ROOT)
    import bench
    assert bench.batch_per_gpu("C1a", 8) == 256 and bench.batch_per_gpu("C2a", 8) == 4096
    assert bench.batch_per_gpu("C3", 4) == 16384
    assert bench.batch_per_gpu("C4", 8) == 8192 and bench.batch_per_gpu("C4", 1) == 65536
    assert bench.batch_per_gpu("C5", 1) == 65536 and bench.batch_per_gpu("C5", 2) == 32768
    assert not bench.PLAN["C4"]["train"] and bench.PLAN["C5"]["train"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="", RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=120, env=env)
    assert res.returncode == 0 and not [l for l in res.stdout.splitlines() if l.startswith("{")]
