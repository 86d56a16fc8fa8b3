"""bench.py's bookkeeping, checked on the CPU: the roofline numerator is SURVEY 8d's algorithmic FLOP count, the
workloads are BASELINE.json's configurations, and the reference arm prints the contract's keys (it runs the oracle
port on the host cores; no GPU involved)."""
import importlib.util
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_algorithmic_flops_are_the_surveys():
    """F = 2 (n M_e + M_s); EB-CADRL net 245,300 MAC per entity + 132,000 per state -> 8.11 MFLOP at n = 16,
    14.98 at n = 30; SARL baseline 62,050 + 33,500 -> 0.6875 MFLOP at n = 5 (a17)."""
    b = _bench()
    eb, _ = b.value_net_weights(fixture="weights_ebcadrl.npz")
    sarl, _ = b.value_net_weights(fixture="weights_sarl_baseline.npz")
    assert b.flops_per_eval(eb, 16) == 2 * (16 * 245300 + 132000) == 8113600
    assert b.flops_per_eval(eb, 30) == 2 * (30 * 245300 + 132000)
    assert b.flops_per_eval(sarl, 5) == 2 * (5 * 62050 + 33500) == 687500


def test_workloads_are_the_baseline_configs():
    b = _bench()
    shape, cfg = b.workload("cfg2")                       # BASELINE configs[1]: 10 typed humans + 3 walls, 4096 episodes
    assert (shape.H, shape.Smax, b.WORKLOADS["cfg2"][4], cfg.with_agent_type, cfg.robot_kinematics) == (10, 6, 4096, True, "holonomic")
    shape, cfg = b.workload("cfg3")                       # configs[2]: unicycle robot, 16,384 episodes
    assert (shape.H, shape.Smax, b.WORKLOADS["cfg3"][4], cfg.robot_kinematics) == (5, 0, 16384, "unicycle")
    shape, cfg = b.workload("cfg4")                       # configs[3]: 20 humans + 10 walls, 8,192 episodes per GPU
    assert (shape.H, shape.num_walls, b.WORKLOADS["cfg4"][4]) == (20, 10, 8192)
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert len(base["configs"]) == 5 and "4096" in base["configs"][1] and "16384" in base["configs"][2]
    peaks = b.measured_peaks()
    assert peaks["hbm_gbs"] > 0 and peaks["tensor_tflops"] > 0


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, env=dict(os.environ, RANK="0", WORLD_SIZE="1"))
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert line["impl"] == "reference" and line["metric"] == "agent_steps_per_sec" and line["higher_is_better"] is True
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] == (os.cpu_count() or 1)
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["config"]["workload"] == "cfg2_h10_typed_3walls" and line["value"] > 0
    # rank > 0 of a torchrun launch exits 0 without work
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=120, env=dict(os.environ, RANK="1", WORLD_SIZE="2"))
    assert out.returncode == 0 and not [l for l in out.stdout.splitlines() if l.startswith("{")]
