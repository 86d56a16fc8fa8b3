"""the angular local map of simulator/env.py:468-628.

Golden vectors: tests/golden/local_map_angular.npz, written by tests/golden/make_local_map_golden.py from the UNMODIFIED
reference (EntityBasedCollisionAvoidance.get_local_map_angular) for scenes with walls, circular obstacles, a forward fan
of 72 sectors and no obstacles, at fp32-representable robot poses.  Stated tolerance: 1e-12 absolute on distances of at
most a few metres (the arithmetic is the reference's float64 sequence; numpy's / glibc's / CUDA's cos, sin, atan2 differ
in the last place) -- and every sector must be touched or untouched exactly as in the reference.
"""
import ctypes
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "eb-cadrl_b200"))
sys.path.insert(0, HERE)

import oracle_backend as ob  # noqa: E402
from ebc import abi  # noqa: E402
from ebc.config import SimConfig  # noqa: E402
from ebc.engine import BatchedSim  # noqa: E402

TOL = 1e-12
Z = np.load(os.path.join(ob.GOLDEN, "local_map_angular.npz"))
CASES = list(range(int(Z["n_cases"][0])))


def _run(case, device, backend, normalize):
    par, verts, poses = Z["c%d_params" % case], Z["c%d_verts" % case], Z["c%d_poses" % case]
    N, P, dim = len(poses), max(len(verts), 1), int(par[3])
    sim = BatchedSim(SimConfig(), N, 1, 0, 0, 1, device=device, backend=backend)
    p32 = torch.tensor(poses, dtype=torch.float32)
    sim.rob_pv[:, 0], sim.rob_pv[:, 1], sim.rob_gr[:, 3], sim.rob_theta[:] = p32[:, 0], p32[:, 1], p32[:, 2], p32[:, 3]
    xy = torch.zeros(N, P, 4, 2, dtype=torch.float64)
    if len(verts):
        xy[:, :len(verts)] = torch.as_tensor(verts)
    cnt = torch.full((N,), len(verts), dtype=torch.int32)
    out = sim.local_map_angular(xy.to(device), cnt.to(device), par[0], par[1], par[2], dim, normalize=normalize)
    return out.cpu().numpy(), par


def _check(got, case, normalize, par):
    gold = Z["c%d_%s" % (case, "norm" if normalize else "raw")]
    full = 1.0 if normalize else par[0]
    assert np.array_equal(got == full, gold == full), "a sector is touched in one map and not in the other"
    assert np.abs(got - gold).max() <= TOL


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("normalize", [True, False])
def test_oracle_angular_map_matches_reference(case, normalize):
    got, par = _run(case, "cpu", ob.OracleBackend(), normalize)
    _check(got, case, normalize, par)


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("normalize", [True, False])
def test_gpu_angular_map_matches_reference_and_oracle(case, normalize):
    got, par = _run(case, "cuda:0", None, normalize)
    _check(got, case, normalize, par)
    ref, _ = _run(case, "cpu", ob.OracleBackend(), normalize)
    assert np.abs(got - ref).max() <= TOL


@pytest.mark.gpu
def test_gpu_angular_map_large_batch_and_ragged_counts():
    """4096 episodes with 0..6 quadrilaterals each (more sweeps than lanes for the largest), random poses: kernel
    against the oracle, and an episode's map does not depend on its neighbours (permutation)."""
    rng = np.random.default_rng(5)
    N, P, dim = 4096, 6, 48
    verts = np.zeros((N, P, 4, 2))
    c = rng.uniform(-4, 4, (N, P, 2)); w = rng.uniform(0.1, 1.5, (N, P, 2))
    for k, (sx, sy) in enumerate(((1, 1), (-1, 1), (-1, -1), (1, -1))):
        verts[:, :, k, 0] = c[..., 0] + sx * w[..., 0]
        verts[:, :, k, 1] = c[..., 1] + sy * w[..., 1]
    cnt = rng.integers(0, P + 1, N).astype(np.int32)
    pose = np.stack([rng.uniform(-4.5, 4.5, N), rng.uniform(-4.5, 4.5, N), rng.uniform(0.2, 0.5, N),
                     rng.uniform(-np.pi, np.pi, N)], 1).astype(np.float32)
    outs = {}
    for name, dev, be, perm in (("gpu", "cuda:0", None, None), ("ref", "cpu", ob.OracleBackend(), None),
                                ("gpu_perm", "cuda:0", None, rng.permutation(N))):
        idx = np.arange(N) if perm is None else perm
        sim = BatchedSim(SimConfig(), N, 1, 0, 0, 1, device=dev, backend=be)
        p32 = torch.tensor(pose[idx])
        sim.rob_pv[:, 0], sim.rob_pv[:, 1], sim.rob_gr[:, 3], sim.rob_theta[:] = p32[:, 0], p32[:, 1], p32[:, 2], p32[:, 3]
        out = sim.local_map_angular(torch.tensor(verts[idx]).to(dev), torch.tensor(cnt[idx]).to(dev), 3.0, -np.pi, np.pi, dim)
        outs[name] = out.cpu().numpy()
        if perm is not None:
            inv = np.empty(N, np.int64); inv[perm] = np.arange(N)
            outs[name] = outs[name][inv]
    assert np.array_equal(outs["gpu"], outs["gpu_perm"])
    assert np.array_equal(outs["gpu"] == 1.0, outs["ref"] == 1.0)
    assert np.abs(outs["gpu"] - outs["ref"]).max() <= TOL
    assert (outs["gpu"][cnt == 0] == 1.0).all()


@pytest.mark.gpu
def test_env_mirror_returns_the_angular_map():
    """reset / step of the host mirror return the map like the reference (env.py:194-204, :460-466)."""
    import configparser
    from simulator.agents.robot import Robot
    from simulator.env import EntityBasedCollisionAvoidance
    from simulator.policy.policy_factory import policy_factory
    from simulator.utils.action import ActionXY
    from simulator.utils.state import FullState
    cp = configparser.RawConfigParser()
    cp.read(os.path.join(ob.GOLDEN, "configs", "env_ebcadrl.config"))
    env = EntityBasedCollisionAvoidance()
    env.configure(cp)
    robot = Robot(cp, "robot")
    robot.set_policy(policy_factory["linear"]())
    env.set_robot(robot)
    ob_, local_map = env.reset(phase="test", test_case=3)
    assert local_map.shape == (48,) and local_map.dtype == np.float64 and len(env.local_maps_angular) == 1
    _, lm2, _, _, _ = env.step(ActionXY(0.3, 0.4))
    assert lm2.shape == (48,) and len(env.local_maps_angular) == 2
    # the scene of golden case c0 is this scene (same config, test_case 3): an explicit pose reproduces its vectors
    verts = np.asarray(env.scene.obstacle_vertices, dtype=np.float64).reshape(-1, 4, 2)
    assert np.array_equal(verts, Z["c0_verts"])
    for pose, gold in zip(Z["c0_poses"][:6], Z["c0_norm"][:6]):
        got = env.get_local_map_angular(FullState(pose[0], pose[1], 0, 0, pose[2], 0, 0, 1, pose[3]), append=False)
        assert np.abs(got - gold).max() <= TOL
    env.local_map_enabled = False
    assert env.step(ActionXY(0.0, 0.0))[1] is None


@pytest.mark.gpu
def test_batched_env_local_map_batch():
    """BatchedEnv.local_map_batch: the reset-time `local_map` of N reference scenes in one launch.  Episodes 0-3 are
    the scenes of golden cases c0-c3 (test_case 3..6 = scene_number 1003..1006); the robot stands at its start pose
    (0, -circle_radius, theta = pi / 2), which the oracle evaluates from the same fp32 state."""
    from ebc.batched_env import BatchedEnv
    env = BatchedEnv(os.path.join(ob.GOLDEN, "configs", "env_ebcadrl.config"), None, 4, "cuda:0")
    env.reset_batch("test", [1003, 1004, 1005, 1006])
    for e in range(4):
        assert np.array_equal(env.poly_xy[e, :3].cpu().numpy(), Z["c%d_verts" % e])
    got = env.local_map_batch().cpu().numpy()
    assert got.shape == (4, 48)
    ref_sim = BatchedSim(SimConfig(), 4, 1, 0, 0, 1, device="cpu", backend=ob.OracleBackend())
    ref_sim.rob_pv.copy_(env.sim.rob_pv.cpu()); ref_sim.rob_gr.copy_(env.sim.rob_gr.cpu()); ref_sim.rob_theta.copy_(env.sim.rob_theta.cpu())
    ref = ref_sim.local_map_angular(env.poly_xy.cpu(), env.poly_count.cpu(), 3.0, -np.pi, np.pi, 48).numpy()
    assert np.abs(got - ref).max() <= TOL and (got < 1.0).any()
