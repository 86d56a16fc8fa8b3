"""SURVEY 8f-3, second half: the binary grid sub-map of simulator/env.py:630-708 ([map] use_grid_map = true).

Golden vectors: tests/golden/local_map_grid.npz, written by tests/golden/make_local_map_grid_golden.py from the
UNMODIFIED reference (EntityBasedCollisionAvoidance.get_local_map, which rotates with the container's cv2 4.13) for
scenes with 3 / 6 / 8 walls, windows of 50, 60 and 31 cells, at fp32-representable robot poses: interior, clipped at
every border of the map, missing the map, axis-aligned headings.  The output is binary and the arithmetic behind it is
integer / exactly representable (fixed-point sample positions, weights that are multiples of 1/1024), so the bar is
EXACT equality of every cell -- oracle vs reference, kernel vs reference, kernel vs oracle.  The one inexact step is
cos / sin of the heading (glibc in the reference and the oracle, CUDA on the device: <= 2 ulp apart), which can move a
sample by 1/32 pixel only when a product lands within an ulp of a rounding tie; the large random batch therefore
states a tolerance of 1e-6 of its cells against the oracle (expected and observed: 0).
"""
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
from ebc.scene import rects_from_zero_cells  # noqa: E402

Z = np.load(os.path.join(ob.GOLDEN, "local_map_grid.npz"))
CASES = list(range(int(Z["n_cases"][0])))


def _sim(N, rects_per_episode, map_size_m, res, device, backend):
    R = max(max(len(r) for r in rects_per_episode), 1)
    cfg = SimConfig()
    cfg.map_size_m, cfg.map_resolution = float(map_size_m), float(res)
    sim = BatchedSim(cfg, N, 1, 0, R, 1, device=device, backend=backend)
    rect = torch.zeros(N, R, 4, dtype=torch.int16)
    cnt = torch.zeros(N, dtype=torch.int32)
    for e, r in enumerate(rects_per_episode):
        if len(r):
            rect[e, :len(r)] = torch.as_tensor(np.asarray(r, np.int16))
        cnt[e] = len(r)
    sim.rect.copy_(rect)
    sim.rect_count.copy_(cnt)
    return sim


def _run(case, device, backend):
    par, grid, poses = Z["c%d_params" % case], Z["c%d_map" % case], Z["c%d_poses" % case]
    rects = rects_from_zero_cells(grid == 0)
    sim = _sim(len(poses), [rects] * len(poses), par[1], par[2], device, backend)
    p32 = torch.tensor(poses, dtype=torch.float32)
    sim.rob_pv[:, 0], sim.rob_pv[:, 1], sim.rob_theta[:] = p32[:, 0], p32[:, 1], p32[:, 2]
    return sim.local_map_grid(float(par[0])).cpu().numpy()


@pytest.mark.parametrize("case", CASES)
def test_oracle_grid_map_matches_reference(case):
    got = _run(case, "cpu", ob.OracleBackend())
    gold = Z["c%d_grid" % case]
    assert got.shape == gold.shape and got.dtype == np.uint8
    assert np.array_equal(got, gold), "%d cells differ" % int((got != gold).sum())


def test_oracle_grid_map_rejects_bad_sizes():
    sim = _sim(1, [[]], 9.0, 0.1, "cpu", ob.OracleBackend())
    out = torch.zeros(1, 10, 10, dtype=torch.uint8)
    m = abi.EbcGridMap(5.0, 10, 0)       # size is not round(5 / 0.1)
    import ctypes
    with pytest.raises(abi.EbcError):
        sim.be.call("local_map_grid", sim.h, ctypes.byref(m), ctypes.c_void_p(out.data_ptr()))
    with pytest.raises(abi.EbcError):   # a window larger than the map (the reference raises a shape error there)
        sim.local_map_grid(9.5)


def _random_batch(N, seed):
    rng = np.random.default_rng(seed)
    rects = []
    for e in range(N):
        k = int(rng.integers(0, 7))
        r = []
        for _ in range(k):
            x0, y0 = rng.integers(-5, 95, 2)
            w, h = rng.integers(1, 40, 2)
            r.append((max(x0, 0), max(y0, 0), min(x0 + w, 90), min(y0 + h, 90)))
        rects.append([q for q in r if q[2] > q[0] and q[3] > q[1]])
    poses = np.stack([rng.uniform(-6.5, 6.5, N), rng.uniform(-6.5, 6.5, N), rng.uniform(-4, 4, N)], 1)
    poses[::7, 2] = np.array([0.0, np.pi / 2, np.pi, -np.pi / 2])[np.arange(len(poses[::7])) % 4]
    return rects, poses


def _run_batch(rects, poses, submap, device, backend):
    sim = _sim(len(poses), rects, 9.0, 0.1, device, backend)
    p32 = torch.tensor(poses, dtype=torch.float32)
    sim.rob_pv[:, 0], sim.rob_pv[:, 1], sim.rob_theta[:] = p32[:, 0], p32[:, 1], p32[:, 2]
    return sim.local_map_grid(submap).cpu().numpy()


def _window(grid, px, py, S, map_size_m=9.0, res=0.1):
    """env.py:637-684 with the reference's own index arithmetic: the clipped window of scene.map, unrotated."""
    G = grid.shape[0]
    cx, cy = int(round((px + map_size_m / 2.0) / res)), int(round((py + map_size_m / 2.0) / res))
    win = np.ones((S, S), np.uint8)
    six, siy = cx - S // 2, cy - S // 2
    eix, eiy = six + S - 1, siy + S - 1
    sgx = sgy = 0
    egx = egy = S - 1
    if six < 0:
        sgx, six = -six, 0
    elif eix > G - 1:
        egx, eix = egx - (eix - (G - 1)), G - 1
    if siy < 0:
        sgy, siy = -siy, 0
    elif eiy > G - 1:
        egy, eiy = egy - (eiy - (G - 1)), G - 1
    if not (sgy > egy or siy > eiy or six > eix or sgx > egx):
        win[sgx:egx, sgy:egy] = grid[six:eix, siy:eiy]
    return win


def test_oracle_grid_map_is_a_rotation_of_the_window():
    """Size-independent property: with the heading along +y the rotation angle is 0 and the map is the clipped
    window itself (last row / column left at 1, env.py:682-684)."""
    rects, poses = _random_batch(64, 3)
    poses[:, 2] = np.float32(np.pi / 2)
    # (-theta32 + pi/2) is ~4e-8 rad, not 0: every sample stays within 1/32 pixel of its cell centre -> identity
    got = _run_batch(rects, poses, 5.0, "cpu", ob.OracleBackend())
    S, G = 50, 90
    for e in range(64):
        grid = np.ones((G, G), np.uint8)
        for (x0, y0, x1, y1) in rects[e]:
            grid[x0:x1, y0:y1] = 0
        win = _window(grid, float(np.float32(poses[e, 0])), float(np.float32(poses[e, 1])), S)
        assert np.array_equal(got[e], win), e


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_gpu_grid_map_matches_reference_and_oracle(case):
    got = _run(case, "cuda:0", None)
    assert np.array_equal(got, Z["c%d_grid" % case])
    assert np.array_equal(got, _run(case, "cpu", ob.OracleBackend()))


@pytest.mark.gpu
@pytest.mark.parametrize("submap", [5.0, 6.0, 3.1, 7.0])      # 7 m = 70 cells: the general (byte-grid) kernel; the others fit 64-bit rows
def test_gpu_grid_map_large_batch(submap):
    """4096 episodes with 0..6 rectangles each and random poses (a third of them with the window off the map's
    edge): kernel against the oracle, and an episode's map does not depend on its neighbours (permutation)."""
    rects, poses = _random_batch(4096, 11)
    got = _run_batch(rects, poses, submap, "cuda:0", None)
    ref = _run_batch(rects, poses, submap, "cpu", ob.OracleBackend())
    assert (got != ref).sum() <= 1e-6 * got.size, "%d cells differ" % int((got != ref).sum())
    assert 0.02 < (got == 0).mean() < 0.9
    perm = np.random.default_rng(1).permutation(4096)
    got_p = _run_batch([rects[i] for i in perm], poses[perm], submap, "cuda:0", None)
    assert np.array_equal(got_p, got[perm])
    # both kernels (64-bit rows + lookup table for windows of <= 62 cells, byte grid + integer taps otherwise) agree
    os.environ["EBC_GRID_MAP_BYTES"] = "1"
    try:
        assert np.array_equal(_run_batch(rects, poses, submap, "cuda:0", None), got)
    finally:
        del os.environ["EBC_GRID_MAP_BYTES"]


@pytest.mark.gpu
def test_gpu_env_api_returns_the_grid_map(tmp_path):
    """The reference's call: [map] use_grid_map = true makes env.reset / env.step return get_local_map(robot state)
    (simulator/env.py:195-204, :460-466).  Same config and test case as golden case 0: the replayed scene has the
    reference's scene.map and get_local_map(ob) gives the reference's arrays."""
    import configparser
    from simulator.utils.state import FullState
    from simulator.utils.test_utils import configure_env_policy_robot
    cfg_dir = os.path.join(ob.GOLDEN, "configs")
    cp = configparser.RawConfigParser()
    cp.read(os.path.join(cfg_dir, "env_ebcadrl.config"))
    cp.set("map", "use_grid_map", "true")
    cp.set("map", "submap_size_m", "5")
    path = str(tmp_path / "env_grid.config")
    with open(path, "w") as f:
        cp.write(f)
    env, policy, robot = configure_env_policy_robot(path, os.path.join(cfg_dir, "policy.config"), policy="linear")
    ob_, local_map = env.reset("test", test_case=3)
    assert np.array_equal((np.asarray(env.scene.map) > 0).astype(np.uint8), Z["c0_map"])
    assert local_map.shape == (50, 50) and local_map.dtype == np.float64 and set(np.unique(local_map)) <= {0.0, 1.0}
    assert len(env.local_maps) == 1
    for k, (px, py, theta) in enumerate(Z["c0_poses"]):
        got = env.get_local_map(FullState(px, py, 0.0, 0.0, 0.3, 0.0, 0.0, 1.0, theta), append=False)
        assert np.array_equal(got, Z["c0_grid"][k].astype(np.float64)), k
    action = robot.act(ob_, local_map=local_map, env=env)
    ob_, local_map2, reward, done, info = env.step(action)
    assert local_map2.shape == (50, 50) and len(env.local_maps) == 2
    # the heading of the holonomic robot stays pi/2: the map is the (clipped) window of scene.map itself
    win = _window((np.asarray(env.scene.map) > 0).astype(np.uint8), float(robot.px), float(robot.py), 50)
    assert np.array_equal(local_map2, win.astype(np.float64))
