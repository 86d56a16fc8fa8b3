"""ORCA obstacle half-planes (spec: RVO2 computeNewVelocity obstacle section + processObstacles,
reference caller simulator/policy/orca_obstacles.py:83-152).  The reference never adds an obstacle on its live path,
so there is no reference golden: the checks are (i) the product's host packer against the oracle's independent
restatement of RVO2's addObstacle / buildObstacleTree, byte for byte, (ii) behavioural properties of the oracle that
RVO2 guarantees (agents never enter a wall, they slide along it at radius + 0.01), (iii) flag off == no obstacles,
and -- on the GPU -- (iv) the CUDA kernels bit-exact against the oracle."""
import ctypes

import numpy as np
import pytest
import torch

import oracle_backend as ob
from ebc import abi, synth
from ebc.actions import build_action_space
from ebc.config import SimConfig
from ebc.engine import BatchedSim
from ebc.scene import pack_obstacles, wall_polygon

WALLS5 = synth.SceneShape("walls5_h12", [(0, 12, (1.0, 1.0), (0.3, 0.3))], rule="mixed", square_width=13.0,
                          circle_radius=6.0, robot_v_pref=1.0, num_walls=5, wall_len=(2, 4), map_size_m=14.0)


def _cfg(shape, flag):
    c = SimConfig()
    c.map_size_m, c.map_resolution, c.time_limit = shape.map_size_m, shape.map_resolution, 100.0
    c.orca_obstacles = flag
    return c


def _both_packers(oracle, polys, cap=64):
    xy = np.asarray([c for p in polys for v in p for c in v], dtype=np.float32)
    sizes = np.asarray([len(p) for p in polys], dtype=np.int32)
    res = []
    for fn in (oracle.lib.ebc_ref_pack_obstacles, abi.load().ebc_pack_obstacles):
        out = np.zeros(64, dtype=abi.OBST_DTYPE)
        n = ctypes.c_int32(-1)
        rc = fn(xy.ctypes.data_as(ctypes.c_void_p), sizes.ctypes.data_as(ctypes.c_void_p), len(polys),
                out.ctypes.data_as(ctypes.c_void_p), cap, ctypes.byref(n))
        res.append((rc, n.value if rc == 0 else -1, out.tobytes() if rc == 0 else b""))
    return res


def test_packers_agree_byte_for_byte(oracle):
    """ebc_pack_obstacles (product, C++) vs ebc_ref_pack_obstacles (oracle, C): same vertices, links, convexity,
    split points and kd-tree masks on random wall layouts, including layouts whose splitting overflows 64 vertices."""
    rng = np.random.default_rng(0)
    splits = overflow = 0
    for _ in range(600):
        R = int(rng.integers(1, 11))
        polys = []
        for _ in range(R):
            length, horiz = int(rng.integers(2, 5)), rng.random() > 0.5
            cx, cy = np.floor(rng.uniform(-70, 70, 2)) * 0.1
            polys.append(wall_polygon(cx, cy, length if horiz else 1, 1 if horiz else length))
        a, b = _both_packers(oracle, polys)
        assert a == b
        splits += a[1] > 4 * R
        if a[1] > 4 * R:                            # no room for the split vertices: both must refuse
            c, d = _both_packers(oracle, polys, cap=4 * R)
            assert c == d and c[0] == abi.ERR_INVALID
            overflow += 1
    assert splits > 100 and overflow > 100         # both regimes were exercised
    # a triangle with a reflex neighbour: convexity flags (leftOf(prev, cur, next) >= 0)
    a, b = _both_packers(oracle, [[(0, 0), (2, 0), (1, 0.2), (1, 2)]])
    assert a == b and a[0] == 0
    rec = np.frombuffer(a[2], dtype=abi.OBST_DTYPE)[:a[1]]
    assert rec["convex"][:4].tolist() == [1, 1, 0, 1]
    # two-vertex obstacle (a line): both ends convex, unit directions opposite
    a, b = _both_packers(oracle, [[(0, 0), (3, 0)]])
    rec = np.frombuffer(a[2], dtype=abi.OBST_DTYPE)[:2]
    assert a == b and rec["convex"].tolist() == [1, 1] and rec["ux"].tolist() == [1.0, -1.0]


def test_kd_tree_masks_describe_one_tree(oracle):
    """anc / anc_left of the packed records: exactly one root, every node's ancestors form a chain."""
    sc = synth.generate(synth.CFG4, np.arange(32), max_obst=64)
    for e in range(32):
        n = int(sc["obst_count"][e])
        rec = sc["obst"][e, :n]
        anc = [int(a) for a in rec["anc"]]
        assert sum(a == 0 for a in anc) == 1
        for i in range(n):
            assert not (anc[i] >> i) & 1 and int(rec["anc_left"][i]) & ~anc[i] == 0
            chain = sorted((bin(anc[z]).count("1"), z) for z in range(n) if (anc[i] >> z) & 1)
            assert [d for d, _ in chain] == list(range(len(chain)))          # depths 0, 1, 2, ... : a path
            for (_, hi), (_, lo) in zip(chain, chain[1:]):
                assert anc[lo] == anc[hi] | (1 << hi)
        assert rec["next"][rec["prev"]].tolist() == list(range(n))           # prev / next are inverse links


def _clearance(sim, rect_m):
    p = sim.hum_pv.numpy()[:, :, :2].astype(np.float64)
    r = sim.hum_gr.numpy()[:, :, 3]
    dx = np.maximum(np.maximum(rect_m[:, None, :, 0] - p[:, :, None, 0], 0), p[:, :, None, 0] - rect_m[:, None, :, 2])
    dy = np.maximum(np.maximum(rect_m[:, None, :, 1] - p[:, :, None, 1], 0), p[:, :, None, 1] - rect_m[:, None, :, 3])
    return (np.hypot(dx, dy) - r[:, :, None]).min(2)


def test_oracle_humans_never_enter_walls(oracle):
    """What RVO2's obstacle lines guarantee: an agent that starts clear of every wall never penetrates one (it may
    touch it at its ORCA radius r + 0.01); with the flag off the same crowd walks straight through."""
    shape, N = WALLS5, 192
    sc = synth.generate(shape, np.arange(N), max_obst=64)
    assert sc["obst_count"].max() <= 64
    rect_m = sc["rect"].astype(np.float64) * shape.map_resolution - shape.map_size_m / 2.0
    worst = {}
    for flag in (True, False):
        sim = BatchedSim(_cfg(shape, flag), N, shape.H, shape.Smax, shape.Rmax, 81, device="cpu", backend=oracle, max_obst=64)
        sim.set_actions(build_action_space(1.0))
        synth.load(sim, sc)
        c0 = _clearance(sim, rect_m)
        clear = c0 > 0.05
        mn = c0.copy()
        zero = torch.zeros(N, dtype=torch.int32)
        for _ in range(100):
            sim.step(action_idx=zero, fused_orca=True)
            mn = np.minimum(mn, _clearance(sim, rect_m))
        assert np.isfinite(sim.hum_pv.numpy()).all()
        worst[flag] = mn[clear].min()
    assert worst[True] > 0.0099 - 1e-4, worst         # slides along the wall at the +0.01 radius pad
    assert worst[False] < -0.25, worst                # reference behaviour: humans ignore walls


def test_flag_off_ignores_bound_obstacles(oracle):
    """orca_obstacles = 0 (the default, the reference's live behaviour): bound obstacle arrays change nothing."""
    shape, N = synth.CFG2, 64
    sc = synth.generate(shape, np.arange(N), max_obst=32)
    plain = {k: v for k, v in sc.items() if not k.startswith("obst")}
    out = []
    for scenes, omax in ((sc, 32), (plain, 0)):
        sim = BatchedSim(_cfg(shape, False), N, shape.H, shape.Smax, shape.Rmax, 81, device="cpu", backend=oracle, max_obst=omax)
        sim.set_actions(build_action_space(shape.robot_v_pref))
        synth.load(sim, scenes)
        for _ in range(5):
            sim.step(action_idx=torch.zeros(N, dtype=torch.int32), fused_orca=True)
        out.append(sim.hum_pv.numpy().copy())
    assert np.array_equal(out[0], out[1])


@pytest.mark.gpu
@pytest.mark.parametrize("shape_name,N,steps", [("CFG2", 2048, 24), ("CFG4", 768, 16), ("WALLS5", 1024, 40)])
def test_gpu_obstacle_orca_is_bit_exact(oracle, shape_name, N, steps):
    """K1 with obstacle half-planes (vertices staged in shared memory, visiting-order ranking from the kd-tree masks,
    cover resolution by ballots, 64-line warp LP with numObstLines honoured in linearProgram3) == the oracle's
    sequential RVO2 restatement, bit for bit, step after step; the ORCA robot too; and flag off == round-1 kernels."""
    shape = WALLS5 if shape_name == "WALLS5" else getattr(synth, shape_name)
    sc = synth.generate(shape, np.arange(N), max_obst=64)
    sims = []
    for dev, be in (("cuda:0", None), ("cpu", oracle)):
        s = BatchedSim(_cfg(shape, True), N, shape.H, shape.Smax, shape.Rmax, 81, device=dev, backend=be, max_obst=64)
        s.set_actions(build_action_space(shape.robot_v_pref))
        synth.load(s, sc)
        sims.append(s)
    g, r = sims
    n_lp3 = 0
    for t in range(steps):
        for s in sims:
            s.orca()
        torch.cuda.synchronize()
        a, b = g.hum_nv.cpu().numpy(), r.hum_nv.numpy()
        same = (a == b) | (np.isnan(a) & np.isnan(b))
        assert same.all(), "step %d: %d of %d ORCA velocities differ, first at %s" % (t, (~same).sum(), same.size, np.argwhere(~same)[0])
        ga, ra = g.robot_orca(0.15), r.robot_orca(0.15)
        torch.cuda.synchronize()
        assert np.array_equal(ga.cpu().numpy(), ra.numpy()), "robot ORCA action, step %d" % t
        idx = torch.as_tensor(np.random.default_rng(t).integers(0, 81, N), dtype=torch.int32)
        # alternate the two launch shapes: fused K1 + K2, and K2 after a separate K1
        if t % 2:
            g.step(action_idx=idx.cuda(), fused_orca=True); r.step(action_idx=idx, fused_orca=True)
        else:
            g.step(action_idx=idx.cuda()); r.step(action_idx=idx)
        torch.cuda.synchronize()
        assert np.array_equal(g.hum_pv.cpu().numpy(), r.hum_pv.numpy()), "human states, step %d" % t
        sp = np.linalg.norm(a, axis=2)
        n_lp3 += int((sp < 1e-3).sum())
    # walls really constrain somebody: the same crowd without the flag moves differently
    off = BatchedSim(_cfg(shape, False), N, shape.H, shape.Smax, shape.Rmax, 81, device="cuda:0", max_obst=64)
    off.set_actions(build_action_space(shape.robot_v_pref))
    synth.load(off, sc)
    base = BatchedSim(_cfg(shape, False), N, shape.H, shape.Smax, shape.Rmax, 81, device="cuda:0")
    base.set_actions(build_action_space(shape.robot_v_pref))
    synth.load(base, {k: v for k, v in sc.items() if not k.startswith("obst")})
    for s in (off, base):
        s.orca()
    torch.cuda.synchronize()
    assert torch.equal(off.hum_nv, base.hum_nv)                       # flag off: bit-identical to no obstacles at all
    synth.load(g, sc)
    g.orca()
    torch.cuda.synchronize()
    assert not torch.equal(g.hum_nv, off.hum_nv)
