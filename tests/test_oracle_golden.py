"""Pin the CPU oracle (oracle/ebc_oracle.c) against vectors produced by the UNMODIFIED
reference (tests/golden/make_golden.py).  CPU only.

Bars (BASELINE.json north_star): event / done flags bit-exact on single-step comparisons from
identical (fp32-representable) states; positions, velocities, rewards within 1e-5; argmax
identical wherever the reference's own top-2 value gap exceeds the fp32 value-network noise.
"""
import json
import os

import numpy as np
import pytest
import torch

import oracle_backend as ob
from ebc.actions import build_action_space
from ebc.config import SimConfig
from ebc.engine import BatchedSim

POS_TOL = 1e-5      # positions / velocities / rewards (north_star)
VIN_TOL = 2e-5      # rotated fp32 rows: atan2f/cosf/sinf of glibc vs torch differ by ulps
VALUE_TOL = 2e-4    # fp32 value network, different summation order (MKL vs sequential)
ARGMAX_GAP = 5e-4   # below this top-2 gap an argmax flip is fp32 noise, not an error


def make_sim(tr, backend, n_actions=81, weights=None):
    H, S, R = tr.dims()
    cfg = tr.sim_config()
    sim = BatchedSim(cfg, 1, H, S, R, n_actions, device="cpu", backend=backend)
    if "actions" in tr.z.files:
        sim.set_actions(tr.z["actions"])
    if weights:
        sim.set_weights(ob.load_weights(weights))
    return sim


def test_collision_known_answers(oracle):
    """tests/test_collisions.py:13-143 of the reference: 6 hand-built swept-disc cases."""
    cases = json.load(open(os.path.join(ob.GOLDEN, "collisions_known_answers.json")))
    assert len(cases) == 6
    for c in cases:
        cfg = SimConfig()
        cfg.time_step = c["time_step"]
        cfg.collision_penalty_adult = -0.25
        sim = BatchedSim(cfg, 1, 1, 0, 0, 1, device="cpu", backend=oracle)
        sim.load_episodes(0, np.array([[[c["adult_pos"][0], c["adult_pos"][1], 0, 0]]], np.float32),
                          np.array([[[0, 0, 1.0, c["adult_radius"]]]], np.float32), np.zeros((1, 1), np.uint8),
                          np.array([1], np.int32), None, np.array([0], np.int32), None, np.array([0], np.int32),
                          np.array([[c["robot_pos"][0], c["robot_pos"][1], 0, 0]], np.float32),
                          np.array([[100.0, 100.0, 1.0, c["robot_radius"]]], np.float32),
                          np.array([np.pi / 2], np.float32), np.array([0.0]))
        sim.hum_nv.zero_()
        act = torch.tensor([c["action"]], dtype=torch.float64)
        sim.step(action=act)
        collided = int(sim.event[0]) == ob.abi.EV_COLLISION_ADULT
        assert collided == c["collision"], c["name"]
        if not c["collision"]:
            # radii 0.9 / 1.2 are not fp32-representable: fp32 state moves dmin by ~1e-8
            assert abs(float(sim.dmin[0, 0]) - c["dmin"]) < 1e-6


def test_action_tables():
    z = np.load(os.path.join(ob.GOLDEN, "action_tables.npz"))
    for key in z.files:
        kin, vp = key.rsplit("_", 1)
        mine = build_action_space(float(vp), kin)
        assert mine.shape == (81, 2)
        assert np.array_equal(mine, z[key]), key     # float64 bit-exact


@pytest.mark.parametrize("name", sorted(ob.TRACE_WEIGHTS))
def test_trace_single_steps(oracle, name):
    tr = ob.Trace(name)
    sim = make_sim(tr, oracle, weights=ob.TRACE_WEIGHTS[name])
    n_flip, n_dec, worst_vin, worst_val, n_last = 0, 0, 0.0, 0.0, 0
    for t in range(tr.n_steps):
        H, S = tr.load_into(sim, t)
        st = tr.steps[t]
        if tr.has(t, "last_state"):
            # multi_human_rl.py:84-85,128-149: the replay sample = transform(CURRENT joint state), pinned by the
            # reference's own policy.last_state (train phase)
            ref = tr.get(t, "last_state")
            mine = sim.transform()[0].numpy()
            assert ref.shape == (H + S, sim.D)
            np.testing.assert_allclose(mine[:H + S], ref, rtol=0, atol=VIN_TOL)
            assert not mine[H + S:].any()
            n_last += 1
        sim.orca()
        # humans' ORCA actions: fp32 inside rvo2 on both sides -> bit-exact; `linear` humans (linear.py:17-23) are
        # float64 in the reference and fp32 state here: equal after rounding to fp32
        nv, ref_nv = sim.hum_nv[0, :H].numpy(), tr.get(t, "orca")
        lin = np.array([tr.cfg_dict["human_policy"][int(ty)] == 1 for ty in tr.get(t, "hum_type")])
        assert np.array_equal(nv[~lin].astype(np.float64), ref_nv[~lin]), (name, t)
        assert np.array_equal(nv[lin], ref_nv[lin].astype(np.float32)), (name, t)
        if tr.has(t, "la_reward"):
            sim.lookahead()
            assert np.array_equal(sim.la_event[0].numpy(), tr.get(t, "la_event")), (name, t)
            assert np.array_equal(sim.la_done[0].numpy(), tr.get(t, "la_done")), (name, t)
            np.testing.assert_allclose(sim.la_reward[0].numpy(), tr.get(t, "la_reward"), rtol=0, atol=POS_TOL)
            if tr.has(t, "vin"):
                ref = tr.get(t, "vin")                     # A x n_e x D
                mine = sim.vin[0, :, :H + S].numpy()
                worst_vin = max(worst_vin, float(np.abs(mine - ref).max()))
                np.testing.assert_allclose(mine, ref, rtol=0, atol=VIN_TOL)
                assert not sim.vin[0, :, H + S:].any()
            sim.value()
            dv = float(np.abs(sim.values[0].numpy() - tr.get(t, "la_value")).max())
            worst_val = max(worst_val, dv)
            assert dv < VALUE_TOL, (name, t, dv)
            sim.select()
            np.testing.assert_allclose(sim.action_values[0].numpy(), tr.get(t, "action_values"), rtol=0,
                                       atol=VALUE_TOL)
            n_dec += 1
            if int(sim.argmax[0]) != st["argmax"]:
                n_flip += 1
                assert st["top2_gap"] < ARGMAX_GAP, (name, t, st["top2_gap"])
        elif "argmax" in st:
            # reach_destination short-cut: zero action without lookahead
            sim.la_reward.zero_(); sim.values.zero_()
            sim.select()
            assert int(sim.argmax[0]) == 0
        # committed step with the REFERENCE's chosen action
        act = torch.tensor(tr.get(t, "action")[None], dtype=torch.float64)
        sim.step(action=act)
        assert int(sim.event[0]) == st["event"], (name, t)
        assert bool(sim.done[0]) == st["done"], (name, t)
        assert abs(float(sim.reward[0]) - st["reward"]) < POS_TOL
        info = tr.get(t, "info")
        if not np.isnan(info[0]):      # info.py:141-146: Nothing carries no dist_to_goal
            assert abs(float(sim.dist_to_goal[0]) - info[0]) < POS_TOL
        for k in range(3):
            a, b = float(sim.dmin[0, k]), info[1 + k]
            assert (np.isinf(a) and np.isinf(b)) or abs(a - b) < POS_TOL, (name, t, k, a, b)
        np.testing.assert_allclose(sim.hum_pv[0, :H].numpy(), tr.get(t, "after_hum_pv"), rtol=0, atol=POS_TOL)
        np.testing.assert_allclose(sim.rob_pv[0].numpy(), tr.get(t, "after_rob_pv"), rtol=0, atol=POS_TOL)
        assert abs(float(sim.rob_theta[0]) - float(tr.get(t, "after_rob_theta"))) < POS_TOL
        assert float(sim.time[0]) == float(tr.get(t, "after_time"))
    print("%s: decisions=%d argmax flips=%d max|dvin|=%.2e max|dV|=%.2e" % (name, n_dec, n_flip, worst_vin, worst_val))
    assert n_flip <= max(1, n_dec // 50)
    if "train" in name:
        assert n_last >= 20        # every decision of a train-phase trace carries last_state
    if "linear" in name:
        assert 1 in tr.cfg_dict["human_policy"]


@pytest.mark.parametrize("name", ob.LINEAR_SCENES)
def test_reference_collision_scenes_rollout(oracle, name):
    """tests/test_collisions_simulation.py:12-39 of the reference: linear robot, ORCA humans, the
    episode must end in the pinned Info class.  Free-running fp32 rollout (no re-sync)."""
    tr = ob.Trace(name)
    H, S, R = tr.dims()
    cfg = tr.sim_config()
    sim = BatchedSim(cfg, 1, H, S, R, 1, device="cpu", backend=oracle)
    tr.load_into(sim, 0)
    done, t = False, 0
    while not done and t < 500:
        # simulator/policy/linear.py:17-23 for the robot (host side, float64)
        px, py = sim.rob_pv[0, 0].item(), sim.rob_pv[0, 1].item()
        gx, gy, vp = sim.rob_gr[0, 0].item(), sim.rob_gr[0, 1].item(), sim.rob_gr[0, 2].item()
        th = np.arctan2(gy - py, gx - px)
        act = torch.tensor([[np.cos(th) * vp, np.sin(th) * vp]], dtype=torch.float64)
        sim.step(action=act, fused_orca=True)
        done = bool(sim.done[0])
        t += 1
    assert int(sim.event[0]) == tr.meta["final_event"], (name, t)
    assert abs(t - tr.n_steps) <= 1


def test_robot_orca_trace(oracle):
    """Imitation-learning robot (rl/train.py:99-143): ORCA with safety_space 0.15 over humans + static discs."""
    tr = ob.Trace("trace_orca_robot_h10_seed5")
    H, S, R = tr.dims()
    sim = BatchedSim(tr.sim_config(), 1, H, S, R, 1, device="cpu", backend=oracle)
    for t in range(tr.n_steps):
        tr.load_into(sim, t)
        a = sim.robot_orca(0.15)
        assert np.array_equal(a[0].numpy(), tr.get(t, "action")), t
        sim.step(action=a, fused_orca=True)
        assert int(sim.event[0]) == tr.steps[t]["event"]
        np.testing.assert_allclose(sim.hum_pv[0, :H].numpy(), tr.get(t, "after_hum_pv"), rtol=0, atol=POS_TOL)


def test_rect_decomposition_matches_grid():
    """env.py:227-261 window-sum test on scene.map == AABB overlap on our rectangles."""
    from ebc.scene import rects_from_zero_cells
    for name in ("trace_cfg2_h10_seed7", "trace_adults3_bikes3_static2_seed1002", "scene_collision_with_static"):
        zero = ob.Trace(name).map_zero()
        rects = rects_from_zero_cells(zero)
        rebuilt = np.zeros_like(zero)
        for x0, y0, x1, y1 in rects:
            assert not rebuilt[x0:x1, y0:y1].any()      # disjoint
            rebuilt[x0:x1, y0:y1] = True
        assert np.array_equal(rebuilt, zero)
    rng = np.random.default_rng(0)
    zero = rng.random((40, 40)) < 0.3
    rects = rects_from_zero_cells(zero)
    rebuilt = np.zeros_like(zero)
    for x0, y0, x1, y1 in rects:
        rebuilt[x0:x1, y0:y1] = True
    assert np.array_equal(rebuilt, zero)
