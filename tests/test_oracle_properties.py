"""Size-independent properties of the path, checked on the CPU oracle (the checker the GPU tests compare against;
oracle/ebc_oracle.c).  The reference pins no float at the rvo2 boundary, so besides the golden traces the
oracle is held to what the algorithms themselves guarantee:

  * RVO2 (a9): the new velocity lies inside the speed disc; an agent without neighbours takes its preferred velocity;
    the result depends on relative positions only; reciprocal agents that start apart never overlap
    (the ORCA guarantee, here with the reference's +0.01 radius pad of simulator/policy/orca.py:113-127);
  * env.py:207-209: `onestep_lookahead(a)` IS `step(a, update=False)` -- reward, done and event of every action equal
    what the committed step then returns from the same state, and the lookahead leaves the state untouched;
  * episodes never interact: any permutation of the batch permutes the results, bit for bit.
"""
import numpy as np
import torch

from ebc import synth
from ebc.actions import build_action_space
from ebc.config import SimConfig
from ebc.engine import BatchedSim


def _sim(oracle, shape, N, visible=False, time_limit=100.0):
    c = SimConfig()
    c.map_size_m, c.map_resolution, c.time_limit = shape.map_size_m, shape.map_resolution, time_limit
    c.robot_visible = visible
    # reward block of data/eb-cadrl/adults_8_..._fix_static.config:26-44 (the cfg2 workload of bench.py)
    c.new_reward, c.time_max, c.time_good, c.max_goal_distance = True, 35.0, 10.0, 10.0
    c.collision_penalty_adult, c.collision_penalty_bicycle = -1.0, -1.5
    c.collision_penalty_child, c.collision_penalty_obstacle = -2.0, -0.5
    c.discomfort_dist = c.discomfort_dist_adult = 0.1
    c.discomfort_dist_bicycle = c.discomfort_dist_child = 0.2
    c.discomfort_penalty_factor_bicycle = c.discomfort_penalty_factor_child = 1.0
    sim = BatchedSim(c, N, shape.H, shape.Smax, shape.Rmax, 81, device="cpu", backend=oracle)
    sim.set_actions(build_action_space(shape.robot_v_pref))
    return sim


def test_orca_velocity_inside_speed_disc_and_lone_agent(oracle):
    shape, N = synth.CFG2, 256
    sim = _sim(oracle, shape, N, visible=True)
    synth.load(sim, synth.generate(shape, np.arange(N)))
    sim.orca()
    H = sim.hum_count.numpy()
    speed = np.linalg.norm(sim.hum_nv.numpy().astype(np.float64), axis=2)
    v_pref = sim.hum_gr.numpy()[:, :, 2].astype(np.float64)
    live = np.arange(shape.H)[None] < H[:, None]
    # linearProgram1/2 keep the result on or inside the maxSpeed disc up to fp32 rounding of `point + t * direction`
    # (t from sqrt(discriminant): measured excess <= 3.6e-5 relative on these scenes)
    assert (speed[live] <= v_pref[live] * (1 + 1e-4)).all()
    assert (speed[live] > 0).any()
    # one human, invisible robot, no neighbour: the preferred velocity (orca.py:136-140: goal offset, unit length beyond
    # 1 m -- not scaled by v_pref) clamped to the speed disc by the LP
    lone = synth.SceneShape("lone", [(0, 1, (0.4, 1.6), (0.3, 0.3))], rule="circle_crossing", circle_radius=4.0, num_walls=0)
    sim = _sim(oracle, lone, 64)
    sc = synth.generate(lone, np.arange(64))
    sc["hum_gr"][::4, 0, 0:2] = sc["hum_pv"][::4, 0, 0:2] + np.float32(0.25)       # some goals closer than 1 m
    synth.load(sim, sc)
    sim.orca()
    p, g = sc["hum_pv"][:, 0, 0:2].astype(np.float64), sc["hum_gr"][:, 0, 0:2].astype(np.float64)
    vp = sc["hum_gr"][:, 0, 2].astype(np.float64)
    d = g - p
    dist = np.linalg.norm(d, axis=1, keepdims=True)
    pref = np.where(dist > 1.0, d / dist, d)
    pn = np.linalg.norm(pref, axis=1, keepdims=True)
    want = np.where(pn > vp[:, None], pref / pn * vp[:, None], pref)
    assert np.allclose(sim.hum_nv.numpy()[:, 0].astype(np.float64), want, atol=2e-6)


def test_orca_depends_on_relative_positions_only(oracle):
    """Positions on a 2^-8 grid, the whole scene shifted by (16, -32): every difference RVO2 forms is the same fp32
    number, so the velocities are bit-identical."""
    shape, N = synth.CFG2, 128
    sc = synth.generate(shape, np.arange(N))
    q = np.float32(256.0)
    for key in ("hum_pv", "hum_gr", "rob_pv", "rob_gr"):
        sc[key][..., 0:2] = np.round(sc[key][..., 0:2] * q) / q
    out = []
    for shift in ((0.0, 0.0), (16.0, -32.0)):
        s2 = {k: np.array(v, copy=True) for k, v in sc.items()}
        for key in ("hum_pv", "hum_gr", "rob_pv", "rob_gr"):
            s2[key][..., 0] += np.float32(shift[0])
            s2[key][..., 1] += np.float32(shift[1])
        sim = _sim(oracle, shape, N, visible=True)
        synth.load(sim, s2)
        sim.orca()
        out.append(sim.hum_nv.numpy().copy())
    assert np.array_equal(out[0], out[1])
    assert np.abs(out[0]).max() > 0.1


def test_reciprocal_orca_agents_never_overlap(oracle):
    """Circle crossing through the centre: every pair that starts apart stays apart (RVO2's guarantee for agents
    that all run ORCA; the radii carry the reference's +0.01 pad, so the discs themselves keep a small gap)."""
    shape = synth.SceneShape("cross8", [(0, 8, (1.0, 1.0), (0.3, 0.3))], rule="circle_crossing", circle_radius=4.0,
                             robot_v_pref=1.0, num_walls=0)
    N = 96
    sim = _sim(oracle, shape, N)                     # invisible robot: the humans see each other only
    sc = synth.generate(shape, np.arange(N))
    sc["rob_pv"][:, 0:2] = 40.0                      # and it stands far away
    sc["rob_gr"][:, 0:2] = 40.0
    synth.load(sim, sc)
    r = sim.hum_gr.numpy()[:, :, 3].astype(np.float64)
    need = r[:, :, None] + r[:, None, :]
    iu = np.triu_indices(shape.H, 1)
    zero = torch.zeros(N, dtype=torch.int32)
    worst, moved = np.inf, 0.0
    p0 = sim.hum_pv.numpy()[:, :, 0:2].astype(np.float64)
    for _ in range(80):
        sim.step(action_idx=zero, fused_orca=True)
        p = sim.hum_pv.numpy()[:, :, 0:2].astype(np.float64)
        gap = np.linalg.norm(p[:, :, None] - p[:, None, :], axis=3) - need
        worst = min(worst, gap[:, iu[0], iu[1]].min())
        moved = max(moved, np.linalg.norm(p - p0, axis=2).max())
    assert worst > 0.0, worst
    assert moved > 6.0                               # they did cross the circle


def test_lookahead_is_the_uncommitted_step(oracle):
    shape, N = synth.CFG2, 48
    sc = synth.generate(shape, np.arange(N))
    sim = _sim(oracle, shape, N, visible=True, time_limit=35.0)
    synth.load(sim, sc)
    for _ in range(6):                               # a few steps in, so that Danger / collision events occur
        sim.orca()
        sim.step(action_idx=torch.full((N,), 40, dtype=torch.int32))
    keep = {k: getattr(sim, k).clone() for k in ("hum_pv", "rob_pv", "rob_theta", "time")}
    sim.orca()
    sim.lookahead(build_inputs=False)
    for k, v in keep.items():                        # onestep_lookahead must not mutate (env.py:207-209)
        assert torch.equal(getattr(sim, k), v), k
    la_r, la_d, la_e = sim.la_reward.clone(), sim.la_done.clone(), sim.la_event.clone()
    seen = set()
    for a in (0, 3, 17, 40, 64, 80):
        for k, v in keep.items():
            getattr(sim, k).copy_(v)
        sim.orca()
        sim.step(action_idx=torch.full((N,), a, dtype=torch.int32))
        assert torch.equal(sim.reward, la_r[:, a]), a
        assert torch.equal(sim.done, la_d[:, a]), a
        assert torch.equal(sim.event, la_e[:, a]), a
        seen |= set(sim.event.tolist())
    assert len(seen) >= 3 and not torch.isnan(la_r).any()


def test_episodes_do_not_interact(oracle):
    shape, N = synth.CFG2, 96
    sc = synth.generate(shape, np.arange(N))
    perm = np.random.default_rng(5).permutation(N)
    res = []
    for order in (np.arange(N), perm):
        sim = _sim(oracle, shape, N, visible=True)
        synth.load(sim, {k: v[order] for k, v in sc.items()})
        for _ in range(3):
            sim.orca()
            sim.lookahead(build_inputs=False)
            sim.step(action_idx=torch.full((N,), 21, dtype=torch.int32))
        res.append((sim.hum_pv.numpy().copy(), sim.la_reward.numpy().copy(), sim.la_event.numpy().copy(),
                    sim.reward.numpy().copy(), sim.event.numpy().copy()))
    for a, b in zip(*res):
        assert np.array_equal(a[perm], b)
