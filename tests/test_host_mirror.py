"""Host-side mirror of the reference API: CPU-only checks (no compute on the device)."""
import configparser
import os

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

import oracle_backend as ob

CFG = os.path.join(ob.GOLDEN, "configs")


def read(name):
    cp = configparser.RawConfigParser()
    assert cp.read(os.path.join(CFG, name))
    return cp


@pytest.mark.parametrize("tag,cfg", [("adults5", "env_adults_5.config"), ("ebcadrl", "env_ebcadrl.config")])
def test_scene_generator_replays_reference_rng(tag, cfg):
    """Same seed -> same agents, static discs and occupancy grid as the reference (tests/golden/scenes.npz)."""
    from simulator.agents.robot import Robot
    from simulator.scene.scene_generator import SceneGenerator
    z = np.load(os.path.join(ob.GOLDEN, "scenes.npz"))
    cp = read(cfg)
    sg = SceneGenerator(cp)
    rb = Robot(cp, "robot")
    rb.policy = type("P", (), dict(multiagent_training=True, name="x", kinematics="holonomic"))()
    sg.set_robot(rb)
    for seed in (1002, 1003, 1000000):
        rb.set(0, -sg.circle_radius, 0, sg.circle_radius, 0, 0, np.pi / 2)
        sg.generate_random_scene({"test": 1000, "train": 2000, "val": 0}, "test", scene_number=seed)
        hs = sg.adults + sg.bicycles + sg.children
        assert np.array_equal(np.array([[h.px, h.py, h.vx, h.vy] for h in hs]), z["%s_%d_hum_pv" % (tag, seed)])
        assert np.array_equal(np.array([[h.gx, h.gy, h.v_pref, h.radius] for h in hs]), z["%s_%d_hum_gr" % (tag, seed)])
        assert np.array_equal(np.array([int(h.agent_type) for h in hs], np.uint8), z["%s_%d_hum_type" % (tag, seed)])
        st = np.array([[s.px, s.py, s.radius, 0.0] for s in sg.static_obstacles_as_pedestrians]).reshape(-1, 4)
        assert np.array_equal(st, z["%s_%d_stat" % (tag, seed)])
        assert np.array_equal(np.packbits((sg.map == 0).astype(np.uint8), axis=None), z["%s_%d_map_zero" % (tag, seed)])


def test_scene_save_load_round_trip(tmp_path):
    """tests/test_save_load_map.py of the reference: scene.map identical after save_scene -> load_scene."""
    from simulator.agents.robot import Robot
    from simulator.scene.scene_generator import SceneGenerator
    cp = read("env_adults_3_bikes_3_static_2.config")
    sg = SceneGenerator(cp)
    rb = Robot(cp, "robot")
    rb.policy = type("P", (), dict(multiagent_training=True, name="x", kinematics="holonomic"))()
    rb.set(0, -sg.circle_radius, 0, sg.circle_radius, 0, 0, np.pi / 2)
    sg.set_robot(rb)
    path = str(tmp_path / "scene.json")
    sg.generate_random_scene({"test": 1000, "train": 2000, "val": 0}, "test", save_scene_path=path, scene_number=5)
    grid, discs = sg.map.copy(), [(s.px, s.py, s.radius) for s in sg.static_obstacles_as_pedestrians]
    agents = [(a.px, a.py, a.gx, a.gy) for a in sg.adults + sg.bicycles]
    sg.load_scene("test", path)
    assert np.array_equal(sg.map, grid)
    assert discs == [(s.px, s.py, s.radius) for s in sg.static_obstacles_as_pedestrians]
    assert agents == [(a.px, a.py, a.gx, a.gy) for a in sg.adults + sg.bicycles]


def test_value_network_state_dict_and_masking():
    """Shipped state_dicts load unchanged; zero-padded rows + row_count reproduce the unpadded forward."""
    from rl.policy.policy_factory import policy_factory
    pol = policy_factory["sarl"]()
    pol.configure(read("policy_ebcadrl.config"))
    sd = {k: torch.as_tensor(v) for k, v in ob.load_weights("weights_ebcadrl.npz").items()}
    pol.get_model().load_state_dict(sd)
    assert sum(p.numel() for p in pol.get_model().parameters()) == 379202
    x = torch.randn(3, 7, 17)
    pad = torch.cat([x, torch.zeros(3, 5, 17)], 1)
    a = pol.get_model()(x)
    b = pol.get_model()(pad, torch.tensor([7, 7, 7]))
    assert torch.allclose(a, b, atol=1e-6)
    with pytest.raises(NotImplementedError):
        policy_factory["lstm_rl"]()


def test_info_and_state_layout():
    from simulator.utils import info
    from simulator.utils.state import FullState, ObservableState
    row = FullState(1, 2, 3, 4, 5, 6, 7, 8, 9) + ObservableState(10, 11, 12, 13, 14, 2)
    assert row == (1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 2)       # state.py:19-30,65-66
    d = info.from_code(1, 3.0, (0.5, 0.15, float("inf")), (0.1, 0.2, 0.2))
    assert isinstance(d, info.Danger) and d.min_dist == 0.15 and str(d) == "Too close"
    assert str(info.from_code(2, 0.1, (1, 1, 1))) == "Reaching goal"
    assert info.from_code(0, 0.0, (1, 2, 3)).dist_to_goal is None


def _kin_agent(kinematics, dt):
    from simulator.agents.agent import Agent
    cfg = configparser.RawConfigParser()
    cfg.add_section("a")
    for k, v in (("visible", "true"), ("v_pref", "1"), ("radius", "0.3"), ("policy", "none"), ("sensor", "coordinates")):
        cfg.set("a", k, v)
    agent = Agent(cfg, "a")
    agent.set_policy(type("P", (), {"kinematics": kinematics})())
    agent.time_step = dt
    return agent


def test_agent_kinematics_match_reference():
    """Host-side kinematics of the plugin classes against the unmodified reference's Agent on 256 random
    poses x 3 action kinds (tests/golden/agent_kinematics.npz, written by make_agent_golden.py from
    /root/reference/simulator/agents/agent.py:80-93,164-234).  Bit-exact: same float64 operations in the same order."""
    from simulator.utils.action import ActionRot, ActionXY, ActionXYRot
    from simulator.utils.utils import AgentType
    g = np.load(os.path.join(ob.GOLDEN, "agent_kinematics.npz"))
    pose, act, dt = g["pose"], g["act"], float(g["dt"])
    for kind in ("xy", "rot", "xyrot"):
        agent = _kin_agent("holonomic" if kind == "xy" else "unicycle", dt)
        for i in range(len(pose)):
            a = {"xy": ActionXY(act[i, 0], act[i, 1]), "rot": ActionRot(act[i, 0], act[i, 2]),
                 "xyrot": ActionXYRot(act[i, 0], act[i, 1], act[i, 2])}[kind]
            agent.set(*pose[i, 0:7], radius=pose[i, 7], v_pref=pose[i, 8], agent_type=AgentType.ADULT)
            assert bool(agent.reached_destination()) == bool(g[kind + "_reach"][i])
            assert tuple(agent.compute_position(a, dt)) == tuple(g[kind + "_pos"][i])
            if kind != "xy":
                assert tuple(agent.compute_velocity(a)) == tuple(g[kind + "_vel"][i])
            if kind != "xyrot":
                o = agent.get_next_observable_state(a)
                assert (o.px, o.py, o.vx, o.vy, o.radius) == tuple(g[kind + "_next"][i]) and o.obj_type == AgentType.ADULT
            agent.step(a)
            if kind != "xyrot":
                assert (agent.px, agent.py, agent.vx, agent.vy, agent.theta) == tuple(g[kind + "_step"][i])
            else:   # the reference's step() raises on ActionXYRot (reads action.v); here it follows compute_velocity
                th = (pose[i, 6] + act[i, 2]) % (2 * np.pi)
                assert agent.theta == th and (agent.px, agent.py) == tuple(g["xyrot_pos"][i])
                assert agent.vx == act[i, 0] * np.cos(th) - act[i, 1] * np.sin(th)
    holo = _kin_agent("holonomic", dt)
    holo.set(0, 0, 1, 1, 0, 0, 0)
    with pytest.raises(AssertionError):                      # agent.py:158-162
        holo.compute_position(ActionRot(1.0, 0.1), dt)
    d = holo.get_state_dict()
    other = _kin_agent("holonomic", dt)
    other.set_from_state_dict(d)
    assert other.get_full_state()._tuple() == holo.get_full_state()._tuple()
    assert other.get_observable_state()._tuple() == holo.get_observable_state()._tuple()


def _rank_main(rank, world, port, tmp):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from rl.policy.policy_factory import policy_factory
    from rl.utils.explorer import Explorer
    from rl.utils.memory import ReplayMemory
    from rl.utils.trainer import Trainer
    torch.manual_seed(0)
    pol = policy_factory["sarl"]()
    pol.configure(read("policy.config"))
    model = pol.get_model()
    mem = ReplayMemory(64, device="cpu")
    g = torch.Generator().manual_seed(100 + rank)             # every rank owns a different replay shard
    mem.push_batch(torch.randn(40, 5, 13, generator=g), torch.full((40,), 5), torch.randn(40, generator=g))
    tr = Trainer(model, mem, "cpu", 8, policy=pol)
    tr.set_optimizer(0.01)
    tr.optimize_batch(3)
    # imitation-learning shape: per-rank shards whose sizes straddle the batch size (ceil(40/16) = 3 vs ceil(9/16) = 1
    # steps per epoch if nobody agreed) and, in a second pass, one EMPTY shard -- every rank must still issue the
    # same number of all-reduces or the job deadlocks
    small = ReplayMemory(64, device="cpu")
    if rank == 0:
        small.push_batch(torch.randn(9, 5, 13, generator=g), torch.full((9,), 5), torch.randn(9, generator=g))
    t2 = Trainer(model, mem if rank == 1 else small, "cpu", 16, policy=pol)
    t2.set_optimizer(0.01)
    t2.optimize_epoch(2)                                       # 40 vs 9 samples
    empty = ReplayMemory(64, device="cpu")
    t3 = Trainer(model, mem if rank == 0 else empty, "cpu", 16, policy=pol)
    t3.set_optimizer(0.01)
    t3.optimize_epoch(1)                                       # 40 vs 0 samples
    t3.optimize_batch(2)
    flat = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
    out = [torch.zeros_like(flat) for _ in range(world)]
    dist.all_gather(out, flat)
    same = all(torch.equal(out[0], o) for o in out)            # grad all-reduce keeps the replicas identical
    # episode statistics: rank-local arrays are gathered before logging
    ex = Explorer(type("E", (), dict(time_limit=25, N=1))(), robot=object(), device="cpu")
    stats = type("S", (), dict(event=np.array([2, 3]) if rank == 0 else np.array([2, 7]), time=np.array([5.0, 6.0]),
                               steps=np.array([20, 24]), cum_reward=np.array([0.5, -0.2]), too_close=np.array([1, 0]),
                               min_dist_sum=np.array([0.1, 0.0])))()
    summary = ex._gather(stats)
    m = ex.log_results(summary, "test")
    torch.save({"same": same, "version": pol.weights_version, "n": int(summary[13]), "sr": m["success_rate"]},
               os.path.join(tmp, "r%d.pt" % rank))
    dist.destroy_process_group()


def test_two_rank_gloo_training_and_stats(tmp_path):
    """world_size 2 on CPU: the N > 1 host logic (gradient all-reduce, statistics gather)."""
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_rank_main, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        d = torch.load(os.path.join(str(tmp_path), "r%d.pt" % r))
        assert d["same"] and d["version"] == 3 + 2 * 3 + 3 + 2 and d["n"] == 4 and d["sr"] == 0.5


def test_episode_sharding_is_independent_of_world_size():
    """Synthetic scenes are keyed by the global episode id: any split over ranks gives the same episodes."""
    from ebc import synth
    whole = synth.generate(synth.CFG2, np.arange(64))
    for world in (2, 4, 8):
        per = 64 // world
        parts = [synth.generate(synth.CFG2, np.arange(r * per, (r + 1) * per)) for r in range(world)]
        for k in whole:
            assert np.array_equal(np.concatenate([p[k] for p in parts]), whole[k]), (world, k)


def _explorer_with(gamma, dt, vp, target_model=None):
    from rl.utils.explorer import Explorer
    from rl.utils.memory import ReplayMemory
    env = type("E", (), dict(time_limit=25, N=1, time_step=dt, robot=type("R", (), dict(v_pref=vp))()))()
    ex = Explorer(env, robot=object(), device="cpu", memory=ReplayMemory(4096, device="cpu"), gamma=gamma)
    if target_model is not None:
        ex.update_target_model(target_model)
    return ex


def test_update_memory_matches_reference():
    """a20: the (state, value) pairs our Explorer.update_memory pushes == the reference's own Explorer.update_memory /
    ParallelExplorer.update_memory (rl/utils/explorer.py:151-200, parallel_explorer.py:286-330) on trajectories the
    reference produced itself (tests/golden/update_memory.npz, written by make_golden.py: update_memory_golden).
    TD target r_i + gamma^(dt v_pref) V_target(s_{i+1}) (terminal: r_i), IL discounted returns, to 1e-5; and the
    sequential explorer's admission rule (only success / collision episodes, explorer.py:82-92)."""
    from rl.policy.policy_factory import policy_factory
    z = np.load(os.path.join(ob.GOLDEN, "update_memory.npz"))
    gamma, dt, vp = [float(x) for x in z["gamma_dt_vpref"]]
    pol = policy_factory["sarl"]()
    pol.configure(read("policy.config"))
    pol.get_model().load_state_dict({k: torch.as_tensor(v) for k, v in ob.load_weights("weights_sarl_baseline.npz").items()})
    states = torch.as_tensor(z["rl_states"])                     # [T, n, D] = policy.last_state of every step
    rewards = torch.as_tensor(z["rl_rewards"])
    T, n, D = states.shape
    assert T == 74 and np.array_equal(z["rl_mem_states"], z["rl_states"])
    traj = {"states": states[:, None], "rewards": rewards[:, None], "alive": torch.ones(T, 1, dtype=torch.bool),
            "rows": torch.tensor([n], dtype=torch.int32)}
    stats = type("S", (), dict(event=np.array([ob.abi.EV_REACH_GOAL])))()
    # ---- RL: TD targets through the target network ----
    ex = _explorer_with(gamma, dt, vp, pol.get_model())
    ex.update_memory(traj, stats, imitation_learning=False)
    assert len(ex.memory) == T
    assert torch.equal(ex.memory.states[:T], states)
    got = ex.memory.values[:T, 0].numpy()
    np.testing.assert_allclose(got, z["rl_values"], rtol=0, atol=1e-5)
    assert got[-1] == np.float32(rewards[-1])                    # terminal state: value = reward
    # ---- IL: discounted returns sum_{t >= i} gamma^((t - i) dt v_pref) r_t ----
    ex = _explorer_with(gamma, dt, vp)
    ex.update_memory(traj, stats, imitation_learning=True)
    np.testing.assert_allclose(ex.memory.values[:T, 0].numpy(), z["il_values_of_rl_rewards"], rtol=0, atol=1e-5)
    il_r = torch.as_tensor(z["il_rewards"])
    Ti = len(il_r)
    il_states = torch.as_tensor(z["il_states"])                  # transform(JointState) of the ORCA robot's episode
    traj = {"states": il_states[:, None], "rewards": il_r[:, None], "alive": torch.ones(Ti, 1, dtype=torch.bool),
            "rows": torch.tensor([il_states.shape[1]], dtype=torch.int32)}
    ex = _explorer_with(gamma, dt, vp)
    ex.update_memory(traj, stats, imitation_learning=True)
    np.testing.assert_allclose(ex.memory.values[:Ti, 0].numpy(), z["il_values"], rtol=0, atol=1e-5)
    assert torch.equal(ex.memory.states[:Ti], il_states)
    # ---- admission: the reference tried scene 2005 (Timeout: nothing stored) then 2006 (ReachGoal: 41 pairs) ----
    import json
    tried = json.loads(bytes(z["il_tried_json"]).decode())
    assert tried == [[5, "Timeout", 0], [6, "ReachGoal", Ti]]
    for event, stored in ((ob.abi.EV_TIMEOUT, 0), (ob.abi.EV_REACH_GOAL, Ti), (ob.abi.EV_COLLISION_CHILD, Ti),
                          (ob.abi.EV_COLLISION_OBSTACLE, Ti)):
        ex = _explorer_with(gamma, dt, vp)
        ex.update_memory(traj, type("S", (), dict(event=np.array([event])))(), imitation_learning=True, store_all=False)
        assert len(ex.memory) == stored, (event, len(ex.memory))
    ex = _explorer_with(gamma, dt, vp)                            # the parallel explorer stores every episode
    ex.update_memory(traj, type("S", (), dict(event=np.array([ob.abi.EV_TIMEOUT])))(), imitation_learning=True, store_all=True)
    assert len(ex.memory) == Ti
    # ragged batch: two episodes of different length in one trajectory tensor (alive mask) give the same pairs
    T2 = 30
    st2 = torch.zeros(T, 2, n, D); st2[:, 0] = states; st2[:T2, 1] = states[:T2]
    rw2 = torch.zeros(T, 2, dtype=torch.float64); rw2[:, 0] = rewards; rw2[:T2, 1] = rewards[:T2]
    al2 = torch.zeros(T, 2, dtype=torch.bool); al2[:, 0] = True; al2[:T2, 1] = True
    ex = _explorer_with(gamma, dt, vp, pol.get_model())
    ex.update_memory({"states": st2, "rewards": rw2, "alive": al2, "rows": torch.tensor([n, n], dtype=torch.int32)},
                     type("S", (), dict(event=np.array([2, 3])))(), imitation_learning=False)
    assert len(ex.memory) == T + T2
    vals = ex.memory.values[:T + T2, 0].numpy()
    # time-major order: steps 0 .. T2-1 hold both episodes, then episode 0 alone
    ep0 = np.concatenate([vals[0:2 * T2:2], vals[2 * T2:]])
    ep1 = vals[1:2 * T2:2]
    np.testing.assert_allclose(ep0, z["rl_values"], rtol=0, atol=1e-5)
    np.testing.assert_allclose(ep1[:-1], z["rl_values"][:T2 - 1], rtol=0, atol=1e-5)
    assert ep1[-1] == np.float32(rewards[T2 - 1])                # its last step is terminal: no bootstrap


def test_synth_wall_geometry_matches_reference():
    """f-1: walls -> occupancy grid and -> static discs of the synthetic generator (ebc/synth.py: wall_geometry, which
    the device generator ebc_generate is held equal to on the GPU) against the REFERENCE's own place_obstacles_on_map /
    create_observation_from_static_obstacles (tests/golden/static_decomposition.npz: 80 scenes x 4 walls, lengths 1-5,
    walls hanging over the map border included)."""
    from ebc import synth
    z = np.load(os.path.join(ob.GOLDEN, "static_decomposition.npz"))
    n_clipped = n_square = 0
    for k in range(int(z["n_scenes"][0])):
        walls = z["s%03d_walls" % k]
        size, res = z["s%03d_map" % k]
        G = int(round(size / res))
        zero = np.unpackbits(z["s%03d_zero" % k])[:G * G].reshape(G, G).astype(bool)
        mine = np.zeros_like(zero)
        discs = []
        for lx, ly, xd, yd in walls:
            r, dl = synth.wall_geometry(np.array([float(lx)]), np.array([float(ly)]), np.array([float(xd)]),
                                        np.array([float(yd)]), G, res, 8)
            x0, y0, x1, y1 = [int(v) for v in r[0]]
            mine[x0:x1, y0:y1] = True
            n_clipped += (x1 - x0) * (y1 - y0) != round(xd / res) * round(yd / res)
            n_square += xd == yd
            discs += [(sx[0], sy[0], rr[0]) for sx, sy, rr, live in dl if live[0]]
        assert np.array_equal(mine, zero), k
        ref = z["s%03d_discs" % k]
        assert len(discs) == len(ref), (k, len(discs), len(ref))
        np.testing.assert_allclose(np.array(discs).reshape(-1, 3), ref, rtol=0, atol=1e-12)
    assert n_clipped >= 10 and n_square >= 10      # border-clipped and square walls were part of the fixture


@pytest.mark.parametrize("shape_name", ["CFG1", "CFG2", "CFG3", "CFG4", "MIXED20", "ONE_STATIC"])
def test_synth_scenes_obey_reference_placement_rules(shape_name):
    """f-1: start / goal geometry, minimum separation, wall clearance, disc radii of scene_generator.py on the host
    generator's output (tests/scene_rules.py restates each rule with its reference line)."""
    from ebc import synth
    from scene_rules import check_scene_rules
    shape = getattr(synth, shape_name)
    res = check_scene_rules(shape, synth.generate(shape, np.arange(5000, 5000 + 768)))
    print(shape.name, res)


def test_trainer_flat_sgd_equals_torch_sgd():
    """The trainer's flat-buffer update == torch.optim.SGD(lr, momentum=0.9) of the reference (trainer.py:21-27,74-100)
    on identical batches: same parameters after every step (to fp32 rounding of one fused multiply-add)."""
    import copy
    from rl.policy.policy_factory import policy_factory
    from rl.utils.memory import ReplayMemory
    from rl.utils.trainer import Trainer
    torch.manual_seed(1)
    pol = policy_factory["sarl"]()
    pol.configure(read("policy.config"))
    model = pol.get_model()
    ref = copy.deepcopy(model)
    mem = ReplayMemory(256, device="cpu")
    g = torch.Generator().manual_seed(7)
    mem.push_batch(torch.randn(200, 5, 13, generator=g), torch.randint(1, 6, (200,), generator=g), torch.randn(200, generator=g))
    tr = Trainer(model, mem, "cpu", 32, policy=pol)
    tr.set_optimizer(0.01)
    opt = torch.optim.SGD(ref.parameters(), lr=0.01, momentum=0.9)
    crit = torch.nn.MSELoss()
    total = 0.0
    for step in range(6):
        idx = torch.randint(0, 200, (32,), generator=g)
        tr._step(index=idx)
        opt.zero_grad()
        loss = crit(ref(mem.states[idx], mem.rows[idx]), mem.values[idx])
        loss.backward()
        opt.step()
        total += loss.item()
        for (k, a), b in zip(model.state_dict().items(), ref.state_dict().values()):
            assert torch.allclose(a, b, rtol=1e-6, atol=1e-8), (step, k, (a - b).abs().max())
    assert abs(tr._take_loss() - total) < 1e-6 and pol.weights_version == 6
    # the module's tensors are views of the flat buffers (no gather / scatter around the all-reduce)
    assert all(p.data.untyped_storage().data_ptr() == tr._flat.untyped_storage().data_ptr() for p in model.parameters())
    sd = copy.deepcopy(model.state_dict())
    model.load_state_dict(sd)                      # checkpoint I/O still works on the views
