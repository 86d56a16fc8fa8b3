#!/usr/bin/env python
"""Generate golden input/output vectors by running the UNMODIFIED reference.

Runs only in the build container (needs /root/reference); the outputs (*.npz, *.json in
this directory) are committed and are what the tests read on the GPU box.

    PYTHONPATH=oracle/shims:/root/reference python tests/golden/make_golden.py   (cwd = /root/reference)
or simply
    python tests/golden/make_golden.py          (re-executes itself with that environment)

What is recorded (all from reference code, nothing from this repo's package):
  * the reference's own known-answer cases (tests/test_collisions.py:13-143)
  * the action tables of rl/policy/cadrl.py:91-116
  * step traces: from an fp32-representable state, the humans' ORCA actions, all 81
    onestep_lookahead outcomes, the rotated value-network inputs and outputs, the argmax,
    and the committed env.step result; the state is re-rounded to fp32 after every step so
    that every step is a single-step comparison "from identical states".
  * episode outcomes of the reference's fixed scenes (tests/test_collisions_simulation.py).
  * a few generated scenes (scene_generator.py) for the host-side scene generator.

`rvo2` is supplied by oracle/shims/rvo2.py (the C restatement in oracle/ebc_oracle.c); see
that file's header for what this does and does not pin.
"""
import configparser
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("EBC_REFERENCE", "/root/reference")

if os.environ.get("EBC_GOLDEN_CHILD") != "1":
    env = dict(os.environ)
    env["EBC_GOLDEN_CHILD"] = "1"
    env["PYTHONPATH"] = os.pathsep.join([os.path.join(REPO, "oracle", "shims"), REF, os.path.join(REF, "tests")])
    sys.exit(subprocess.call([sys.executable, os.path.abspath(__file__)] + sys.argv[1:], cwd=REF, env=env))

import numpy as np  # noqa: E402
import torch  # noqa: E402
import gym  # noqa: E402
from simulator.agents.robot import Robot  # noqa: E402
from simulator.utils.state import ObservableState  # noqa: E402
from simulator.utils.info import *  # noqa: E402,F401,F403
from simulator.utils import info as info_mod  # noqa: E402
from rl.policy.policy_factory import policy_factory  # noqa: E402

torch.set_num_threads(1)

EVENT_CODE = {"Nothing": 0, "Danger": 1, "ReachGoal": 2, "CollisionAdult": 3, "CollisionBicycle": 4,
              "CollisionChild": 5, "CollisionObstacle": 6, "Timeout": 7}


def f32(x):
    return float(np.float32(x))


def make(env_cfg_path, policy_cfg_path, model_path, policy_name, phase="test", env_overrides=None,
         policy_overrides=None):
    env_config = configparser.RawConfigParser()
    env_config.read(env_cfg_path)
    for (sec, key), val in (env_overrides or {}).items():
        env_config.set(sec, key, str(val))
    env = gym.make("EntityBasedCollisionAvoidance-v0")
    env.configure(env_config)
    env.get_local_map_angular = lambda *a, **k: None  # unused by every shipped policy
    robot = Robot(env_config, "robot")
    env.set_robot(robot)
    policy = policy_factory[policy_name]()
    policy_config = None
    if policy_cfg_path is not None:
        policy_config = configparser.RawConfigParser()
        policy_config.read(policy_cfg_path)
        for (sec, key), val in (policy_overrides or {}).items():
            policy_config.set(sec, key, str(val))
        policy.configure(policy_config)
    if model_path is not None:
        policy.get_model().load_state_dict(torch.load(model_path, map_location="cpu"))
    robot.set_policy(policy)
    policy.set_phase(phase)
    policy.set_device("cpu")
    return env, policy, robot, env_config, policy_config


def humans_of(env):
    return env.scene.adults + env.scene.bicycles + env.scene.children


def round_state_to_f32(env):
    """Make the reference state fp32-representable so that both sides see identical inputs."""
    for ag in [env.robot] + humans_of(env):
        for name in ("px", "py", "vx", "vy", "gx", "gy", "radius", "v_pref", "theta"):
            setattr(ag, name, f32(getattr(ag, name)))
    new_static = []
    for s in env.scene.static_obstacles_as_pedestrians:
        new_static.append(ObservableState(f32(s.px), f32(s.py), 0, 0, f32(s.radius), s.obj_type))
    env.scene.static_obstacles_as_pedestrians = new_static
    ob = [h.get_observable_state() for h in humans_of(env)]
    ob += env.scene.static_obstacles_as_pedestrians
    return ob


def snapshot(env):
    hs = humans_of(env)
    rb = env.robot
    d = {
        "hum_pv": np.array([[h.px, h.py, h.vx, h.vy] for h in hs], dtype=np.float64).reshape(-1, 4),
        "hum_gr": np.array([[h.gx, h.gy, h.v_pref, h.radius] for h in hs], dtype=np.float64).reshape(-1, 4),
        "hum_type": np.array([int(h.agent_type) for h in hs], dtype=np.uint8),
        "stat": np.array([[s.px, s.py, s.radius, 0.0] for s in env.scene.static_obstacles_as_pedestrians],
                         dtype=np.float64).reshape(-1, 4),
        "rob_pv": np.array([rb.px, rb.py, rb.vx, rb.vy], dtype=np.float64),
        "rob_gr": np.array([rb.gx, rb.gy, rb.v_pref, rb.radius], dtype=np.float64),
        "rob_theta": np.float64(rb.theta),
        "time": np.float64(env.global_time),
    }
    return d


def config_dict(env, env_config, policy, policy_config):
    rw = env.reward
    kin = getattr(policy, "kinematics", "holonomic")
    d = {
        "time_step": env.time_step, "time_limit": float(env.time_limit),
        "new_reward": int(bool(rw.new_reward)),
        "has_max_goal_distance": int(rw.max_goal_distance is not None),
        "time_max": rw.time_max if rw.time_max is not None else 0.0,
        "time_good": rw.time_good,
        "max_goal_distance": rw.max_goal_distance if rw.max_goal_distance is not None else 0.0,
        "success_reward": rw.success_reward,
        "collision_penalty_adult": rw.collision_penalty_adult or 0.0,
        "collision_penalty_bicycle": rw.collision_penalty_bicycle or 0.0,
        "collision_penalty_obstacle": rw.collision_penalty_obstacle or 0.0,
        "collision_penalty_child": rw.collision_penalty_child or 0.0,
        "discomfort_dist_adult": rw.discomfort_dist_adult,
        "discomfort_dist_bicycle": rw.discomfort_dist_bicycle,
        "discomfort_dist_child": rw.discomfort_dist_child,
        "discomfort_penalty_factor_adult": rw.discomfort_penalty_factor_adult,
        "discomfort_penalty_factor_bicycle": rw.discomfort_penalty_factor_bicycle,
        "discomfort_penalty_factor_child": rw.discomfort_penalty_factor_child,
        "rotation_penalty_factor": rw.rotation_penalty_factor,
        "map_size_m": env.scene.map_size_m, "map_resolution": env.scene.map_resolution,
        "robot_kinematics": 0 if kin == "holonomic" else 1,
        "rotate_theta": int(kin == "unicycle"),
        "with_agent_type": int(bool(getattr(policy, "with_agent_type", False))),
        "robot_visible": int(bool(env.robot.visible)),
        "gamma": getattr(policy, "gamma", None) or 0.9,
        "human_policy": [0 if env_config.get(sec, "policy", fallback="orca") == "orca" else 1
                         for sec in ("adults", "bicycles", "children") if True],
    }
    return d


def info_fields(info):
    def g(name):
        v = getattr(info, name, None)
        return np.nan if v is None else float(v)
    return np.array([g("dist_to_goal"), g("dmin_adult"), g("dmin_bicycle"), g("dmin_child"), g("min_dist")],
                    dtype=np.float64)


def trace(name, env, policy, robot, env_config, policy_config, reset_kwargs, n_steps, tensor_steps,
          out):
    """Record a step trace.  `tensor_steps` = steps whose full value-net inputs are kept."""
    env.reset("test", **reset_kwargs)
    ob = round_state_to_f32(env)
    sarl = hasattr(policy, "action_space")
    rec = {"cfg": config_dict(env, env_config, policy, policy_config), "steps": []}
    zero = (env.scene.map == 0)
    rec_map = np.packbits(zero.astype(np.uint8), axis=None)
    arrays = {"map_zero_packed": rec_map, "map_shape": np.array(env.scene.map.shape)}

    look = []
    orig_look = env.onestep_lookahead

    def look_wrap(action):
        r = orig_look(action)
        look.append(r)
        return r

    env.onestep_lookahead = look_wrap
    vin, vout = [], []
    if sarl and policy.get_model() is not None:
        def fwd_hook(m, i, o):
            vin.append(i[0].detach().numpy().copy())
            vout.append(float(o.item()))

        policy.get_model().register_forward_hook(fwd_hook)
    orca_actions = []
    orig_update = env.compute_step_update

    def update_wrap(action, agents_actions, all_agents):
        orca_actions.append([(a.vx, a.vy) for a in agents_actions])
        return orig_update(action, agents_actions, all_agents)

    env.compute_step_update = update_wrap

    done = False
    t = 0
    while not done and t < n_steps:
        del look[:], vin[:], vout[:], orca_actions[:]
        snap = snapshot(env)
        for k, v in snap.items():
            arrays["s%03d_%s" % (t, k)] = v
        action = robot.act(ob, local_map=None, env=env)
        step = {"t": t}
        if sarl and policy.phase == "train" and torch.is_tensor(getattr(policy, "last_state", None)):
            # multi_human_rl.py:84-85: the replay sample = transform(current joint state), n x D
            arrays["s%03d_last_state" % t] = policy.last_state.detach().numpy().astype(np.float32)
        if sarl and len(look) > 0:
            A = len(policy.action_space)
            assert len(look) == A and len(vout) == A
            arrays["s%03d_la_reward" % t] = np.array([r[1] for r in look], dtype=np.float64)
            arrays["s%03d_la_done" % t] = np.array([r[2] for r in look], dtype=np.uint8)
            arrays["s%03d_la_event" % t] = np.array([EVENT_CODE[type(r[3]).__name__] for r in look], dtype=np.uint8)
            arrays["s%03d_la_value" % t] = np.array(vout, dtype=np.float32)
            arrays["s%03d_action_values" % t] = np.array(policy.action_values, dtype=np.float64)
            if t in tensor_steps:
                arrays["s%03d_vin" % t] = np.concatenate(vin, axis=0).astype(np.float32)  # A x n x D
                arrays["s%03d_la_next" % t] = np.array(
                    [[[o.px, o.py, o.vx, o.vy, o.radius, int(o.obj_type)] for o in r[0]] for r in look],
                    dtype=np.float64)
            step["argmax"] = policy.action_space.index(action)
            vals = np.array(policy.action_values)
            srt = np.sort(vals)[::-1]
            step["top2_gap"] = float(srt[0] - srt[1])
        elif sarl:
            step["argmax"] = 0  # reach_destination short-cut (multi_human_rl.py:25-26)
            step["top2_gap"] = None
        arrays["s%03d_action" % t] = np.array(list(action), dtype=np.float64)
        ob_next, _, reward, done, info = env.step(action)
        arrays["s%03d_orca" % t] = np.array(orca_actions[-1], dtype=np.float64).reshape(-1, 2)
        step.update({"reward": float(reward), "done": bool(done), "event": EVENT_CODE[type(info).__name__]})
        arrays["s%03d_info" % t] = info_fields(info)
        after = snapshot(env)
        for k in ("hum_pv", "rob_pv", "rob_theta", "time"):
            arrays["s%03d_after_%s" % (t, k)] = after[k]
        rec["steps"].append(step)
        ob = round_state_to_f32(env)
        t += 1
    rec["n_steps"] = t
    rec["final_event"] = rec["steps"][-1]["event"]
    rec["final_time"] = float(env.global_time)
    if sarl:
        rec["actions"] = [list(a) for a in policy.action_space]
        arrays["actions"] = np.array([list(a) for a in policy.action_space], dtype=np.float64)
    env.onestep_lookahead = orig_look
    env.compute_step_update = orig_update
    arrays["meta_json"] = np.frombuffer(json.dumps(rec).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(out, name + ".npz"), **arrays)
    print("  %-34s steps=%d final_event=%d t=%.2f" % (name, t, rec["final_event"], rec["final_time"]))
    return rec


def save_weights(path, out_name, out):
    sd = torch.load(path, map_location="cpu")
    np.savez_compressed(os.path.join(out, out_name), **{k: v.numpy() for k, v in sd.items()})


def collisions_known_answers(out):
    """tests/test_collisions.py:13-143 restated as data: inputs + the asserted boolean."""
    import re
    src = open(os.path.join(REF, "tests", "test_collisions.py")).read()
    cases = []
    for block in src.split("    def test_")[1:]:
        nums = lambda pat: [float(x) for x in re.search(pat, block).groups()]  # noqa: E731
        rs = nums(r"robot\.set\(([-\d.]+), ([-\d.]+), ")
        rr = nums(r"robot\.radius = ([-\d.]+)")
        as_ = nums(r"adult\.set\(([-\d.]+), ([-\d.]+), ")
        ar = nums(r"adult\.radius = ([-\d.]+)")
        ts = nums(r"time_step = ([-\d.]+)")
        ac = nums(r"ActionXY\(([-\d.]+), ([-\d.]+)\)")
        exp = re.search(r"assertEqual\(result\[1\], (True|False)\)", block).group(1) == "True"
        cases.append({"name": block.split("(")[0], "robot_pos": rs, "robot_radius": rr[0], "adult_pos": as_,
                      "adult_radius": ar[0], "time_step": ts[0], "action": ac, "collision": exp})
    # cross-check against the reference function itself
    from simulator.agents.agents import Adult
    from simulator.utils.collisions import compute_collision_agent_with_robot
    from simulator.utils.action import ActionXY
    cfg = configparser.RawConfigParser()
    cfg.read("configs/env_configs/env_adults_5_bikes_5_static_5.config")
    for c in cases:
        robot = Robot(cfg, "robot")
        robot.kinematics = "holonomic"
        robot.set(c["robot_pos"][0], c["robot_pos"][1], 0, 0, 0, 0, np.pi / 2)
        robot.radius = c["robot_radius"]
        adult = Adult(cfg, "adults")
        adult.set(c["adult_pos"][0], c["adult_pos"][1], 0, 0, 0, 0, 0)
        adult.radius = c["adult_radius"]
        dmin, coll = compute_collision_agent_with_robot(adult, robot, ActionXY(*c["action"]), float("inf"), c["time_step"])
        assert coll == c["collision"], c
        c["dmin"] = None if np.isinf(dmin) else dmin
    json.dump(cases, open(os.path.join(out, "collisions_known_answers.json"), "w"), indent=1)
    print("  collisions_known_answers.json: %d cases" % len(cases))


def action_tables(out):
    from rl.policy.cadrl import CADRL
    res = {}
    for kin in ("holonomic", "unicycle"):
        for vp in (0.6, 0.7, 1.0):
            p = CADRL()
            p.kinematics = kin
            p.speed_samples = 5
            p.rotation_samples = 16
            p.build_action_space(vp)
            res["%s_%g" % (kin, vp)] = np.array([list(a) for a in p.action_space], dtype=np.float64)
    np.savez_compressed(os.path.join(out, "action_tables.npz"), **res)
    print("  action_tables.npz")


def scenes(out):
    """Generated scenes for a few seeds (host scene generator)."""
    res = {}
    for tag, cfgp, over in [
        ("adults5", "configs/test_configs/test_env_configs/env_adults_5.config", None),
        ("ebcadrl", "data/eb-cadrl/adults_8_bikes_8_child_8_static_3_35_sec_new_reward_fix_static.config", None),
    ]:
        env, policy, robot, ec, pc = make(cfgp, None, None, "linear", env_overrides=over)
        for seed in (1002, 1003, 1000000):
            env.reset("test", scene_number=seed)
            s = snapshot(env)
            for k in ("hum_pv", "hum_gr", "hum_type", "stat"):
                res["%s_%d_%s" % (tag, seed, k)] = s[k]
            res["%s_%d_map_zero" % (tag, seed)] = np.packbits((env.scene.map == 0).astype(np.uint8), axis=None)
            res["%s_%d_obstacles" % (tag, seed)] = np.array(
                [[o.location_x, o.location_y, o.dim[0], o.dim[1]] for o in env.scene.obstacles], dtype=np.int64).reshape(-1, 4)
    np.savez_compressed(os.path.join(out, "scenes.npz"), **res)
    print("  scenes.npz")


def copy_configs(out):
    """INI fixtures: the reference's own config files (data, key names verbatim) that the tests configure the
    env / policy / trainer with on the GPU box, where /root/reference does not exist."""
    import shutil
    dst = os.path.join(out, "configs")
    os.makedirs(dst, exist_ok=True)
    for src, name in [
        ("configs/test_configs/test_env_configs/env_adults_5.config", "env_adults_5.config"),
        ("configs/test_configs/test_env_configs/env_adults_3_bikes_3_static_2.config", "env_adults_3_bikes_3_static_2.config"),
        ("configs/test_configs/test_env_configs/env_adults_5_bikes_5_static_5.config", "env_adults_5_bikes_5_static_5.config"),
        ("configs/test_configs/test_env_configs/env_adults_3_bikes_3_child_3_static_3_fast_train.config", "env_fast_train.config"),
        ("configs/test_configs/test_policy_configs/policy.config", "policy.config"),
        ("configs/test_configs/test_train_configs/test_train.config", "test_train.config"),
        ("data/eb-cadrl/adults_8_bikes_8_child_8_static_3_35_sec_new_reward_fix_static.config", "env_ebcadrl.config"),
        ("data/eb-cadrl/policy_x2_agent_type.config", "policy_ebcadrl.config"),
        ("data/eb-cadrl/train_50k_8x.config", "train_50k_8x.config"),
    ]:
        shutil.copy(os.path.join(REF, src), os.path.join(dst, name))
    print("  configs/: 9 INI fixtures")
    # the reference's fixed test scenes (tests/test_scenes/test_collisions/*.json): golden inputs of its
    # episode-outcome tests, replayed through our env API by tests/test_gpu_mirror.py
    sdst = os.path.join(out, "scenes")
    os.makedirs(sdst, exist_ok=True)
    src_dir = os.path.join(REF, "tests", "test_scenes", "test_collisions")
    for f in sorted(os.listdir(src_dir)):
        shutil.copy(os.path.join(src_dir, f), os.path.join(sdst, f))
    for cfgname in ("env_adults_5_bikes_0_static_5.config", "env_adults_5_child_5_static_5.config"):
        shutil.copy(os.path.join(REF, "configs/test_configs/test_env_configs", cfgname), os.path.join(dst, cfgname))
    print("  scenes/: %d JSON scenes" % len(os.listdir(sdst)))


def update_memory_golden(out):
    """The reference's own Explorer.update_memory / ParallelExplorer.update_memory (rl/utils/explorer.py:151-200,
    parallel_explorer.py:286-330) on trajectories it produced itself: (state, value) pairs of one RL episode (TD
    target through a target network) and one imitation-learning episode (ORCA robot, discounted returns), plus
    which episodes the sequential explorer lets into the memory (explorer.py:82-92)."""
    from rl.utils.explorer import Explorer
    from rl.utils.memory import ReplayMemory
    from rl.utils.parallel_explorer import ParallelExplorer
    POL = "configs/test_configs/test_policy_configs/policy.config"
    ENVC = "configs/test_configs/test_env_configs/env_adults_5.config"
    res = {}
    # ---- RL episode: SARL robot, train phase, epsilon 0, target network = the same weights ----
    env, policy, robot, ec, pc = make(ENVC, POL, "model_weights/sarl_model_baseline.pth", "sarl", phase="train")
    policy.set_epsilon(0.0)
    env.get_local_map_angular = lambda *a, **k: None
    captured = {}
    for tag, store_il in (("rl", False),):
        mem = ReplayMemory(100000)
        ex = Explorer(env, robot, "cpu", mem, policy.gamma, target_policy=policy)
        ex.update_target_model(policy.get_model())
        orig = ex.update_memory

        def spy(states, actions, rewards, imitation_learning=False, _orig=orig):
            captured["states"], captured["rewards"] = states, rewards
            return _orig(states, actions, rewards, imitation_learning)

        ex.update_memory = spy
        np.random.seed(0)
        env.scene.case_counter["train"] = 3
        ex.run_k_episodes(1, "train", update_memory=True, imitation_learning=False)
        states, rewards = captured["states"], captured["rewards"]
        res["rl_states"] = torch.stack(states).numpy().astype(np.float32)
        res["rl_rewards"] = np.array(rewards, dtype=np.float64)
        res["rl_values"] = np.array([float(v) for _, v in mem.memory], dtype=np.float32)
        res["rl_mem_states"] = torch.stack([s for s, _ in mem.memory]).numpy().astype(np.float32)
        # the parallel explorer's twin must give the same pairs
        mem2 = ReplayMemory(100000)
        pex = ParallelExplorer(env, robot, "cpu", 1, mem2, policy.gamma, target_policy=policy)
        pex.update_target_model(policy.get_model())
        pex.update_memory(states, None, rewards, False)
        assert np.array_equal(res["rl_values"], np.array([float(v) for _, v in mem2.memory], dtype=np.float32))
        # and its IL branch on the same rewards (states pass through transform there; values depend on rewards only)
        gamma, dt, vp = policy.gamma, robot.time_step, robot.v_pref
        res["il_values_of_rl_rewards"] = np.array(
            [sum(pow(gamma, max(t - i, 0) * dt * vp) * r * (1 if t >= i else 0) for t, r in enumerate(rewards))
             for i in range(len(rewards))], dtype=np.float32)
        res["gamma_dt_vpref"] = np.array([gamma, dt, vp], dtype=np.float64)
        res["cumulative_reward"] = np.array([ex.calculate_cumulative_reward(rewards)], dtype=np.float64)
    # ---- IL episode: ORCA robot (rl/train.py:99-143), memory values = discounted returns, states = transform ----
    env2, orca, robot2, ec2, _ = make(ENVC, None, None, "orca", phase="train")
    orca.safety_space = 0.15
    orca.multiagent_training = policy.multiagent_training
    robot2.set_policy(orca)
    mem = ReplayMemory(100000)
    ex = Explorer(env2, robot2, "cpu", mem, policy.gamma, target_policy=policy)
    cap2 = {}
    orig2 = ex.update_memory

    def spy2(states, actions, rewards, imitation_learning=False):
        cap2["rewards"] = rewards
        return orig2(states, actions, rewards, imitation_learning)

    ex.update_memory = spy2
    tried = []
    for case in range(5, 25):                     # until one episode is let into the memory (explorer.py:82-92)
        env2.scene.case_counter["train"] = case
        before = len(mem.memory)
        ex.run_k_episodes(1, "train", update_memory=True, imitation_learning=True)
        ended = "ReachGoal" if ex.success else ("Timeout" if ex.timeout else "Collision")
        tried.append((case, ended, len(mem.memory) - before))
        if len(mem.memory) > before:
            break
    res["il_tried_json"] = np.frombuffer(json.dumps(tried).encode(), dtype=np.uint8)
    print("  IL episodes tried:", tried)
    res["il_rewards"] = np.array(cap2.get("rewards", []), dtype=np.float64)
    res["il_values"] = np.array([float(v) for _, v in mem.memory], dtype=np.float32)
    res["il_states"] = (torch.stack([s for s, _ in mem.memory]).numpy().astype(np.float32) if len(mem.memory)
                        else np.zeros((0, 5, 13), np.float32))
    np.savez_compressed(os.path.join(out, "update_memory.npz"), **res)
    print("  update_memory.npz: rl T=%d, il T=%d" % (len(res["rl_rewards"]), len(res["il_rewards"])))


def static_decomposition(out):
    """Walls -> occupancy grid (scene_generator.py:888-922, incl. the per-cell edge path) and -> static discs
    (:380-422) by the REFERENCE's own code, for walls given as (grid location, dimensions) the way generate_wall
    returns them (:194-243).  Pins the wall geometry of the synthetic / device scene generator."""
    class WallCfg(object):      # what generate_static_map_input / generate_wall read from a scene config
        def __init__(self, walls):
            self.walls = walls

        def getint(self, sec, key):
            return {"num_circles": 0, "num_walls": len(self.walls)}[key]

        def getfloat(self, sec, key):
            w = self.walls[int(key)]
            return {"x_locations_walls": w[0], "y_locations_walls": w[1], "x_dim": w[2], "y_dim": w[3]}[sec]

    res = {}
    rng = np.random.default_rng(20261018)
    k = 0
    for cfgp, size in (("configs/test_configs/test_env_configs/env_adults_3_bikes_3_static_2.config", None),
                       ("data/eb-cadrl/adults_8_bikes_8_child_8_static_3_35_sec_new_reward_fix_static.config", None)):
        env, policy, robot, ec, pc = make(cfgp, None, None, "linear")
        sg = env.scene
        robot.set(0, -sg.circle_radius, 0, sg.circle_radius, 0, 0, np.pi / 2)
        G = int(round(sg.map_size_m / sg.map_resolution))
        for scene in range(40):
            walls = []
            while len(walls) < 4:
                lx, ly = int(rng.integers(-G // 2, G // 2)), int(rng.integers(-G // 2, G // 2))
                length = int(rng.integers(1, 6))
                xd, yd = (length, 1) if rng.random() > 0.5 else (1, length)
                xm, ym = lx * sg.map_resolution, ly * sg.map_resolution
                clear = robot.radius + sg.discomfort_dist
                near = any(abs(xm - 0.0) < xd / 2.0 + clear and abs(ym - gy) < yd / 2.0 + clear
                           for gy in (-sg.circle_radius, sg.circle_radius))
                if not near:          # (generate_wall would spin on a fixed wall that touches the robot's start / goal)
                    walls.append((lx, ly, xd, yd))
            sg.generate_static_map_input(sg.map_size_m, "test", config=WallCfg(walls))
            res["s%03d_walls" % k] = np.array(walls, dtype=np.int64)
            res["s%03d_map" % k] = np.array([sg.map_size_m, sg.map_resolution])
            res["s%03d_zero" % k] = np.packbits((sg.map == 0).astype(np.uint8), axis=None)
            res["s%03d_discs" % k] = np.array([[s.px, s.py, s.radius] for s in sg.static_obstacles_as_pedestrians],
                                              dtype=np.float64).reshape(-1, 3)
            res["s%03d_vertices" % k] = np.array(sg.obstacle_vertices, dtype=np.float64).reshape(-1, 4, 2)
            k += 1
    res["n_scenes"] = np.array([k])
    np.savez_compressed(os.path.join(out, "static_decomposition.npz"), **res)
    print("  static_decomposition.npz: %d scenes x 4 walls" % k)


def round2(out):
    """Fixtures added in round 2 (the earlier ones are left untouched)."""
    POL = "configs/test_configs/test_policy_configs/policy.config"
    BASE_W = "model_weights/sarl_model_baseline.pth"
    EB_ENV = "data/eb-cadrl/adults_8_bikes_8_child_8_static_3_35_sec_new_reward_fix_static.config"
    EB_POL = "data/eb-cadrl/policy_x2_agent_type.config"
    EB_W = "data/eb-cadrl/rl_model_val.pth"
    # humans driven by the `linear` policy (simulator/policy/linear.py:17-23): all of them, and mixed with ORCA
    e = make("configs/test_configs/test_env_configs/env_adults_5.config", POL, BASE_W, "sarl",
             env_overrides={("adults", "policy"): "linear"})
    trace("trace_linear_adults5_seed1006", *e, {"test_case": 6}, 40, {0, 7}, out)
    e = make("configs/test_configs/test_env_configs/env_adults_3_bikes_3_static_2.config", POL, BASE_W, "sarl",
             env_overrides={("bicycles", "policy"): "linear"})
    trace("trace_mixed_linear_bikes_seed1007", *e, {"test_case": 7}, 40, {0, 9}, out)
    # train phase (epsilon 0): records policy.last_state = transform(current state) (multi_human_rl.py:84-85,128-149)
    ov = {("sim", "adult_num"): 4, ("sim", "bicycle_num"): 3, ("sim", "children_num"): 3}
    e = make(EB_ENV, EB_POL, EB_W, "sarl", phase="train", env_overrides=ov)
    e[1].set_epsilon(0.0)
    trace("trace_train_cfg2_h10_seed13", *e, {"scene_number": 13}, 24, {0}, out)
    e = make("configs/test_configs/test_env_configs/env_adults_5.config", POL, BASE_W, "sarl", phase="train",
             policy_overrides={("action_space", "kinematics"): "unicycle"},
             env_overrides={("reward", "rotation_penalty_factor"): -0.01})
    e[1].set_epsilon(0.0)
    trace("trace_train_unicycle_adults5_seed1008", *e, {"test_case": 8}, 24, {0}, out)
    update_memory_golden(out)
    static_decomposition(out)


def main():
    out = HERE
    if "--configs-only" in sys.argv:
        copy_configs(out)
        return
    if "--round2" in sys.argv:
        round2(out)
        return
    POL = "configs/test_configs/test_policy_configs/policy.config"
    BASE_W = "model_weights/sarl_model_baseline.pth"
    EB_ENV = "data/eb-cadrl/adults_8_bikes_8_child_8_static_3_35_sec_new_reward_fix_static.config"
    EB_POL = "data/eb-cadrl/policy_x2_agent_type.config"
    EB_W = "data/eb-cadrl/rl_model_val.pth"
    print("writing goldens to", out)
    collisions_known_answers(out)
    action_tables(out)
    save_weights(BASE_W, "weights_sarl_baseline.npz", out)
    save_weights(EB_W, "weights_ebcadrl.npz", out)

    # cfg1: the reference's own CPU-runnable case (BASELINE.json configs[0]); full episode
    e = make("configs/test_configs/test_env_configs/env_adults_5.config", POL, BASE_W, "sarl")
    trace("trace_cfg1_adults5_seed1002", *e, {"test_case": 2}, 400, {0, 10, 40}, out)
    # with walls + bicycles (static discs in the observation, obstacle grid test live)
    e = make("configs/test_configs/test_env_configs/env_adults_3_bikes_3_static_2.config", POL, BASE_W, "sarl")
    trace("trace_adults3_bikes3_static2_seed1002", *e, {"test_case": 2}, 400, {0, 20}, out)
    # EB-CADRL shipped config: 24 humans (maxNeighbors truncation live), 3 walls, D = 17
    e = make(EB_ENV, EB_POL, EB_W, "sarl")
    trace("trace_ebcadrl_h24_seed1000000", *e, {"scene_number": 1000000}, 16, {0, 8}, out)
    # cfg2 shape: 4 adults + 3 bicycles + 3 children + 3 walls, D = 17
    ov = {("sim", "adult_num"): 4, ("sim", "bicycle_num"): 3, ("sim", "children_num"): 3}
    e = make(EB_ENV, EB_POL, EB_W, "sarl", env_overrides=ov)
    trace("trace_cfg2_h10_seed7", *e, {"scene_number": 7}, 400, {0, 12}, out)
    e = make(EB_ENV, EB_POL, EB_W, "sarl", env_overrides=ov)
    trace("trace_cfg2_h10_seed11", *e, {"scene_number": 11}, 400, {3}, out)
    # unicycle robot, kinematics string "unicycle" (theta_rel live) and "nonholonomic" (theta_rel = 0)
    for kin in ("unicycle", "nonholonomic"):
        e = make("configs/test_configs/test_env_configs/env_adults_5.config", POL, BASE_W, "sarl",
                 policy_overrides={("action_space", "kinematics"): kin},
                 env_overrides={("reward", "rotation_penalty_factor"): -0.01})
        trace("trace_%s_adults5_seed1003" % kin, *e, {"test_case": 3}, 60, {0, 5}, out)
    # robot visible to the humans (env.py:401-402)
    e = make("configs/test_configs/test_env_configs/env_adults_5.config", POL, BASE_W, "sarl",
             env_overrides={("robot", "visible"): "true"})
    trace("trace_visible_adults5_seed1004", *e, {"test_case": 4}, 40, {0}, out)
    # linear robot on the reference's fixed collision scenes (tests/test_collisions_simulation.py:12-39)
    scenes_dir = os.path.join(REF, "tests", "test_scenes", "test_collisions")
    for cfgp, names in [
        ("configs/test_configs/test_env_configs/env_adults_5_bikes_5_static_5.config",
         ["collision_with_adult", "collision_with_bicycle", "collision_with_static", "no_collisions"]),
        ("configs/test_configs/test_env_configs/env_adults_5_bikes_0_static_5.config",
         ["bikes_0_collision_with_adult_1", "bikes_0_collision_with_adult_2", "bikes_0_no_collisions"]),
        ("configs/test_configs/test_env_configs/env_adults_5_child_5_static_5.config", ["collision_with_child"]),
    ]:
        for nm in names:
            e = make(cfgp, None, None, "linear")
            trace("scene_" + nm, *e, {"load_scene_path": os.path.join(scenes_dir, nm + ".json")}, 400, set(), out)
    # ORCA robot (imitation learning, rl/train.py:99-143): safety_space 0.15 since the robot is invisible
    e = make(EB_ENV, None, None, "orca", env_overrides=ov)
    e[1].safety_space = 0.15
    trace("trace_orca_robot_h10_seed5", *e, {"scene_number": 5}, 60, set(), out)
    scenes(out)
    copy_configs(out)
    round2(out)


if __name__ == "__main__":
    main()
