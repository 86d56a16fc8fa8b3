#!/usr/bin/env python
"""The UNMODIFIED Python reference timed in this build container: agent-steps/s of
`robot.act` (81 x onestep_lookahead + value network + argmax) + `env.step` per episode step, one process per core.

The reference cannot travel to the GPU box (no /root/reference there, no rvo2 wheel anywhere), so bench.py's CPU arm
is the C / OpenMP oracle port, which is far faster than the Python original.  This tool records the original's own
rate next to it so that the conservatism of the bench's GPU / CPU ratio is on record.

    python tools/python_reference_rate.py [--procs P] [--steps S]     -> profiles/r2_python_reference_cpu.json

`rvo2` is not installed in this image: ORCA runs through oracle/shims/rvo2.py (the C restatement behind the Python-RVO2
interface) -- the rest (env.step, collisions, reward, angular local map, lookahead, rotate, torch value network) is the
reference's own Python.  The real Python-RVO2 is C++ as well, so this does not flatter either side.
"""
import configparser
import json
import multiprocessing as mp
import os
import subprocess
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
REF = os.environ.get("EBC_REFERENCE", "/root/reference")

if os.environ.get("EBC_GOLDEN_CHILD") != "1":
    env = dict(os.environ)
    env["EBC_GOLDEN_CHILD"] = "1"
    env["OMP_NUM_THREADS"] = "1"
    env["PYTHONPATH"] = os.pathsep.join([os.path.join(REPO, "oracle", "shims"), REF, os.path.join(REF, "tests")])
    sys.exit(subprocess.call([sys.executable, os.path.abspath(__file__)] + sys.argv[1:], cwd=REF, env=env))

CASES = {
    # BASELINE configs[0]: SARL baseline, 5 ORCA humans, holonomic robot
    "cfg1_sarl_baseline_h5": dict(env="configs/test_configs/test_env_configs/env_adults_5.config",
                                  pol="configs/test_configs/test_policy_configs/policy.config",
                                  w="model_weights/sarl_model_baseline.pth", over={}),
    # BASELINE configs[1] shape: 4 adults + 3 bicycles + 3 children + 3 walls, entity-typed network
    "cfg2_h10_typed_3walls": dict(env="data/eb-cadrl/adults_8_bikes_8_child_8_static_3_35_sec_new_reward_fix_static.config",
                                  pol="data/eb-cadrl/policy_x2_agent_type.config", w="data/eb-cadrl/rl_model_val.pth",
                                  over={("sim", "adult_num"): 4, ("sim", "bicycle_num"): 3, ("sim", "children_num"): 3}),
    # the shipped EB-CADRL test configuration: 24 humans + 3 walls
    "ebcadrl_h24_3walls": dict(env="data/eb-cadrl/adults_8_bikes_8_child_8_static_3_35_sec_new_reward_fix_static.config",
                               pol="data/eb-cadrl/policy_x2_agent_type.config", w="data/eb-cadrl/rl_model_val.pth", over={}),
}


def worker(args):
    name, seed, steps = args
    import gym
    import torch
    from simulator.agents.robot import Robot
    from rl.policy.policy_factory import policy_factory
    torch.set_num_threads(1)
    c = CASES[name]
    ec = configparser.RawConfigParser()
    ec.read(c["env"])
    for (sec, key), val in c["over"].items():
        ec.set(sec, key, str(val))
    pc = configparser.RawConfigParser()
    pc.read(c["pol"])
    env = gym.make("EntityBasedCollisionAvoidance-v0")
    env.configure(ec)
    robot = Robot(ec, "robot")
    env.set_robot(robot)
    policy = policy_factory["sarl"]()
    policy.configure(pc)
    policy.get_model().load_state_dict(torch.load(c["w"], map_location="cpu"))
    robot.set_policy(policy)
    policy.set_phase("test")
    policy.set_device("cpu")
    policy.set_env(env) if hasattr(policy, "set_env") else None
    ob, local_map = env.reset(phase="test", test_case=seed)
    humans = len(env.scene.adults + env.scene.bicycles + env.scene.children)
    done_steps, t0 = 0, time.perf_counter()
    while done_steps < steps:
        action = robot.act(ob, local_map, env)
        ob, local_map, reward, done, info = env.step(action)
        done_steps += 1
        if done:
            seed += 1000
            ob, local_map = env.reset(phase="test", test_case=seed % 1000)
    dt = time.perf_counter() - t0
    return humans, done_steps, dt


def main():
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--procs", type=int, default=min(os.cpu_count() or 1, 8))
    ap.add_argument("--steps", type=int, default=12)
    a = ap.parse_args()
    out = {"what": "unmodified Python reference (robot.act + env.step per step), one process per core, this build container",
           "cpu_count": os.cpu_count(), "procs": a.procs, "steps_per_proc": a.steps, "cases": {}}
    for name in CASES:
        with mp.Pool(a.procs) as pool:
            t0 = time.perf_counter()
            res = pool.map(worker, [(name, 1 + i, a.steps) for i in range(a.procs)])
            wall = time.perf_counter() - t0
        H = res[0][0]
        per_proc = [(H + 1) * n / dt for _, n, dt in res]
        out["cases"][name] = {"humans": H, "agent_steps_per_sec_per_core": sum(per_proc) / len(per_proc),
                              "agent_steps_per_sec_all_procs": sum(per_proc), "env_steps_per_sec_per_core":
                              sum(n / dt for _, n, dt in res) / len(res), "wall_s_incl_setup": wall}
        print(name, json.dumps(out["cases"][name]), flush=True)
    path = os.path.join(REPO, "profiles", "r2_python_reference_cpu.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", path)


if __name__ == "__main__":
    main()
