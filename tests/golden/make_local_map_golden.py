#!/usr/bin/env python
"""Golden vectors for the angular local map, produced by the UNMODIFIED reference.

Runs only in the build container (needs /root/reference); writes tests/golden/local_map_angular.npz.

    python tests/golden/make_local_map_golden.py      (re-executes itself with the reference on PYTHONPATH)

For several generated scenes (walls, circular obstacles, none) and robot poses (fp32-representable px, py, radius,
theta, since the hot path's state is fp32) it records scene.obstacle_vertices and the 48-sector vector returned by
EntityBasedCollisionAvoidance.get_local_map_angular (simulator/env.py:570-628), normalised and raw.
"""
import configparser
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
import gym  # noqa: E402
from simulator.agents.robot import Robot  # noqa: E402
from simulator.utils.state import FullState  # noqa: E402
from simulator.policy.policy_factory import policy_factory  # noqa: E402

CFG = os.path.join(HERE, "configs")
CASES = [  # (config, overrides, seeds)
    ("env_ebcadrl.config", {}, [3, 4, 5, 6]),                                            # 3 walls
    ("env_ebcadrl.config", {("map", "num_walls"): 5, ("map", "num_circles"): 2}, [7, 8]),  # walls + circular obstacles
    ("env_ebcadrl.config", {("map", "angular_map_dim"): 72, ("map", "angular_map_max_range"): 4.5,
                            ("map", "angle_min"): -0.5, ("map", "angle_max"): 0.5}, [9, 10]),   # a forward fan
    ("env_adults_5.config", {}, [11]),                                                    # no obstacles at all
]
POSES = 24


def main():
    rng = np.random.RandomState(20261018)
    out = {}
    n = 0
    for cfg_name, overrides, seeds in CASES:
        cp = configparser.RawConfigParser()
        cp.read(os.path.join(CFG, cfg_name))
        for (sec, key), val in overrides.items():
            cp.set(sec, key, str(val))
        env = gym.make("EntityBasedCollisionAvoidance-v0")
        env.configure(cp)
        robot = Robot(cp, "robot")
        robot.set_policy(policy_factory["linear"]())
        env.set_robot(robot)
        for seed in seeds:
            env.reset(phase="test", test_case=seed, compute_local_map=False)
            verts = np.array(env.scene.obstacle_vertices, dtype=np.float64).reshape(-1, 4, 2)
            half = env.scene.map_size_m / 2.0
            poses = np.zeros((POSES, 4), np.float64)
            raw = np.zeros((POSES, env.angular_map_dim), np.float64)
            nrm = np.zeros_like(raw)
            for k in range(POSES):
                px, py = rng.uniform(-half, half, 2)
                if k % 6 == 5 and len(verts):      # a pose hugging an obstacle corner (short distances, wrapped sweeps)
                    v = verts[rng.randint(len(verts)), rng.randint(4)]
                    px, py = v[0] + rng.uniform(-0.6, 0.6), v[1] + rng.uniform(-0.6, 0.6)
                radius = rng.uniform(0.2, 0.5)
                theta = rng.uniform(-np.pi, np.pi) if k % 4 else [0.0, np.pi / 2, -np.pi / 2, np.pi][(k // 4) % 4]
                px, py, radius, theta = (float(np.float32(x)) for x in (px, py, radius, theta))
                ob = FullState(px, py, 0.0, 0.0, radius, 0.0, 0.0, 1.0, theta)
                nrm[k] = env.get_local_map_angular(ob, normalize=True, append=False)
                raw[k] = env.get_local_map_angular(ob, normalize=False, append=False)
                poses[k] = (px, py, radius, theta)
            tag = "c%d" % n
            out[tag + "_params"] = np.array([env.angular_map_max_range, env.angular_map_min_angle,
                                             env.angular_map_max_angle, env.angular_map_dim], np.float64)
            out[tag + "_verts"] = verts
            out[tag + "_poses"] = poses
            out[tag + "_raw"] = raw
            out[tag + "_norm"] = nrm
            print(tag, cfg_name, "seed", seed, "obstacles", len(verts), "min", raw.min(), "touched sectors",
                  int((raw < env.angular_map_max_range).sum()))
            n += 1
    out["n_cases"] = np.array([n])
    np.savez_compressed(os.path.join(HERE, "local_map_angular.npz"), **out)


if __name__ == "__main__":
    main()
