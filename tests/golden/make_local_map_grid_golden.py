#!/usr/bin/env python
"""Golden vectors for the binary grid sub-map ([map] use_grid_map = true), produced by the UNMODIFIED
reference and the container's OpenCV (cv2 4.13).

Runs only in the build container (needs /root/reference); writes tests/golden/local_map_grid.npz.

    python tests/golden/make_local_map_grid_golden.py      (re-executes itself with the reference on PYTHONPATH)

For generated scenes with walls it records scene.map and, for robot poses with fp32-representable px, py, theta (the
hot path's state is fp32) the array returned by EntityBasedCollisionAvoidance.get_local_map
(simulator/env.py:630-708: window of the occupancy grid around the robot, cv2.warpAffine rotation, threshold).
Poses cover the interior, windows clipped at each map border, windows that miss the map, and the axis-aligned headings.
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
    ("env_ebcadrl.config", {("map", "submap_size_m"): 5}, [3, 4]),                                  # 3 walls, 50 x 50
    ("env_ebcadrl.config", {("map", "submap_size_m"): 6, ("map", "num_walls"): 6}, [5, 6]),           # 60 x 60
    ("env_ebcadrl.config", {("map", "submap_size_m"): 3.1, ("map", "num_walls"): 8}, [7]),            # odd size: 31 x 31
]
POSES = 40


def main():
    rng = np.random.RandomState(20261019)
    out = {}
    n = 0
    for cfg_name, overrides, seeds in CASES:
        cp = configparser.RawConfigParser()
        cp.read(os.path.join(CFG, cfg_name))
        cp.set("map", "use_grid_map", "true")
        for (sec, key), val in overrides.items():
            cp.set(sec, key, str(val))
        env = gym.make("EntityBasedCollisionAvoidance-v0")
        env.configure(cp)
        robot = Robot(cp, "robot")
        robot.set_policy(policy_factory["linear"]())
        env.set_robot(robot)
        for seed in seeds:
            env.reset(phase="test", test_case=seed, compute_local_map=False)
            half = env.scene.map_size_m / 2.0
            size = int(round(env.submap_size_m / env.scene.map_resolution))
            poses = np.zeros((POSES, 3), np.float64)
            maps = np.zeros((POSES, size, size), np.uint8)
            for k in range(POSES):
                px, py = rng.uniform(-half, half, 2)
                if k % 5 == 1:      # window clipped at a border of the map
                    side = rng.randint(4)
                    edge = rng.uniform(half - 1.5, half + 1.0) * (1 if side & 1 else -1)
                    px, py = (edge, py) if side & 2 else (px, edge)
                if k % 10 == 7:     # far outside: the window misses the map
                    px = (half + env.submap_size_m) * (1 if k % 20 == 7 else -1)
                theta = rng.uniform(-np.pi, np.pi) if k % 4 else [0.0, np.pi / 2, -np.pi / 2, np.pi][(k // 4) % 4]
                px, py, theta = (float(np.float32(x)) for x in (px, py, theta))
                ob = FullState(px, py, 0.0, 0.0, 0.3, 0.0, 0.0, 1.0, theta)
                g = env.get_local_map(ob, append=False)
                assert g.shape == (size, size) and set(np.unique(g)) <= {0.0, 1.0}
                maps[k] = g.astype(np.uint8)
                poses[k] = (px, py, theta)
            tag = "c%d" % n
            out[tag + "_params"] = np.array([env.submap_size_m, env.scene.map_size_m, env.scene.map_resolution], np.float64)
            out[tag + "_map"] = (np.asarray(env.scene.map) > 0).astype(np.uint8)
            out[tag + "_poses"] = poses
            out[tag + "_grid"] = maps
            print(tag, cfg_name, "seed", seed, "size", size, "zero cells in the map", int((env.scene.map == 0).sum()),
                  "zero cells over the poses", int((maps == 0).sum()), "all-ones maps", int((maps.min(axis=(1, 2)) == 1).sum()))
            n += 1
    out["n_cases"] = np.array([n])
    np.savez_compressed(os.path.join(HERE, "local_map_grid.npz"), **out)


if __name__ == "__main__":
    main()
