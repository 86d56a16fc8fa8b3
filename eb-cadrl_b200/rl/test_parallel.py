"""rl/test_parallel.py of the reference (:24-176): one CSV row per test episode, same columns, same flags.

The reference maps `run_episode` over a process pool (multiprocessing.Pool(cpu_count()), :165-173); here the
episodes [--start, --end) are batches of --batch episodes per GPU on the device, and with several ranks
(`python -m torch.distributed.run --nproc-per-node W -m rl.test_parallel ...`) every rank takes a contiguous
slice and writes `<csv>.rank<r>`; rank 0 concatenates them after a barrier.

Columns (test_parallel.py:112-129): episode, time, reward, success, collision, collision_child, collision_adult,
collision_bicycle, collision_obstacle, timeout, too_close, min_dist, dist_to_goal, dmin_adult, dmin_bicycle,
dmin_child -- the last three and min_dist are per-step lists, exactly as pandas writes the reference's lists."""
import configparser
import logging
import os
import sys
import time

import numpy as np
import torch

COLUMNS = ["episode", "time", "reward", "success", "collision", "collision_child", "collision_adult",
           "collision_bicycle", "collision_obstacle", "timeout", "too_close", "min_dist", "dist_to_goal",
           "dmin_adult", "dmin_bicycle", "dmin_child"]


def episode_rows(seeds, stats, cfg):
    """EpisodeStats (+ dmin trace) of one batch -> the reference's per-episode dicts."""
    from ebc import abi
    dd = (cfg.discomfort_dist_adult, cfg.discomfort_dist_bicycle, cfg.discomfort_dist_child)
    rows = []
    for i, seed in enumerate(seeds):
        ev = int(stats.event[i])
        alive = stats.alive_trace[:, i]
        dm = stats.dmin_trace[alive, i]                       # [T_i, 3] adult, bicycle, child
        # Danger.min_dist per danger step (reward.py:138-166): child > bicycle > adult; a step is a danger step
        # iff its event was Danger, i.e. it did not end the episode and some dmin was below its discomfort distance
        T = len(dm)
        min_dist = []
        for t in range(T):
            terminal = t == T - 1
            if terminal:
                continue      # the last step ended the episode: its Info is not Danger
            a, b, c = dm[t]
            if c < dd[2]:
                min_dist.append(float(c))
            elif b < dd[1]:
                min_dist.append(float(b))
            elif a < dd[0]:
                min_dist.append(float(a))
        assert len(min_dist) == int(stats.too_close[i]), (seed, len(min_dist), int(stats.too_close[i]))
        rows.append({
            "episode": int(seed), "time": float(stats.time[i]), "reward": float(stats.cum_reward[i]),
            "success": int(ev == abi.EV_REACH_GOAL), "collision": 0,
            "collision_child": int(ev == abi.EV_COLLISION_CHILD), "collision_adult": int(ev == abi.EV_COLLISION_ADULT),
            "collision_bicycle": int(ev == abi.EV_COLLISION_BICYCLE),
            "collision_obstacle": int(ev == abi.EV_COLLISION_OBSTACLE), "timeout": int(ev == abi.EV_TIMEOUT),
            "too_close": int(stats.too_close[i]), "min_dist": min_dist, "dist_to_goal": float(stats.dist_to_goal[i]),
            "dmin_adult": [float(x) for x in dm[:, 0]], "dmin_bicycle": [float(x) for x in dm[:, 1]],
            "dmin_child": [float(x) for x in dm[:, 2]],
        })
    return rows


def main(argv=None):
    import pandas as pd
    from ebc.batched_env import BatchedEnv
    from rl.policy.sarl import SARL
    from rl.test import parse_arguments
    args = parse_arguments(argv)
    if not args.csv or args.start is None or args.end is None:
        sys.exit(1)                                            # test_parallel.py:135-136
    logging.basicConfig(level=logging.INFO, format="%(asctime)s, %(levelname)s: %(message)s",
                        datefmt="%Y-%m-%d %H:%M:%S", stream=sys.stdout, force=True)
    model_path = None
    env_config_file, policy_config_file = args.env_config, args.policy_config
    if args.model_dir is not None:
        env_config_file = os.path.join(args.model_dir, os.path.basename(args.env_config))
        policy_config_file = os.path.join(args.model_dir, os.path.basename(args.policy_config))
        model_path = os.path.join(args.model_dir, args.model_name or "rl_model.pth")
    if args.model_path is not None:
        model_path = args.model_path
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    device = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)
    policy = SARL()
    pc = configparser.RawConfigParser()
    pc.read(policy_config_file)
    policy.configure(pc)
    if model_path is None:
        sys.exit(1)
    if model_path.endswith(".npz"):
        z = np.load(model_path)
        policy.get_model().load_state_dict({k: torch.as_tensor(z[k]) for k in z.files})
    else:
        policy.get_model().load_state_dict(torch.load(model_path, map_location="cpu"))
    policy.set_phase("test")
    policy.set_device(device)
    ec = configparser.RawConfigParser()
    ec.read(env_config_file)
    if args.square:
        ec.set("sim", "test_sim_adult", "square_crossing")
    if args.circle:
        ec.set("sim", "test_sim_adult", "circle_crossing")
    episodes = list(range(args.start, args.end))               # scene_number = episode (test_parallel.py:60)
    per = -(-len(episodes) // world)
    mine = episodes[rank * per:(rank + 1) * per]
    t0 = time.time()
    rows = []
    if mine:
        env = BatchedEnv(ec, policy, min(args.batch, len(mine)), device)
        for lo in range(0, len(mine), env.N):
            chunk = mine[lo:lo + env.N]
            stats, _ = env.run_episodes("test", chunk + chunk[:1] * (env.N - len(chunk)), record_dmin=True)
            rows += episode_rows(chunk, stats, env.cfg)
    df = pd.DataFrame(rows, columns=COLUMNS)
    if world > 1:
        df.to_csv("%s.rank%d" % (args.csv, rank))
        dist.barrier()
        if rank == 0:
            parts = [pd.read_csv("%s.rank%d" % (args.csv, r), index_col=0) for r in range(world)]
            pd.concat(parts, ignore_index=True).to_csv(args.csv)
            for r in range(world):
                os.remove("%s.rank%d" % (args.csv, r))
        dist.barrier()
        dist.destroy_process_group()
    else:
        df.to_csv(args.csv)
    logging.info("Time passed: %s", time.time() - t0)
    return df


if __name__ == "__main__":
    main()
