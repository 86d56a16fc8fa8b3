"""rl/test.py — evaluate a trained policy on the test seeds (rl/test.py:17-151 and rl/test_parallel.py of the
reference), same flags; the episodes are one batch per GPU, one CSV row per episode with --csv."""
import argparse
import configparser
import logging
import os
import sys

import numpy as np
import torch


def parse_arguments(argv=None):
    p = argparse.ArgumentParser("Parse configuration file")
    p.add_argument("--env_config", type=str, default="configs/env_configs/env.config")
    p.add_argument("--policy_config", type=str, default="configs/policy_configs/policy.config")
    p.add_argument("--policy", type=str, default="sarl")
    p.add_argument("--model_dir", type=str, default=None)
    p.add_argument("--model_path", type=str, default=None)
    p.add_argument("--il", default=False, action="store_true")
    p.add_argument("--gpu", default=False, action="store_true")
    p.add_argument("--phase", type=str, default="test")
    p.add_argument("--test_case", type=int, default=None)
    p.add_argument("--csv", type=str, default=None)
    p.add_argument("--model_name", type=str, default=None)
    p.add_argument("--square", default=False, action="store_true")
    p.add_argument("--circle", default=False, action="store_true")
    p.add_argument("--processes", type=int, default=None, help="accepted for compatibility (the episodes are a "
                   "device batch, not a process pool)")
    p.add_argument("--start", type=int, default=None)
    p.add_argument("--end", type=int, default=None)
    p.add_argument("--batch", type=int, default=256, help="episodes advanced at once per GPU")
    # rendering flags of the reference (rl/test.py:32,36-37): parsed so that its command lines keep working
    p.add_argument("--visualize", default=False, action="store_true")
    p.add_argument("--video_file", type=str, default=None)
    p.add_argument("--traj", default=False, action="store_true")
    return p.parse_args(argv)


def main(argv=None):
    from ebc.abi import EVENT_NAMES
    from ebc.batched_env import BatchedEnv
    from rl.policy.policy_factory import policy_factory
    from rl.utils.explorer import Explorer
    args = parse_arguments(argv)
    logging.basicConfig(level=logging.INFO, format="%(asctime)s, %(levelname)s: %(message)s", stream=sys.stdout, force=True)
    env_config_path, policy_config_path, model_path = args.env_config, args.policy_config, args.model_path
    if args.model_dir is not None:
        env_config_path = os.path.join(args.model_dir, os.path.basename(args.env_config))
        policy_config_path = os.path.join(args.model_dir, os.path.basename(args.policy_config))
        if args.il:
            model_path = os.path.join(args.model_dir, "il_model.pth")
        else:                                            # rl/test.py:59-64 preference order
            for name in ("resumed_rl_model.pth", "rl_model_val.pth", "rl_model.pth"):
                if os.path.exists(os.path.join(args.model_dir, name)):
                    model_path = os.path.join(args.model_dir, name)
                    break
    device = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    policy = policy_factory[args.policy]()
    pc = configparser.RawConfigParser()
    pc.read(policy_config_path)
    policy.configure(pc)
    if policy.trainable:
        if model_path is None:
            raise ValueError("Trainable policy must be specified with a model weights directory")
        if model_path.endswith(".npz"):
            z = np.load(model_path)
            policy.get_model().load_state_dict({k: torch.as_tensor(z[k]) for k in z.files})
        else:
            policy.get_model().load_state_dict(torch.load(model_path, map_location="cpu"))
    policy.set_phase(args.phase)
    policy.set_device(device)
    ec = configparser.RawConfigParser()
    ec.read(env_config_path)
    if args.start is not None and args.end is not None:
        seeds = list(range(args.start, args.end))        # test_parallel.py: scene_number = episode
    elif args.test_case is not None:
        seeds = [BatchedEnv.COUNTER_OFFSET[args.phase] + args.test_case]
    else:
        seeds = [BatchedEnv.COUNTER_OFFSET[args.phase] + i for i in range(ec.getint("env", "test_size"))]
    if args.visualize:
        raise NotImplementedError("--visualize: rendering (simulator/utils/render.py) is outside the hot path (SURVEY §2 #18); "
                                  "run without it for the per-episode statistics, --csv for one row per episode")
    if args.square:                                      # rl/test.py:100-103
        ec.set("sim", "test_sim_adult", "square_crossing")
    if args.circle:
        ec.set("sim", "test_sim_adult", "circle_crossing")
    env = BatchedEnv(ec, policy, min(args.batch, len(seeds)), device)
    explorer = Explorer(env, None, device, gamma=policy.gamma)
    rows = []
    for lo in range(0, len(seeds), env.N):
        chunk = seeds[lo:lo + env.N]
        stats, _ = env.run_episodes(args.phase, chunk + chunk[:1] * (env.N - len(chunk)))
        for i, s in enumerate(chunk):
            rows.append({"episode": s, "info": EVENT_NAMES[int(stats.event[i])], "time": float(stats.time[i]),
                         "steps": int(stats.steps[i]), "cumulative_reward": float(stats.cum_reward[i]),
                         "too_close": int(stats.too_close[i]), "min_dist_sum": float(stats.min_dist_sum[i])})
    arrays = {k: np.array([r[v] for r in rows]) for k, v in (("time", "time"), ("steps", "steps"),
              ("cum_reward", "cumulative_reward"), ("too_close", "too_close"), ("min_dist_sum", "min_dist_sum"))}
    arrays["event"] = np.array([EVENT_NAMES.index(r["info"]) for r in rows])
    metrics = explorer.log_results(arrays, args.phase, seeds=seeds, print_failure=True)
    if args.csv:
        import pandas as pd
        pd.DataFrame(rows).to_csv(args.csv, index=False)
    return metrics


if __name__ == "__main__":
    main()
