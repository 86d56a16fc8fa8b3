"""Explorer (rl/utils/explorer.py:20-340, rl/utils/parallel_explorer.py:148-338): run k episodes, keep the
reference's statistics and log lines, fill the replay memory.

The reference runs one episode at a time (or 8 in a process pool); here the k episodes are one batch on the
device (ebc.batched_env.BatchedEnv.run_episodes).  Memory semantics are kept: the sequential Explorer stores
only success / collision episodes (explorer.py:82-92), the parallel one stores all of them
(parallel_explorer.py:185-192) -> `store_all`; imitation learning uses discounted returns, RL uses the TD
target r + gamma^(dt v_pref) V_target(s') with the terminal state's value = its reward (explorer.py:151-200).
With torch.distributed initialised, every rank runs its slice of the seeds and the outcome counters are
exchanged with one fixed-size all_gather before logging (the only collective of evaluation)."""
import copy
import logging

import numpy as np
import torch

from ebc import abi

_COLLISIONS = (abi.EV_COLLISION_ADULT, abi.EV_COLLISION_BICYCLE, abi.EV_COLLISION_CHILD, abi.EV_COLLISION_OBSTACLE)


class Explorer(object):
    PHASES = ["val", "test"]
    ORCA_POLICY = "ORCA"

    def __init__(self, env, robot=None, device="cuda:0", memory=None, gamma=None, target_policy=None):
        self.env = env                      # ebc.batched_env.BatchedEnv
        self.robot = robot if robot is not None else env.robot
        self.device = device
        self.memory = memory
        self.gamma = gamma
        self.target_policy = target_policy
        self.target_model = None
        self.metrics = {}

    def update_target_model(self, target_model):
        """explorer.py:29-30.  The copy is what TD targets are computed with; on a CUDA env its weights are also
        handed to a second K4 handle (`env.target_value`), so the target network runs on the tensor-core kernels."""
        self.target_model = copy.deepcopy(target_model)
        tv = getattr(self.env, "target_value", None)
        if tv is not None:
            tv.set_model(self.target_model)

    def _target_values(self, states, rows):
        """V_target(states) -> float64 [B]; states [B, n, D] fp32, rows [B] real row counts."""
        tv = getattr(self.env, "target_value", None)
        if tv is not None and states.is_cuda:
            return tv(states, rows).double()
        model = self.target_model.to(states.device)
        return model(states, rows).reshape(-1).double()

    # ---- statistics (explorer.py:96-126, 214-340) ----------------------------------------------------------
    # One fixed-size vector per rank: 8 event counts, sum of success times, sum of returns, danger steps,
    # sum of min distances in danger, step count, episode count.  Ranks exchange it with ONE all_gather of 14 float64.
    SUMMARY = 14

    @staticmethod
    def summarise(arrays):
        ev = np.asarray(arrays["event"])
        v = np.zeros(Explorer.SUMMARY, np.float64)
        for code in range(8):
            v[code] = float((ev == code).sum())
        v[8] = float(np.asarray(arrays["time"])[ev == abi.EV_REACH_GOAL].sum())
        v[9] = float(np.asarray(arrays["cum_reward"]).sum())
        v[10] = float(np.asarray(arrays["too_close"]).sum())
        v[11] = float(np.asarray(arrays["min_dist_sum"]).sum())
        v[12] = float(np.asarray(arrays["steps"]).sum())
        v[13] = float(len(ev))
        return v

    def _gather(self, stats):
        """Rank-local per-episode arrays -> the job-wide summary vector (sum over ranks of `summarise`)."""
        import torch.distributed as dist
        arrays = stats if isinstance(stats, dict) else {k: np.asarray(getattr(stats, k)) for k in (
            "event", "time", "steps", "cum_reward", "too_close", "min_dist_sum")}
        v = self.summarise(arrays)
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dev = self.device if dist.get_backend() == "nccl" else "cpu"
            mine = torch.as_tensor(v, dtype=torch.float64, device=dev)
            out = torch.empty(dist.get_world_size(), self.SUMMARY, dtype=torch.float64, device=dev)
            dist.all_gather_into_tensor(out, mine.reshape(1, -1))
            v = out.sum(0).cpu().numpy()
        return v

    def log_results(self, summary, phase, episode=None, seeds=None, print_failure=False, events=None):
        """`summary`: the vector of `_gather` (or a dict of per-episode arrays, summarised here without a collective)."""
        if isinstance(summary, dict):
            events = summary["event"] if events is None else events
            summary = self.summarise(summary)
        v, n = summary, max(float(summary[13]), 1.0)
        rate = lambda code: float(v[code]) / n  # noqa: E731
        m = {
            "success_rate": rate(abi.EV_REACH_GOAL), "collision_rate": 0.0,
            "collision_rate_adult": rate(abi.EV_COLLISION_ADULT), "collision_rate_bicycle": rate(abi.EV_COLLISION_BICYCLE),
            "collision_rate_child": rate(abi.EV_COLLISION_CHILD), "collision_rate_obstacle": rate(abi.EV_COLLISION_OBSTACLE),
            "timeout_rate": rate(abi.EV_TIMEOUT),
            "avg_nav_time": float(v[8] / v[abi.EV_REACH_GOAL]) if v[abi.EV_REACH_GOAL] else float(self.env.time_limit),
            "total_reward": float(v[9] / n),
        }
        extra = "" if episode is None else "in episode {} ".format(episode)
        logging.info("{:<5} {}has success rate: {:.2f}, self.collision rate: {:.2f}, nav time: {:.2f}, total reward: {:.4f}".format(
            phase.upper(), extra, m["success_rate"], m["collision_rate"], m["avg_nav_time"], m["total_reward"]))
        logging.info("{:<5} {}self.collision rate adult: {:.2f}, self.collision rate bicycle: {:.2f},"
                     "self.collision rate child: {:.2f}, self.collision rate obstacle: {:.4f}".format(
                         phase.upper(), extra, m["collision_rate_adult"], m["collision_rate_bicycle"],
                         m["collision_rate_child"], m["collision_rate_obstacle"]))
        if phase in self.PHASES:
            num_step, close = int(v[12]), int(v[10])
            m["danger_frequency"] = close / max(num_step, 1)
            m["avg_min_dist"] = float(v[11] / close) if close else 0.0
            logging.info("Frequency of being in danger: {:.2f} and average min separate distance in danger: {:.2f}".format(
                m["danger_frequency"], m["avg_min_dist"]))
        if print_failure and seeds is not None and events is not None:      # this rank's own cases
            for name, code in (("Collision adult", abi.EV_COLLISION_ADULT), ("Collision bicycle", abi.EV_COLLISION_BICYCLE),
                               ("Collision child", abi.EV_COLLISION_CHILD), ("Collision obstacle", abi.EV_COLLISION_OBSTACLE),
                               ("Timeout", abi.EV_TIMEOUT)):
                cases = [int(s) for s, e in zip(seeds, events) if e == code]
                logging.info("%s cases: %s", name, " ".join(str(c) for c in cases))
        self.metrics = m
        return m

    # ---- episodes ----------------------------------------------------------------------------------------
    def run_k_episodes(self, num_episodes, phase, update_memory=False, imitation_learning=False, episode=None,
                       print_failure=False, seeds=None, epsilon=None, store_all=False, safety_space=0.0):
        env = self.env
        if seeds is None:
            off = env.COUNTER_OFFSET[phase]
            start = env.scene.case_counter[phase]
            seeds = [off + start + i for i in range(num_episodes)]
            env.scene.case_counter[phase] = (start + num_episodes) % env.scene.case_size[phase]
        all_stats, done_seeds = [], []
        for lo in range(0, len(seeds), env.N):
            chunk = list(seeds[lo:lo + env.N])
            pad = env.N - len(chunk)
            stats, traj = env.run_episodes(phase, chunk + chunk[:1] * pad, epsilon=epsilon, record=update_memory,
                                           imitation=imitation_learning, safety_space=safety_space)
            keep = slice(0, len(chunk))
            if update_memory:
                self.update_memory(traj, stats, imitation_learning, store_all, keep)
            for k in ("event", "time", "steps", "cum_reward", "too_close", "min_dist_sum"):
                setattr(stats, k, getattr(stats, k)[keep])
            all_stats.append(stats)
            done_seeds += chunk
        merged = type("S", (), {k: np.concatenate([getattr(s, k) for s in all_stats]) for k in (
            "event", "time", "steps", "cum_reward", "too_close", "min_dist_sum")})()
        summary = self._gather(merged)
        return self.log_results(summary, phase, episode, done_seeds, print_failure, events=merged.event)

    @torch.no_grad()
    def update_memory(self, traj, stats, imitation_learning=False, store_all=False, keep=slice(None)):
        if self.memory is None or self.gamma is None:
            raise ValueError("Memory or gamma value is not set!")
        states, rewards, alive, rows = traj["states"], traj["rewards"], traj["alive"], traj["rows"]
        T, N = rewards.shape
        ev = torch.as_tensor(stats.event, device=rewards.device)
        use = torch.zeros(N, dtype=torch.bool, device=rewards.device)
        use[keep] = True
        if not store_all:       # only success / collision episodes (explorer.py:82-92)
            ok = ev == abi.EV_REACH_GOAL
            for c in _COLLISIONS:
                ok |= ev == c
            use &= ok
        gamma_bar = self.gamma ** (self.env.time_step * self.env.robot.v_pref)
        if imitation_learning:
            # value_i = sum_{t >= i} gamma_bar^(t - i) r_t  (explorer.py:159-173), one backward scan
            values = torch.zeros_like(rewards)
            run = torch.zeros(N, dtype=rewards.dtype, device=rewards.device)
            for t in range(T - 1, -1, -1):
                run = torch.where(alive[t], rewards[t] + gamma_bar * run, run)
                values[t] = run
        else:
            # value_i = r_i + gamma_bar V_target(s_{i+1}); terminal: r_i  (explorer.py:174-186,
            # parallel_explorer.py:313-321).  Every next state of the batch in ONE target-network evaluation.
            values = rewards.clone()
            if T > 1:
                nxt = alive[1:]
                if bool(nxt.any()):
                    t_idx, e_idx = nxt.nonzero(as_tuple=True)
                    v = self._target_values(states[1:][t_idx, e_idx].contiguous(), rows[e_idx].contiguous())
                    values[:-1][t_idx, e_idx] += gamma_bar * v
        sel = alive & use[None]
        t_idx, e_idx = sel.nonzero(as_tuple=True)
        if len(t_idx):
            self.memory.push_batch(states[t_idx, e_idx], rows[e_idx], values[t_idx, e_idx].float())
