"""BatchedEnv — N reference episodes at once: the batched face of simulator/env.py.

    env = BatchedEnv(env_config, policy, n_episodes, device)      # policy: rl.policy.sarl.SARL (configured)
    env.reset_batch("test", seeds)         # env.reset x N   (host scene generator replaying the reference RNG)
    env.lookahead_batch()                  # onestep_lookahead x 81 x N  -> la_reward / la_event / vin
    env.step_batch(action_idx=...)         # env.step x N
    stats = env.run_episodes("test", seeds)                        # the Explorer's inner loop, on the device

Episode r of rank k uses seed seeds[r]; sharding over ranks is the caller's slice of `seeds` (no collective).
"""
import configparser

import numpy as np
import torch

from . import abi
from .actions import build_action_space
from .config import SimConfig
from .engine import BatchedSim
from .scene import rects_from_zero_cells


class _RobotStub(object):
    """What the scene generator needs from the robot (position, goal, radius, policy flags)."""

    def __init__(self, config, policy):
        self.radius = config.getfloat("robot", "radius")
        self.v_pref = config.getfloat("robot", "v_pref")
        self.visible = config.getboolean("robot", "visible")
        self.policy = policy
        self.px = self.py = self.gx = self.gy = self.vx = self.vy = self.theta = 0.0

    def set(self, px, py, gx, gy, vx, vy, theta):
        self.px, self.py, self.gx, self.gy, self.vx, self.vy, self.theta = px, py, gx, gy, vx, vy, theta


class EpisodeStats(object):
    """Per-episode outcome arrays (what rl/utils/explorer.py:33-94 accumulates one episode at a time)."""

    def __init__(self, n):
        self.event = np.zeros(n, np.int64)          # terminal event code
        self.time = np.zeros(n, np.float64)         # env.global_time at the end
        self.steps = np.zeros(n, np.int64)
        self.cum_reward = np.zeros(n, np.float64)   # sum_t gamma^(t dt v_pref) r_t
        self.too_close = np.zeros(n, np.int64)      # Danger steps
        self.min_dist_sum = np.zeros(n, np.float64)


class BatchedEnv(object):
    COUNTER_OFFSET = {"train": 2000, "val": 0, "test": 1000}     # env.py:153-158

    def __init__(self, env_config, policy=None, n_episodes=1, device="cuda:0", max_humans=None, max_statics=None,
                 max_rects=None):
        from simulator.scene.scene_generator import SceneGenerator   # host scene generation (reference RNG)
        if not isinstance(env_config, configparser.RawConfigParser):
            cp = configparser.RawConfigParser()
            cp.read(env_config)
            env_config = cp
        self.config = env_config
        self.policy = policy
        self.cfg = SimConfig.from_ini(env_config)
        if policy is not None:
            self.cfg.robot_kinematics = policy.kinematics or "holonomic"
            self.cfg.with_agent_type = bool(getattr(policy, "with_agent_type", False))
            self.cfg.gamma = policy.gamma or self.cfg.gamma
        self.scene = SceneGenerator(env_config)
        self.robot = _RobotStub(env_config, policy if policy is not None else type("P", (), dict(
            multiagent_training=True, name="none", kinematics="holonomic"))())
        self.cfg.robot_visible = self.robot.visible
        self.scene.set_robot(self.robot)
        self.time_step, self.time_limit = self.cfg.time_step, self.cfg.time_limit
        sg = self.scene
        H = max_humans or max(sg.adult_num + sg.bicycle_num + sg.children_num, 20 if "mixed_20" in (
            sg.test_sim_adult, sg.train_val_sim_adult) else 1)
        walls, circles = sg.num_walls or 0, sg.num_circles or 0
        discs_per_wall = int(np.ceil(sg.max_wall_length / np.sqrt(2.0)))
        S = max_statics if max_statics is not None else walls * discs_per_wall + circles
        R = max_rects if max_rects is not None else 4 * (walls + circles)
        self.N = n_episodes
        n_actions = policy.n_actions if policy is not None and hasattr(policy, "n_actions") else 81
        self.sim = BatchedSim(self.cfg, n_episodes, H, S, R, n_actions, device=device)
        self.device = self.sim.device
        if policy is not None:
            if policy.action_space is None:
                policy.build_action_space(self.robot.v_pref)
            self.sim.set_actions(np.array([tuple(a)[:2] for a in policy.action_space], dtype=np.float64))
            self.sync_weights()
        self._weights_version = getattr(policy, "weights_version", 0)
        # epsilon-greedy exploration stream: ONE generator for the life of the env, keyed by (seed, rank), so that
        # neither successive run_episodes calls nor different ranks replay the same explore mask / random actions
        # (the reference draws from numpy's global stream, reseeded per scene: multi_human_rl.py:31)
        rank = int(__import__("os").environ.get("RANK", "0"))
        self._explore_rng = np.random.default_rng(np.random.SeedSequence([20261018, rank]))

    def seed_exploration(self, seed, rank=0):
        self._explore_rng = np.random.default_rng(np.random.SeedSequence([int(seed), int(rank)]))

    def sync_weights(self):
        model = self.policy.get_model()
        sd = {k: v.detach().cpu().numpy() for k, v in model.state_dict().items()}
        self.sim.set_weights(sd, with_global_state=model.with_global_state, self_state_dim=model.self_state_dim)
        if getattr(self.policy, "value_mode", None):
            self.sim.set_value_mode(self.policy.value_mode)
        self._weights_version = getattr(self.policy, "weights_version", 0)

    # ---- env.reset x N ---------------------------------------------------------------------------------
    def reset_batch(self, phase, seeds):
        """seeds[e] = scene_number of episode e (the reference: counter offset + case counter, or explicit)."""
        sim, sg = self.sim, self.scene
        assert len(seeds) == self.N
        H, S, R = sim.Hmax, max(sim.Smax, 1), max(sim.Rmax, 1)
        pv = np.zeros((self.N, H, 4), np.float32); gr = np.zeros((self.N, H, 4), np.float32)
        ty = np.zeros((self.N, H), np.uint8); hc = np.zeros(self.N, np.int32)
        sd = np.zeros((self.N, S, 4), np.float32); sc = np.zeros(self.N, np.int32)
        rc = np.zeros((self.N, R, 4), np.int16); rcc = np.zeros(self.N, np.int32)
        rad = sg.circle_radius
        for e, seed in enumerate(seeds):
            self.robot.set(0, -rad, 0, rad, 0, 0, np.pi / 2)
            sg.generate_random_scene(self.COUNTER_OFFSET, phase, scene_number=int(seed))
            hs = sg.adults + sg.bicycles + sg.children
            if len(hs) > sim.Hmax or len(sg.static_obstacles_as_pedestrians) > sim.Smax:
                raise abi.EbcError("scene %d exceeds the batch capacity (H=%d, S=%d)" % (
                    seed, len(hs), len(sg.static_obstacles_as_pedestrians)))
            for i, h in enumerate(hs):
                pv[e, i] = (h.px, h.py, h.vx, h.vy)
                gr[e, i] = (h.gx, h.gy, h.v_pref, h.radius)
                ty[e, i] = int(h.agent_type)
            hc[e] = len(hs)
            for i, s in enumerate(sg.static_obstacles_as_pedestrians):
                sd[e, i, :3] = (s.px, s.py, s.radius)
            sc[e] = len(sg.static_obstacles_as_pedestrians)
            rects = rects_from_zero_cells(sg.map == 0)
            if len(rects) > sim.Rmax:
                raise abi.EbcError("scene %d needs %d grid rectangles (capacity %d)" % (seed, len(rects), sim.Rmax))
            rc[e, :len(rects)] = rects
            rcc[e] = len(rects)
        rob_pv = np.tile(np.array([0, -rad, 0, 0], np.float32), (self.N, 1))
        rob_gr = np.tile(np.array([0, rad, self.robot.v_pref, self.robot.radius], np.float32), (self.N, 1))
        sim.load_episodes(0, pv, gr, ty, hc, sd, sc, rc, rcc, rob_pv, rob_gr,
                          np.full(self.N, np.pi / 2, np.float32), np.zeros(self.N))

    # ---- thin batched calls ------------------------------------------------------------------------------
    def lookahead_batch(self):
        self.sim.orca()
        self.sim.lookahead()
        return self.sim.la_reward, self.sim.la_done, self.sim.la_event

    def decide_batch(self):
        if getattr(self.policy, "weights_version", 0) != self._weights_version:
            self.sync_weights()
        return self.sim.decide()

    def step_batch(self, action_idx=None, action=None, active=None):
        self.sim.step(action_idx=action_idx, action=action, active=active, fused_orca=action_idx is None)
        return self.sim.reward, self.sim.done, self.sim.event

    # ---- the explorer's inner loop on the device (rl/utils/explorer.py:33-94) ------------------------------
    @torch.no_grad()
    def run_episodes(self, phase, seeds, epsilon=None, record=False, imitation=False, safety_space=0.0, rng=None):
        """Run every episode to its end.  Returns (EpisodeStats, trajectory or None); trajectory =
        dict(states [T, N, n, D], rewards [T, N], alive [T, N], rows [N]) on the device."""
        sim = self.sim
        self.reset_batch(phase, seeds)
        N, dev = self.N, self.device
        active = torch.ones(N, dtype=torch.uint8, device=dev)
        st = EpisodeStats(N)
        gamma_bar = self.cfg.gamma ** (self.time_step * self.robot.v_pref)
        disc = torch.ones(N, dtype=torch.float64, device=dev)
        cum = torch.zeros(N, dtype=torch.float64, device=dev)
        too_close = torch.zeros(N, dtype=torch.int64, device=dev)
        min_sum = torch.zeros(N, dtype=torch.float64, device=dev)
        steps = torch.zeros(N, dtype=torch.int64, device=dev)
        final_event = torch.zeros(N, dtype=torch.int64, device=dev)
        states, rewards, alive = [], [], []
        gen = rng if rng is not None else self._explore_rng
        max_steps = int(self.time_limit / self.time_step) + 2
        dd = torch.tensor([self.cfg.discomfort_dist_adult, self.cfg.discomfort_dist_bicycle,
                           self.cfg.discomfort_dist_child], dtype=torch.float64, device=dev)
        for _ in range(max_steps):
            if not bool(active.any()):
                break
            if record:
                states.append(sim.transform().clone())
            if imitation:
                act = sim.robot_orca(safety_space)
                sim.orca()
                sim.step(action=act, active=active)
            else:
                idx = self.decide_batch()
                if epsilon:
                    explore = torch.as_tensor(gen.random(N) < epsilon, device=dev)
                    rnd = torch.as_tensor(gen.integers(0, sim.A, N), dtype=torch.int32, device=dev)
                    idx = torch.where(explore, rnd, idx)
                sim.step(action_idx=idx.contiguous(), active=active)
            a = active.bool()
            r = torch.where(a, sim.reward, torch.zeros_like(sim.reward))
            cum += disc * r
            disc = torch.where(a, disc * gamma_bar, disc)
            steps += a.long()
            danger = a & (sim.event == abi.EV_DANGER)
            too_close += danger.long()
            # Danger.min_dist: the first type below its discomfort distance, child > bicycle > adult
            dm = sim.dmin
            md = torch.where(dm[:, 2] < dd[2], dm[:, 2], torch.where(dm[:, 1] < dd[1], dm[:, 1], dm[:, 0]))
            min_sum += torch.where(danger, md, torch.zeros_like(md))
            finished = a & sim.done.bool()
            final_event = torch.where(finished, sim.event.long(), final_event)
            if record:
                rewards.append(r.clone())
                alive.append(a.clone())
            active = (a & ~finished).to(torch.uint8)
        st.event = final_event.cpu().numpy()
        st.time = sim.time.cpu().numpy().copy()
        st.steps = steps.cpu().numpy()
        st.cum_reward = cum.cpu().numpy()
        st.too_close = too_close.cpu().numpy()
        st.min_dist_sum = min_sum.cpu().numpy()
        st.still_running = active.cpu().numpy().astype(bool)
        traj = None
        if record and states:
            traj = {"states": torch.stack(states), "rewards": torch.stack(rewards), "alive": torch.stack(alive),
                    "rows": (sim.hum_count + sim.stat_count).clone()}
        return st, traj
