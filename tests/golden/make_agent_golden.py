#!/usr/bin/env python
"""Golden vectors for the host-side agent kinematics, from the UNMODIFIED reference.

Runs only in the build container (needs /root/reference); writes tests/golden/agent_kinematics.npz,
which is committed and read by tests/test_host_mirror.py::test_agent_kinematics_match_reference.

    python tests/golden/make_agent_golden.py          (re-executes itself with the shim environment)

Recorded, per case, from /root/reference/simulator/agents/agent.py: compute_position :164-188,
compute_velocity :190-200, get_next_observable_state :80-93, step :202-228, reached_destination :230-234,
for holonomic agents (ActionXY) and non-holonomic ones (ActionRot; ActionXYRot where the reference
supports it -- its step() reads `action.v` of an ActionXYRot and raises, so step is recorded for the
first two kinds only)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("EBC_REFERENCE", "/root/reference")

if os.environ.get("EBC_GOLDEN_CHILD") != "1":
    env = dict(os.environ)
    env["EBC_GOLDEN_CHILD"] = "1"
    env["PYTHONPATH"] = os.pathsep.join([os.path.join(REPO, "oracle", "shims"), REF])
    sys.exit(subprocess.call([sys.executable, os.path.abspath(__file__)] + sys.argv[1:], cwd=REF, env=env))

import configparser  # noqa: E402

import numpy as np  # noqa: E402
from simulator.agents.agent import Agent  # noqa: E402
from simulator.utils.action import ActionRot, ActionXY, ActionXYRot  # noqa: E402
from simulator.utils.utils import AgentType  # noqa: E402

N = 256
DT = 0.25


class _Kin:
    def __init__(self, kinematics):
        self.kinematics = kinematics


def make_agent(kinematics):
    cfg = configparser.RawConfigParser()
    cfg.add_section("a")
    for k, v in (("visible", "true"), ("v_pref", "1"), ("radius", "0.3"), ("policy", "none"), ("sensor", "coordinates")):
        cfg.set("a", k, v)
    agent = Agent(cfg, "a")
    agent.set_policy(_Kin(kinematics))
    agent.time_step = DT
    return agent


def main():
    rng = np.random.default_rng(20261019)
    pose = np.empty((N, 9))                       # px py gx gy vx vy theta radius v_pref
    pose[:, 0:4] = rng.uniform(-6, 6, (N, 4))
    pose[:, 4:6] = rng.uniform(-1.5, 1.5, (N, 2))
    pose[:, 6] = rng.uniform(-0.5, 2 * np.pi + 0.5, N)
    pose[:, 7] = rng.uniform(0.1, 0.6, N)
    pose[:, 8] = rng.uniform(0.3, 1.5, N)
    pose[::17, 2:4] = pose[::17, 0:2] + 0.05      # some agents inside their goal disc
    act = np.empty((N, 3))                        # (vx, vy, r) / (v, -, r)
    act[:, 0:2] = rng.uniform(-1.5, 1.5, (N, 2))
    act[:, 2] = rng.uniform(-np.pi / 4, np.pi / 4, N)
    out = {"pose": pose, "act": act, "dt": np.float64(DT)}
    for kind in ("xy", "rot", "xyrot"):
        agent = make_agent("holonomic" if kind == "xy" else "unicycle")
        pos = np.empty((N, 2)); vel = np.empty((N, 2)); nxt = np.empty((N, 5)); stp = np.empty((N, 5)); reach = np.empty(N, bool)
        for i in range(N):
            a = {"xy": ActionXY(act[i, 0], act[i, 1]), "rot": ActionRot(act[i, 0], act[i, 2]),
                 "xyrot": ActionXYRot(act[i, 0], act[i, 1], act[i, 2])}[kind]
            agent.set(*pose[i, 0:7], radius=pose[i, 7], v_pref=pose[i, 8], agent_type=AgentType.ADULT)
            reach[i] = agent.reached_destination()
            pos[i] = agent.compute_position(a, DT)
            if kind != "xy":
                vel[i] = agent.compute_velocity(a)
            if kind != "xyrot":
                o = agent.get_next_observable_state(a)
                nxt[i] = (o.px, o.py, o.vx, o.vy, o.radius)
                agent.step(a)
                stp[i] = (agent.px, agent.py, agent.vx, agent.vy, agent.theta)
        out[kind + "_pos"] = pos
        out[kind + "_reach"] = reach
        if kind != "xy":
            out[kind + "_vel"] = vel
        if kind != "xyrot":
            out[kind + "_next"] = nxt
            out[kind + "_step"] = stp
    np.savez_compressed(os.path.join(HERE, "agent_kinematics.npz"), **out)
    print("agent_kinematics.npz:", {k: np.shape(v) for k, v in out.items()})


if __name__ == "__main__":
    main()
