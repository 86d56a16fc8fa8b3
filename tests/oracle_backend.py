"""Test-only adapter: the CPU oracle (oracle/_ref/libebc_oracle.so, `ebc_ref_*`) behind the
same backend interface the engine uses for libebcadrl.so.  Lives in tests/ on purpose: the
product package never loads the oracle."""
import ctypes
import json
import os
import subprocess

import numpy as np

from ebc import abi
from ebc.config import SimConfig

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "_ref", "libebc_oracle.so")
GOLDEN = os.path.join(ROOT, "tests", "golden")


def build_oracle():
    if not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(
            os.path.join(ROOT, "oracle", "ebc_oracle.c")):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")], stdout=subprocess.DEVNULL)
    return ORACLE_SO


class OracleBackend(object):
    name = "oracle"

    def __init__(self):
        self.lib = ctypes.CDLL(build_oracle())
        vp, c_i32 = ctypes.c_void_p, ctypes.c_int32
        L = self.lib
        L.ebc_ref_create.argtypes = [ctypes.POINTER(abi.EbcConfig), ctypes.POINTER(vp)]
        L.ebc_ref_destroy.argtypes = [vp]
        L.ebc_ref_last_error.argtypes = [vp]
        L.ebc_ref_last_error.restype = ctypes.c_char_p
        L.ebc_ref_bind.argtypes = [vp, ctypes.POINTER(abi.EbcState)]
        L.ebc_ref_set_actions.argtypes = [vp, vp, c_i32]
        L.ebc_ref_set_weights.argtypes = [vp, ctypes.POINTER(abi.EbcWeights)]
        L.ebc_ref_orca.argtypes = [vp]
        L.ebc_ref_robot_orca.argtypes = [vp, ctypes.c_double, vp]
        L.ebc_ref_lookahead.argtypes = [vp] * 5
        L.ebc_ref_value.argtypes = [vp, vp, ctypes.c_int64, vp, vp]
        L.ebc_ref_select.argtypes = [vp] * 6
        L.ebc_ref_set_attention_output.argtypes = [vp, vp]
        L.ebc_ref_step.argtypes = [vp] * 9
        L.ebc_ref_transform.argtypes = [vp, vp]
        L.ebc_ref_reset.argtypes = [vp, ctypes.POINTER(abi.EbcState), c_i32, vp, vp]
        L.ebc_ref_local_map_angular.argtypes = [vp, ctypes.POINTER(abi.EbcAngularMap), vp, vp, vp]
        L.ebc_ref_local_map_grid.argtypes = [vp, ctypes.POINTER(abi.EbcGridMap), vp]
        L.ebc_ref_set_threads.argtypes = [ctypes.c_int]

    def set_threads(self, n):
        self.lib.ebc_ref_set_threads(n)

    def create(self, cfg, device_index):
        h = ctypes.c_void_p()
        rc = self.lib.ebc_ref_create(ctypes.byref(cfg), ctypes.byref(h))
        if rc != 0:
            raise abi.EbcError("ebc_ref_create: %s" % self.lib.ebc_ref_last_error(None).decode())
        return h

    def destroy(self, h):
        self.lib.ebc_ref_destroy(h)

    def last_error(self, h):
        return self.lib.ebc_ref_last_error(h).decode()

    def call(self, name, h, *args, stream=None):
        if name == "reserve":     # the CPU twin has no scratch to reserve
            return 0
        if name == "orca_step":   # the oracle has no fused twin: same result by definition
            rc = self.lib.ebc_ref_orca(h)
            if rc == 0:
                rc = self.lib.ebc_ref_step(h, *args)
        else:
            rc = getattr(self.lib, "ebc_ref_" + name)(h, *args)
        if rc != 0:
            raise abi.EbcError("ebc_ref_%s failed (%d): %s" % (name, rc, self.last_error(h)))

    def launch_count(self, h):
        return 0


# ---- golden fixtures -------------------------------------------------------------------
class Trace(object):
    """One tests/golden/*.npz step trace written by tests/golden/make_golden.py."""

    def __init__(self, name):
        self.name = name
        self.z = np.load(os.path.join(GOLDEN, name + ".npz"))
        self.meta = json.loads(bytes(self.z["meta_json"]).decode())
        self.cfg_dict = self.meta["cfg"]
        self.steps = self.meta["steps"]
        self.n_steps = self.meta["n_steps"]

    def get(self, t, key):
        return self.z["s%03d_%s" % (t, key)]

    def has(self, t, key):
        return ("s%03d_%s" % (t, key)) in self.z.files

    def sim_config(self, safety_space=0.0):
        d = self.cfg_dict
        c = SimConfig()
        for k, v in d.items():
            if k in ("robot_kinematics", "rotate_theta", "human_policy", "has_max_goal_distance"):
                continue
            setattr(c, k, v)
        c.new_reward = bool(d["new_reward"])
        c.with_agent_type = bool(d["with_agent_type"])
        c.robot_visible = bool(d["robot_visible"])
        if not d["has_max_goal_distance"]:
            c.max_goal_distance = None
        if d["robot_kinematics"] == 0:
            c.robot_kinematics = "holonomic"
        else:
            c.robot_kinematics = "unicycle" if d["rotate_theta"] else "nonholonomic"
        c.human_policy = ["orca" if p == 0 else "linear" for p in d["human_policy"]]
        c.orca_safety_space = safety_space
        return c

    def map_zero(self):
        shape = tuple(int(x) for x in self.z["map_shape"])
        bits = np.unpackbits(self.z["map_zero_packed"])[: shape[0] * shape[1]]
        return bits.reshape(shape).astype(bool)

    def load_into(self, sim, t, episode=0):
        """Put the recorded (fp32-representable) state of step t into episode `episode`."""
        from ebc.scene import rects_from_zero_cells
        pv, gr, ty = self.get(t, "hum_pv"), self.get(t, "hum_gr"), self.get(t, "hum_type")
        st = self.get(t, "stat")
        H, S = len(ty), len(st)
        rects = rects_from_zero_cells(self.map_zero())
        hp = np.zeros((1, sim.Hmax, 4), np.float32); hp[0, :H] = pv
        hg = np.zeros((1, sim.Hmax, 4), np.float32); hg[0, :H] = gr
        ht = np.zeros((1, sim.Hmax), np.uint8); ht[0, :H] = ty
        sd = np.zeros((1, max(sim.Smax, 1), 4), np.float32); sd[0, :S] = st
        rc = np.zeros((1, max(sim.Rmax, 1), 4), np.int16); rc[0, :len(rects)] = rects
        assert H <= sim.Hmax and S <= sim.Smax and len(rects) <= sim.Rmax
        sim.load_episodes(episode, hp, hg, ht, np.array([H], np.int32), sd, np.array([S], np.int32),
                          rc, np.array([len(rects)], np.int32),
                          self.get(t, "rob_pv")[None].astype(np.float32),
                          self.get(t, "rob_gr")[None].astype(np.float32),
                          np.array([self.get(t, "rob_theta")], np.float32),
                          np.array([self.get(t, "time")], np.float64))
        return H, S

    def dims(self):
        H = len(self.get(0, "hum_type"))
        S = len(self.get(0, "stat"))
        from ebc.scene import rects_from_zero_cells
        R = len(rects_from_zero_cells(self.map_zero()))
        return H, S, R


def load_weights(name):
    z = np.load(os.path.join(GOLDEN, name))
    return {k: z[k] for k in z.files}


TRACE_WEIGHTS = {
    "trace_cfg1_adults5_seed1002": "weights_sarl_baseline.npz",
    "trace_adults3_bikes3_static2_seed1002": "weights_sarl_baseline.npz",
    "trace_ebcadrl_h24_seed1000000": "weights_ebcadrl.npz",
    "trace_cfg2_h10_seed7": "weights_ebcadrl.npz",
    "trace_cfg2_h10_seed11": "weights_ebcadrl.npz",
    "trace_unicycle_adults5_seed1003": "weights_sarl_baseline.npz",
    "trace_nonholonomic_adults5_seed1003": "weights_sarl_baseline.npz",
    "trace_visible_adults5_seed1004": "weights_sarl_baseline.npz",
    # round 2: `linear` humans (all / mixed with ORCA), train-phase traces that carry policy.last_state
    "trace_linear_adults5_seed1006": "weights_sarl_baseline.npz",
    "trace_mixed_linear_bikes_seed1007": "weights_sarl_baseline.npz",
    "trace_train_cfg2_h10_seed13": "weights_ebcadrl.npz",
    "trace_train_unicycle_adults5_seed1008": "weights_sarl_baseline.npz",
}
LINEAR_SCENES = ["scene_collision_with_adult", "scene_collision_with_bicycle", "scene_collision_with_static",
                 "scene_no_collisions", "scene_bikes_0_collision_with_adult_1",
                 "scene_bikes_0_collision_with_adult_2", "scene_bikes_0_no_collisions",
                 "scene_collision_with_child"]
