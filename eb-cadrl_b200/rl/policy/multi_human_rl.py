"""MultiHumanRL — the one-step-lookahead decision rule (rl/policy/multi_human_rl.py:12-87,128-154).

`predict(state, env)` keeps the reference's contract (same checks and exception types, epsilon-greedy with
numpy's global RNG drawn on every call, strict first-maximum argmax, `action_values`, `last_state`) but the
81-action loop is four kernel launches on the device: ORCA (K1), lookahead + joint-state build (K3), value
network (K4), selection (K5).  OM-SARL's occupancy maps and the `query_env = false` path (dead in the
reference: `compute_reward` raises) are out of scope."""
import numpy as np
import torch

from rl.policy.cadrl import CADRL
from simulator.utils.action import ActionRot, ActionXY


class MultiHumanRL(CADRL):
    def __init__(self):
        super().__init__()
        self.value_mode = None       # None = library default (tcgen05 fp32-accurate); see ebc_set_value_mode

    def predict(self, state, env=None):
        if self.phase is None:
            raise AttributeError("Phase attribute has to be set!")
        if self.device is None:
            raise AttributeError("Device attributes has to be set!")
        if self.phase == "train" and self.epsilon is None:
            raise AttributeError("Epsilon attribute has to be set in training phase")
        if self.reach_destination(state):
            return ActionXY(0, 0) if self.kinematics == "holonomic" else ActionRot(0, 0)
        if self.action_space is None:
            self.build_action_space(state.self_state.v_pref)
        if self.with_om or not self.query_env:
            raise NotImplementedError("occupancy maps / query_env = false are outside the B200 hot path")
        env.bind_policy(self)
        sim = env.native
        probability = np.random.random()
        if self.phase == "train" and probability < self.epsilon:
            max_action = self.action_space[np.random.choice(len(self.action_space))]
        else:
            keeps_attention = hasattr(self, "get_attention_weights") and hasattr(self.model, "_attention")
            if keeps_attention and sim.attention is None:
                sim.enable_attention()
            sim.decide()
            if int(sim.nan_flag[0].item()):
                raise ValueError("Value network is not well trained. ")
            self.action_values = sim.action_values[0].cpu().tolist()
            if keeps_attention:
                # the reference keeps the weights of its LAST forward (sarl.py:71) = the last action of the loop
                rows = int(sim.hum_count[0].item()) + int(sim.stat_count[0].item())
                self.model._attention = sim.attention[sim.A - 1, :rows].clone()
            max_action = self.action_space[int(sim.argmax[0].item())]
        if self.phase == "train":
            self.last_state = self.transform(state, env)
        return max_action

    def transform(self, state, env=None):
        """Rotated CURRENT joint state, (n, D) — the replay sample (multi_human_rl.py:128-149)."""
        if env is not None and getattr(env, "native", None) is not None:
            sim = env.native
            n = int(sim.hum_count[0].item()) + int(sim.stat_count[0].item())
            return sim.transform()[0, :n].to(self.device)
        rows = torch.cat([torch.Tensor([state.self_state + a]).to(self.device) for a in state.agent_states], dim=0)
        return self.rotate(rows)

    def input_dim(self):
        return self.joint_state_dim + (self.cell_num ** 2 * self.om_channel_size if self.with_om else 0)
