"""SARL / EB-CADRL (rl/policy/sarl.py): the attention-pooling value network and its configuration.

`ValueNetwork` is the torch module with the reference's parameter names (mlp1.{0,2}, mlp2.{0,2},
attention.{0,2,4}, mlp3.{0,2,4,6}) — it owns the parameters for training (autograd) and checkpoint I/O;
inference inside the lookahead runs on the device's K4 kernels from the same state_dict."""
import logging

import torch
import torch.nn as nn

from rl.policy.cadrl import mlp
from rl.policy.multi_human_rl import MultiHumanRL


class ValueNetwork(nn.Module):
    def __init__(self, input_dim, self_state_dim, mlp1_dims, mlp2_dims, mlp3_dims, attention_dims,
                 with_global_state, cell_size, cell_num):
        super().__init__()
        self.self_state_dim = self_state_dim
        self.global_state_dim = mlp1_dims[-1]
        self.mlp1 = mlp(input_dim, mlp1_dims, last_relu=True)
        self.mlp2 = mlp(mlp1_dims[-1], mlp2_dims)
        self.with_global_state = with_global_state
        self.attention = mlp(mlp1_dims[-1] * (2 if with_global_state else 1), attention_dims)
        self.cell_size, self.cell_num = cell_size, cell_num
        self.mlp3 = mlp(mlp2_dims[-1] + self.self_state_dim, mlp3_dims)
        self._attention = None

    @property
    def attention_weights(self):
        """sarl.py:71 keeps `weights[0, :, 0].data.cpu().numpy()` on every forward; here the device -> host copy
        happens when somebody reads it (a copy per forward would be a host sync per optimizer step, and is illegal
        inside a CUDA-graph capture)."""
        return None if self._attention is None else self._attention.cpu().numpy()

    def forward(self, state, row_count=None):
        """state: (batch, n, D); optional row_count (batch,) masks zero-padded entity rows."""
        b, n, _ = state.shape
        self_state = state[:, 0, :self.self_state_dim]
        h1 = self.mlp1(state.reshape(b * n, -1))
        h2 = self.mlp2(h1)
        mask = None
        if row_count is not None:
            mask = (torch.arange(n, device=state.device)[None] < row_count[:, None]).to(h1.dtype)
        if self.with_global_state:
            h1v = h1.view(b, n, -1)
            if mask is None:
                g = h1v.mean(1, keepdim=True)
            else:
                g = (h1v * mask[..., None]).sum(1, keepdim=True) / row_count[:, None, None].to(h1.dtype)
            att_in = torch.cat([h1, g.expand(b, n, self.global_state_dim).reshape(b * n, -1)], dim=1)
        else:
            att_in = h1
        scores = self.attention(att_in).view(b, n)
        e = torch.exp(scores) * (scores != 0).float()          # masked softmax, no max subtraction (:69-70)
        if mask is not None:
            e = e * mask
        weights = (e / e.sum(dim=1, keepdim=True)).unsqueeze(2)
        self._attention = weights[0, :, 0].detach()
        pooled = (weights * h2.view(b, n, -1)).sum(dim=1)
        return self.mlp3(torch.cat([self_state, pooled], dim=1))


class SARL(MultiHumanRL):
    def __init__(self):
        super().__init__()
        self.name = "SARL"

    def configure(self, config):
        self.set_common_parameters(config)
        dims = lambda key: [int(x) for x in config.get("sarl", key).split(", ")]  # noqa: E731
        self.with_om = config.getboolean("sarl", "with_om")
        with_global_state = config.getboolean("sarl", "with_global_state")
        if config.has_option("sarl", "with_agent_type"):
            self.with_agent_type = config.getboolean("sarl", "with_agent_type")
            self.agent_type_state_dim = 4 if self.with_agent_type else 0
            self.joint_state_dim = self.self_state_dim + self.agent_state_dim + self.agent_type_state_dim
        self.model = ValueNetwork(self.input_dim(), self.self_state_dim, dims("mlp1_dims"), dims("mlp2_dims"),
                                  dims("mlp3_dims"), dims("attention_dims"), with_global_state, self.cell_size,
                                  self.cell_num)
        self.multiagent_training = config.getboolean("sarl", "multiagent_training")
        if self.with_om:
            self.name = "OM-SARL"
        logging.info("Policy: {} {} global state".format(self.name, "w/" if with_global_state else "w/o"))

    def get_attention_weights(self):
        return self.model.attention_weights
