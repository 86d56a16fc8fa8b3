"""Replay memory (rl/utils/memory.py:4-28): a ring buffer of (rotated joint state, value).  Device-resident:
states are padded to n rows with their real row count kept alongside, so a batch is one gather."""
import torch


class ReplayMemory(object):
    def __init__(self, capacity, n_rows=None, row_dim=None, device="cpu"):
        self.capacity = int(capacity)
        self.device = torch.device(device)
        self.position = 0
        self.size = 0
        self.states = self.rows = self.values = None
        if n_rows is not None:
            self._alloc(n_rows, row_dim)

    def _alloc(self, n_rows, row_dim):
        self.states = torch.zeros(self.capacity, n_rows, row_dim, dtype=torch.float32, device=self.device)
        self.rows = torch.zeros(self.capacity, dtype=torch.int32, device=self.device)
        self.values = torch.zeros(self.capacity, 1, dtype=torch.float32, device=self.device)

    def push(self, item):
        """item = (state [rows, D], value [1]) like the reference; rows may be < n."""
        state, value = item
        self.push_batch(state.unsqueeze(0), torch.tensor([state.shape[0]]), torch.as_tensor(value).reshape(1, 1))

    def push_batch(self, states, rows, values):
        if self.states is None:
            self._alloc(states.shape[1], states.shape[2])
        k = states.shape[0]
        idx = (self.position + torch.arange(k, device=self.device)) % self.capacity
        n = min(states.shape[1], self.states.shape[1])
        self.states[idx] = 0
        self.states[idx, :n] = states[:, :n].to(self.device, torch.float32)
        self.rows[idx] = rows.to(self.device, torch.int32)
        self.values[idx] = values.to(self.device, torch.float32).reshape(k, 1)
        self.position = int((self.position + k) % self.capacity)
        self.size = min(self.size + k, self.capacity)

    def sample(self, batch_size, generator=None):
        idx = torch.randint(0, self.size, (batch_size,), device=self.device, generator=generator)
        return self.states[idx], self.rows[idx], self.values[idx]

    def __getitem__(self, i):
        return self.states[i, :int(self.rows[i])], self.values[i]

    def __len__(self):
        return self.size

    def is_full(self):
        return self.size == self.capacity

    def clear(self):
        self.position = self.size = 0
