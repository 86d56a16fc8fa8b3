"""Trainer (rl/utils/trainer.py:8-100): MSE regression of the value network on (state, value) pairs,
SGD with momentum 0.9 (or Adam).  One process per GPU: every rank draws its own batch from its own replay
shard and the flat gradient is summed with ONE NCCL all-reduce per optimizer step (379,202 fp32 = 1.5 MB,
latency-bound on NVLink) and divided by the world size — the only collective of the training loop besides
the statistics gather."""
import logging

import torch
import torch.nn as nn
import torch.optim as optim


class Trainer(object):
    def __init__(self, model, memory, device, batch_size, policy=None):
        self.model = model
        self.device = device
        self.criterion = nn.MSELoss().to(device)
        self.memory = memory
        self.batch_size = batch_size
        self.optimizer = None
        self.lr_scheduler = None
        self.policy = policy
        self._flat = None

    def set_optimizer(self, learning_rate, algorithm="sgd"):
        logging.info("Current learning rate: %f", learning_rate)
        if algorithm == "adam":
            self.optimizer = optim.Adam(self.model.parameters(), lr=learning_rate)
            self.lr_scheduler = optim.lr_scheduler.ReduceLROnPlateau(self.optimizer, factor=0.5, patience=10)
        else:
            self.optimizer = optim.SGD(self.model.parameters(), lr=learning_rate, momentum=0.9)
            self.lr_scheduler = None

    def set_learning_rate(self, learning_rate):     # reference name kept
        self.set_optimizer(learning_rate)

    def _allreduce_grads(self):
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
            return
        grads = [p.grad for p in self.model.parameters() if p.grad is not None]
        flat = torch.cat([g.reshape(-1) for g in grads])          # one bucket
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        flat /= dist.get_world_size()
        off = 0
        for g in grads:
            g.copy_(flat[off:off + g.numel()].view_as(g))
            off += g.numel()

    def _step(self, states, rows, values):
        self.optimizer.zero_grad()
        outputs = self.model(states, rows)
        loss = self.criterion(outputs, values)
        loss.backward()
        self._allreduce_grads()
        self.optimizer.step()
        if self.policy is not None:
            self.policy.weights_version += 1       # the device copy of the weights is stale now
        return loss.data.item()

    def optimize_epoch(self, num_epochs):
        """Whole passes over the memory in shuffled batches (trainer.py:47-72)."""
        if self.optimizer is None:
            raise ValueError("Learning rate is not set!")
        average_epoch_loss = 0
        for epoch in range(num_epochs):
            perm = torch.randperm(len(self.memory), device=self.memory.device)
            epoch_loss = 0
            for i in range(0, len(perm), self.batch_size):
                idx = perm[i:i + self.batch_size]
                epoch_loss += self._step(self.memory.states[idx].to(self.device), self.memory.rows[idx].to(self.device),
                                         self.memory.values[idx].to(self.device))
            average_epoch_loss = epoch_loss / max(len(self.memory), 1)
            if self.lr_scheduler is not None:
                self.lr_scheduler.step(average_epoch_loss)
            logging.debug("Average loss in epoch %d: %.2E", epoch, average_epoch_loss)
        return average_epoch_loss

    def optimize_batch(self, num_batches):
        """`num_batches` independent random batches (trainer.py:74-100)."""
        if self.optimizer is None:
            raise ValueError("Learning rate is not set!")
        losses = 0
        for _ in range(num_batches):
            states, rows, values = self.memory.sample(self.batch_size)
            losses += self._step(states.to(self.device), rows.to(self.device), values.to(self.device))
        average_loss = losses / max(num_batches, 1)
        if self.lr_scheduler is not None:
            self.lr_scheduler.step(average_loss)
        logging.debug("Average loss : %.2E", average_loss)
        return average_loss
