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

    @staticmethod
    def _dist():
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            return dist
        return None

    def _allreduce_grads(self, contributed):
        """Sum the flat gradient over the ranks (one bucket, one collective) and divide by the number of ranks that
        had a batch.  EVERY rank calls this exactly once per optimizer step, batch or not: the step counts are
        agreed beforehand (`_agree`), so the collectives of different ranks always pair up."""
        dist = self._dist()
        if dist is None:
            return
        params = [p for p in self.model.parameters() if p.requires_grad]
        flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in params] +
                         [torch.full((1,), float(contributed), device=params[0].device, dtype=params[0].dtype)])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        flat[:-1] /= flat[-1].clamp(min=1.0)
        off = 0
        for p in params:
            g = flat[off:off + p.numel()].view_as(p)
            if p.grad is None:
                p.grad = g.clone()
            else:
                p.grad.copy_(g)
            off += p.numel()

    def _agree(self, value, op="max"):
        """The same integer on every rank (max or min over ranks): per-rank replay shards differ in size (imitation
        learning keeps only success / collision episodes), the number of optimizer steps must not."""
        dist = self._dist()
        if dist is None:
            return int(value)
        t = torch.tensor([int(value)], dtype=torch.int64, device=self.device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX if op == "max" else dist.ReduceOp.MIN)
        return int(t.item())

    def _step(self, batch):
        """One optimizer step; `batch` = (states, rows, values) or None when this rank's shard is empty (it then
        contributes a zero gradient but still takes part in the all-reduce)."""
        self.optimizer.zero_grad()
        loss_value = 0.0
        if batch is not None:
            states, rows, values = batch
            outputs = self.model(states, rows)
            loss = self.criterion(outputs, values)
            loss.backward()
            loss_value = loss.data.item()
        self._allreduce_grads(batch is not None)
        self.optimizer.step()
        if self.policy is not None:
            self.policy.weights_version += 1       # the device copy of the weights is stale now
        return loss_value

    def _gather_batch(self, idx):
        m = self.memory
        return m.states[idx].to(self.device), m.rows[idx].to(self.device), m.values[idx].to(self.device)

    def optimize_epoch(self, num_epochs):
        """Whole passes over the memory in shuffled batches (trainer.py:47-72).  With several ranks every rank runs
        ceil(max_r len(memory_r) / batch_size) steps per epoch; a shorter shard is padded by re-drawing from itself
        (sampling with replacement), an empty one sends zero gradients."""
        if self.optimizer is None:
            raise ValueError("Learning rate is not set!")
        average_epoch_loss = 0
        n_own = len(self.memory)
        n_max = self._agree(n_own, "max")
        for epoch in range(num_epochs):
            if n_own:
                perm = torch.randperm(n_own, device=self.memory.device)
                if n_max > n_own:
                    perm = torch.cat([perm, torch.randint(0, n_own, (n_max - n_own,), device=self.memory.device)])
            epoch_loss = 0
            for i in range(0, n_max, self.batch_size):
                epoch_loss += self._step(self._gather_batch(perm[i:i + self.batch_size]) if n_own else None)
            average_epoch_loss = epoch_loss / max(n_max, 1)
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
            batch = None
            if len(self.memory):
                states, rows, values = self.memory.sample(self.batch_size)
                batch = (states.to(self.device), rows.to(self.device), values.to(self.device))
            losses += self._step(batch)
        average_loss = losses / max(num_batches, 1)
        if self.lr_scheduler is not None:
            self.lr_scheduler.step(average_loss)
        logging.debug("Average loss : %.2E", average_loss)
        return average_loss
