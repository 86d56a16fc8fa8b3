"""Trainer (rl/utils/trainer.py:8-100): MSE regression of the value network on (state, value) pairs,
SGD with momentum 0.9 (or Adam).  One process per GPU: every rank draws its own batch from its own replay
shard and the flat gradient is summed with ONE NCCL all-reduce per optimizer step (379,202 fp32 = 1.5 MB,
latency-bound on NVLink) and divided by the number of ranks that had a batch — the only collective of the
training loop besides the statistics gather.

B200 shape of the loop: the reference's optimizer step is ~90 tiny kernels (forward, backward, SGD over 22
parameter tensors) and one host sync for the loss, i.e. pure launch latency (measured here: 1.96 ms per step in
eager PyTorch for a 0.4 M-parameter network).  This trainer keeps all parameters, gradients and momentum in THREE
flat buffers (the module's tensors are views), so the all-reduce needs no gather / scatter copies and the SGD
update is two fused element-wise kernels; on a CUDA device the whole step — batch gather from the replay memory,
forward, loss, backward, update — is captured once in a CUDA graph and replayed (one graph launch per step on one
GPU; gather+forward+backward graph, all-reduce, update graph with several ranks).  The loss is accumulated on the
device and read back once per call.  The arithmetic is the reference's: MSELoss, SGD(lr, momentum = 0.9)."""
import logging

import torch
import torch.nn as nn
import torch.optim as optim

MOMENTUM = 0.9      # trainer.py:21 of the reference


class Trainer(object):
    def __init__(self, model, memory, device, batch_size, policy=None, use_graphs=None):
        self.model = model
        self.device = torch.device(device) if not isinstance(device, torch.device) else device
        self.criterion = nn.MSELoss().to(device)
        self.memory = memory
        self.batch_size = batch_size
        self.optimizer = None          # "sgd" (flat buffers, below) or a torch.optim.Adam
        self.lr_scheduler = None
        self.policy = policy
        self.use_graphs = (self.device.type == "cuda") if use_graphs is None else bool(use_graphs)
        self._flat = self._grad = self._mom = None
        self._graph = None             # (graph_fb, graph_update or None, static index tensor)
        self._graph_key = None

    # ---- flat parameter / gradient / momentum buffers -------------------------------------------------------
    def _flatten(self):
        if self._flat is not None:
            return
        params = [p for p in self.model.parameters()]
        n = sum(p.numel() for p in params)
        dev, dt = params[0].device, params[0].dtype
        self._flat = torch.empty(n, device=dev, dtype=dt)
        self._grad = torch.zeros(n + 1, device=dev, dtype=dt)      # last slot: "this rank had a batch"
        self._mom = torch.zeros(n, device=dev, dtype=dt)
        off = 0
        for p in params:
            k = p.numel()
            self._flat[off:off + k].copy_(p.data.reshape(-1))
            p.data = self._flat[off:off + k].view_as(p)
            p.grad = self._grad[off:off + k].view_as(p)
            off += k
        self._n = n
        self._lr = torch.zeros(1, device=dev, dtype=dt)
        self._loss_acc = torch.zeros(1, device=dev, dtype=torch.float64)

    def set_optimizer(self, learning_rate, algorithm="sgd"):
        logging.info("Current learning rate: %f", learning_rate)
        self._flatten()
        self._graph = None
        if algorithm == "adam":
            self.optimizer = optim.Adam(self.model.parameters(), lr=learning_rate)
            self.lr_scheduler = optim.lr_scheduler.ReduceLROnPlateau(self.optimizer, factor=0.5, patience=10)
        else:
            self.optimizer = "sgd"     # optim.SGD(lr, momentum=0.9) restated on the flat buffers
            self.lr_scheduler = None
            self._lr.fill_(learning_rate)
            self._mom.zero_()          # a new optimizer starts without momentum (the reference builds a new SGD)

    def set_learning_rate(self, learning_rate):     # reference name kept
        self.set_optimizer(learning_rate)

    # ---- the three pieces of an optimizer step ---------------------------------------------------------------
    def _forward_backward(self, states, rows, values):
        self._grad.zero_()
        outputs = self.model(states, rows)
        loss = self.criterion(outputs, values)
        loss.backward()                      # accumulates into the views of self._grad
        self._grad[-1:].fill_(1.0)           # (a kernel, not a host scalar copy: legal inside a graph capture)
        self._loss_acc += loss.detach().double()

    def _apply_update(self, averaged):
        if averaged:
            self._grad[:-1].div_(self._grad[-1].clamp(min=1.0))
        if self.optimizer == "sgd":
            # torch.optim.SGD: buf = momentum * buf + grad (buf = grad on the first step); p -= lr * buf
            self._mom.mul_(MOMENTUM).add_(self._grad[:-1])
            self._flat.addcmul_(self._mom, self._lr.expand_as(self._mom), value=-1.0)
        else:
            self.optimizer.step()

    @staticmethod
    def _dist():
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            return dist
        return None

    def _agree(self, value, op="max"):
        """The same integer on every rank (max or min over ranks): per-rank replay shards differ in size (imitation
        learning keeps only success / collision episodes), the number of optimizer steps must not."""
        dist = self._dist()
        if dist is None:
            return int(value)
        t = torch.tensor([int(value)], dtype=torch.int64, device=self.device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX if op == "max" else dist.ReduceOp.MIN)
        return int(t.item())

    def _capture(self):
        """CUDA graphs of one optimizer step over a static batch-index tensor (SGD on a CUDA device only)."""
        mem = self.memory
        key = (mem.states.data_ptr(), mem.values.data_ptr(), self.batch_size, self._dist() is not None)
        if self._graph is not None and self._graph_key == key:
            return self._graph
        idx = torch.zeros(self.batch_size, dtype=torch.int64, device=self.device)
        multi = self._dist() is not None

        def fb():
            self._forward_backward(mem.states.index_select(0, idx), mem.rows.index_select(0, idx),
                                   mem.values.index_select(0, idx))

        keep = (self._flat.clone(), self._mom.clone(), self._loss_acc.clone())
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):           # warm-up outside capture (autograd / cuBLAS workspaces)
            for _ in range(3):
                fb()
                self._apply_update(multi)
        torch.cuda.current_stream(self.device).wait_stream(side)
        g_fb, g_up = torch.cuda.CUDAGraph(), None
        if multi:
            g_up = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g_fb):
                fb()
            with torch.cuda.graph(g_up):
                self._apply_update(True)
        else:
            with torch.cuda.graph(g_fb):
                fb()
                self._apply_update(False)
        self._flat.copy_(keep[0]); self._mom.copy_(keep[1]); self._loss_acc.copy_(keep[2])   # undo the warm-up steps
        self._graph, self._graph_key = (g_fb, g_up, idx), key
        return self._graph

    def _step(self, batch=None, index=None):
        """One optimizer step.  `index`: a batch of replay-memory indices (graph path when it is a full batch);
        `batch` = (states, rows, values) for the eager path; both None: this rank's shard is empty (it contributes a
        zero gradient but still takes part in the all-reduce, so that the collectives of all ranks pair up)."""
        dist = self._dist()
        graphed = (self.use_graphs and self.optimizer == "sgd" and index is not None and len(index) == self.batch_size)
        if graphed:
            g_fb, g_up, idx = self._capture()
            idx.copy_(index)
            g_fb.replay()
            if dist is not None:
                dist.all_reduce(self._grad, op=dist.ReduceOp.SUM)
                g_up.replay()
        else:
            if index is not None:
                mem = self.memory
                batch = (mem.states[index].to(self.device), mem.rows[index].to(self.device), mem.values[index].to(self.device))
            if batch is not None:
                self._forward_backward(*batch)
            else:
                self._grad.zero_()
            if dist is not None:
                dist.all_reduce(self._grad, op=dist.ReduceOp.SUM)
            self._apply_update(dist is not None)
        if self.policy is not None:
            self.policy.weights_version += 1       # the device copy of the weights is stale now

    def _take_loss(self):
        v = float(self._loss_acc.item())           # the one host read-back of the call
        self._loss_acc.zero_()
        return v

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
            for i in range(0, n_max, self.batch_size):
                self._step(index=perm[i:i + self.batch_size] if n_own else None)
            average_epoch_loss = self._take_loss() / max(n_max, 1)
            if self.lr_scheduler is not None:
                self.lr_scheduler.step(average_epoch_loss)
            logging.debug("Average loss in epoch %d: %.2E", epoch, average_epoch_loss)
        return average_epoch_loss

    def optimize_batch(self, num_batches):
        """`num_batches` independent random batches (trainer.py:74-100)."""
        if self.optimizer is None:
            raise ValueError("Learning rate is not set!")
        n_own = len(self.memory)
        if n_own:       # every batch's indices in one launch (the reference draws random.sample per batch)
            idx_all = torch.randint(0, n_own, (num_batches, self.batch_size), device=self.memory.device)
        for i in range(num_batches):
            self._step(index=idx_all[i] if n_own else None)
        average_loss = self._take_loss() / max(num_batches, 1)
        if self.lr_scheduler is not None:
            self.lr_scheduler.step(average_loss)
        logging.debug("Average loss : %.2E", average_loss)
        return average_loss
