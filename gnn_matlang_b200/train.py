"""Training step of the reference's scripts (Zinc12k.py:354-371, counting.py:400-417, exp_classify.py:318-337):
forward, SUM-reduced loss, backward, Adam(lr=1e-3) -- with data parallelism over the graph-minibatch axis.

Graphs never interact (block-diagonal supports, per-graph readout), so a global minibatch is sharded over the
ranks as whole graphs and the only exchange is ONE SUM all-reduce of the flat gradient buffer per step (SUM,
not mean: every loss of the reference uses reduction='sum', so the single-process result for the same global
batch is reproduced).  Parameters and Adam state are replicated.
"""
import torch
import torch.distributed as dist
import torch.nn.functional as F


def loss_fn(kind, out, y):
    if kind == "l1":                                   # Zinc12k.py:365
        return F.l1_loss(out, y.to(out.dtype).reshape(out.shape), reduction="sum")
    if kind == "mse":                                  # counting.py:411
        return torch.square(out - y.to(out.dtype).reshape(out.shape)).sum()
    if kind == "bce":                                  # exp_classify.py:327-329
        return F.binary_cross_entropy(torch.sigmoid(out), y.to(out.dtype).reshape(out.shape), reduction="sum")
    raise ValueError(kind)


class Trainer(object):
    def __init__(self, model, loss="l1", lr=1e-3, distributed=False):
        self.model = model
        self.loss_kind = loss
        self.distributed = distributed and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        params = [p for p in model.parameters()]
        # one flat gradient bucket; every p.grad is a view into it (autograd accumulates in place)
        self.flat_grad = torch.zeros(sum(p.numel() for p in params), dtype=torch.float32, device=params[0].device)
        off = 0
        for p in params:
            p.grad = self.flat_grad[off:off + p.numel()].view_as(p)
            off += p.numel()
        fused = params[0].is_cuda
        self.opt = torch.optim.Adam(params, lr=lr, fused=fused) if fused else torch.optim.Adam(params, lr=lr)

    def step(self, batch):
        """One optimisation step on a device-resident batch; returns the (device) loss tensor."""
        self.flat_grad.zero_()
        out = self.model(batch)
        loss = loss_fn(self.loss_kind, out, batch.y)
        loss.backward()
        if self.distributed:
            dist.all_reduce(self.flat_grad, op=dist.ReduceOp.SUM)
        self.opt.step()
        return loss.detach()


class HostFeeder(object):
    """Streams pinned host batches to the device one step ahead of the compute stream (H2D overlaps compute)."""

    def __init__(self, device):
        self.device = device
        self.copy_stream = torch.cuda.Stream(device=device)
        self._next = None

    def prefetch(self, host_batch):
        with torch.cuda.stream(self.copy_stream):
            b = host_batch.to(self.device, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(self.copy_stream)
        self._next = (b, ev)

    def get(self):
        b, ev = self._next
        torch.cuda.current_stream().wait_event(ev)
        for v in b.__dict__.values():
            if isinstance(v, torch.Tensor):
                v.record_stream(torch.cuda.current_stream())
        self._next = None
        return b
