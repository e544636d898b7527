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
        self.params = [p for p in model.parameters()]
        fused = self.params[0].is_cuda
        self.opt = torch.optim.Adam(self.params, lr=lr, fused=fused) if fused else torch.optim.Adam(self.params, lr=lr)

    def step(self, batch):
        """One optimisation step on a device-resident batch; returns the (device) loss tensor.

        Gradients start as None, so autograd hands every parameter its freshly computed gradient tensor instead of
        accumulating into a zeroed buffer (44 tiny add kernels per step for the ZINC model).  With more than one rank the
        gradients are flattened into ONE bucket, SUM-all-reduced (all losses use reduction='sum', Zinc12k.py:365) and the
        parameters' .grad become views of the reduced bucket."""
        for p in self.params:
            p.grad = None
        out = self.model(batch)
        loss = loss_fn(self.loss_kind, out, batch.y)
        loss.backward()
        if self.distributed:
            grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in self.params]
            flat = torch.cat([g.reshape(-1) for g in grads])
            dist.all_reduce(flat, op=dist.ReduceOp.SUM)
            off = 0
            for p in self.params:
                p.grad = flat[off:off + p.numel()].view_as(p)
                off += p.numel()
        self.opt.step()
        return loss.detach()


class HostFeeder(object):
    """Streams pinned host batches to the device one step ahead of the compute stream (H2D overlaps compute).

    Host ``Batch`` objects are converted ONCE (first time they are seen) to the compact wire format (``batch.CompactBatch``:
    int32 per-graph sizes instead of the int64 batch vector, graph-local uint8 edge ids instead of int64 global ids, one-hot
    ``x`` as uint8 class codes); every step copies the compact tensors host -> device on the copy stream and the reference's
    attributes are rebuilt on the device (``expand``) on the compute stream."""

    format_note = ("compact wire format: n/e int32 per graph, graph-local uint8 edge ids, one-hot x as uint8 class codes, FP32 "
                   "supports; edge_index2 (int64, global), batch and x are rebuilt on the device")

    def __init__(self, device, onehot_widths=None):
        self.device = device
        self.copy_stream = torch.cuda.Stream(device=device)
        self.onehot_widths = onehot_widths
        self._next = None
        self._compact = {}

    def compact(self, host_batch):
        from .batch import CompactBatch
        key = id(host_batch)
        cb = self._compact.get(key)
        if cb is None:
            if isinstance(host_batch, CompactBatch):
                cb = host_batch
            else:
                cb = CompactBatch.from_batch(host_batch, self.onehot_widths).pin_memory()
            self._compact[key] = cb
        return cb

    def bytes_per_batch(self, host_batch):
        return self.compact(host_batch).nbytes()

    def prefetch(self, host_batch):
        cb = self.compact(host_batch)
        with torch.cuda.stream(self.copy_stream):
            b = cb.to(self.device, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(self.copy_stream)
        self._next = (b, ev)

    def get(self):
        b, ev = self._next
        torch.cuda.current_stream().wait_event(ev)
        for v in b._tensors().values():
            v.record_stream(torch.cuda.current_stream())
        self._next = None
        return b.expand()
