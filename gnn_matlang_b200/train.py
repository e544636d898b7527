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


def shard_graphs(sizes, world, rank=None):
    """Contiguous split of a global minibatch over ``world`` ranks balanced by the graphs' sizes (nodes or support entries) and
    not by their count (30-100-node mixes): the cut points are the positions where the running sum of ``sizes``
    crosses k / world of the total.  Returns the list of (begin, end) graph ranges, or rank's range.  Pure host arithmetic."""
    import numpy as np
    s = np.asarray(sizes, dtype=np.float64)
    c = np.concatenate([[0.0], np.cumsum(s)])
    cuts = [0]
    for k in range(1, world):
        target = c[-1] * k / world
        j = int(np.searchsorted(c, target, side="left"))
        if j > 0 and abs(c[j - 1] - target) <= abs(c[min(j, len(c) - 1)] - target):
            j -= 1
        cuts.append(min(max(j, cuts[-1]), len(s)))
    cuts.append(len(s))
    ranges = [(cuts[k], cuts[k + 1]) for k in range(world)]
    return ranges if rank is None else ranges[rank]


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

    def get_compact(self):
        """The prefetched batch, still in the wire format, on the device (ordered behind its copy on the current stream)."""
        b, ev = self._next
        torch.cuda.current_stream().wait_event(ev)
        for v in b._tensors().values():
            v.record_stream(torch.cuda.current_stream())
        self._next = None
        return b

    def get(self):
        return self.get_compact().expand()


class DesignFeeder(object):
    """BASELINE.json configs[2] (exp_classify.py with the supports rebuilt on the GPU for every batch): designs the NEXT raw
    batches (``synthetic.design_and_collate``: SpectralDesign of all graphs of a batch in one launch + collation) on side streams
    while the current batch trains on the main stream.  At the reference's batch size (50 graphs) the one-block-per-graph Jacobi
    is latency-bound (1.8 ms for 50 blocks on 148 SMs) and independent of the training step: ``depth`` designs in flight on
    ``depth`` streams overlap with each other and with the step.  The raw tensors handed to ``prefetch`` must already be valid
    on the device (the side streams do not wait for the main stream)."""

    def __init__(self, sd, device, depth=2):
        self.sd, self.device = sd, device
        self.streams = [torch.cuda.Stream(device=device) for _ in range(depth)]
        self._queue = []
        self._k = 0

    def prefetch(self, raw, records=False):
        """``records=False``: a collated device ``Batch`` (``design_and_collate``); ``records=True``: the per-graph records of
        ``synthetic.design_raw`` for ``GraphedTrainer.load_designed`` (collation + padding straight into the captured buffers)."""
        from .synthetic import design_and_collate, design_raw
        st = self.streams[self._k % len(self.streams)]
        self._k += 1
        with torch.cuda.stream(st):
            b = design_raw(raw, self.sd, self.device) if records else design_and_collate(raw, self.sd, self.device)
        ev = torch.cuda.Event()
        ev.record(st)
        self._queue.append((b, ev))

    def pending(self):
        return len(self._queue)

    def get(self):
        b, ev = self._queue.pop(0)
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(ev)
        vals = list(b.values()) + list(b["design"].values()) if isinstance(b, dict) else list(b.__dict__.values())
        for v in vals:
            if isinstance(v, torch.Tensor) and v.is_cuda:
                v.record_stream(cur)
        return b


def pad_batch(hb, n_nodes, n_entries):
    """Host ``Batch`` -> the same batch padded to ``n_nodes`` nodes and ``n_entries`` support entries (static shapes for CUDA
    graph replay).  The padding is exactly neutral: the extra nodes are isolated, carry zero features and form ONE extra
    (dummy) graph at the end whose read-out row the loss never sees (``GraphedTrainer`` slices it off); the extra entries are
    self-loops spread round-robin over the dummy nodes with all-zero supports (a bias-free edge MLP maps them to zero, they
    add 0 to those nodes and receive zero gradients).  ``n_nodes`` must exceed the batch's node count (the dummy graph is
    never empty); ``padded_shapes`` picks shapes that keep the dummy nodes' degree small (a hub row would serialise the
    row-parallel kernels)."""
    import numpy as np
    from .batch import Batch
    N, E = hb.x.size(0), hb.edge_index2.size(1)
    if n_nodes <= N or n_entries < E:
        raise ValueError("pad_batch: need n_nodes > %d and n_entries >= %d" % (N, E))
    B = hb.num_graphs
    x = torch.zeros(n_nodes, hb.x.size(1), dtype=hb.x.dtype)
    x[:N] = hb.x
    ei = torch.empty((2, n_entries), dtype=torch.int64)
    ei[:, :E] = hb.edge_index2
    ei[:, E:] = (N + torch.arange(n_entries - E) % (n_nodes - N)).unsqueeze(0)
    ea = torch.zeros(n_entries, hb.edge_attr2.size(1), dtype=hb.edge_attr2.dtype)
    ea[:E] = hb.edge_attr2
    batch = torch.full((n_nodes,), B, dtype=torch.int64)
    batch[:N] = hb.batch
    gp = torch.cat([hb.graph_ptr.to(torch.int32), torch.tensor([n_nodes], dtype=torch.int32)])
    return Batch(x=x, edge_index2=ei, edge_attr2=ea, batch=batch, num_graphs=B + 1, graph_ptr=gp, y=hb.y, real_graphs=B)


def padded_shapes(batches, max_pad_degree=8):
    """(n_nodes, n_entries) for ``pad_batch`` / ``GraphedTrainer`` covering every batch of the list: the largest entry count,
    and enough dummy nodes that the padding self-loops of the SMALLEST batch give no dummy node more than ``max_pad_degree``."""
    nmax = max(int(b.x.shape[0]) for b in batches)
    emax = max(int(b.edge_index2.shape[1]) for b in batches)
    emin = min(int(b.edge_index2.shape[1]) for b in batches)
    return nmax + 1 + (emax - emin + max_pad_degree - 1) // max_pad_degree, emax


class GraphedTrainer(object):
    """The whole optimisation step -- CSR build of the new batch, forward, SUM loss, backward, (all-reduce,) Adam -- captured
    ONCE into a CUDA graph and replayed per step (the step's ~120 launches cost ~2 ms of host time to
    enqueue, more than the GPU needs at the reference's batch sizes).  Batches must have static shapes (``pad_batch``); a
    step copies the new batch into the captured input buffers (device -> device, or host -> device when given a pinned host
    batch) and replays.  Everything inside is capture-safe: the library never synchronises, workspaces are allocated during
    the warm-up steps, tensor-map descriptors are kernel parameters."""

    def __init__(self, model, example, loss="l1", lr=1e-3, distributed=False, warmup=3):
        from .graph import set_range_check
        set_range_check(False)                         # the range check reads a flag back from the device
        self.model, self.loss_kind = model, loss
        self.distributed = distributed and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        self.params = [p for p in model.parameters()]
        self.opt = torch.optim.Adam(self.params, lr=lr, fused=True, capturable=True)
        dev = self.params[0].device
        self.static = example.to(dev, non_blocking=False) if not example.x.is_cuda else example
        self.real = int(getattr(example, "real_graphs", example.num_graphs))
        self.fields = ("x", "edge_index2", "edge_attr2", "batch", "graph_ptr", "y")
        s = torch.cuda.Stream(device=dev)
        s.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(s):
            for _ in range(warmup):
                self._eager_step()
        torch.cuda.current_stream(dev).wait_stream(s)
        torch.cuda.synchronize(dev)
        # Single process: ONE graph holds the whole step.  Data parallel: the NCCL all-reduce stays outside the capture (a
        # captured collective ties the graph to the communicator's stream ordering and watchdog); the step is two graphs --
        # CSR build + forward + loss + backward + gradient flattening | Adam on views of the reduced bucket -- with the one
        # all-reduce launch between them.
        self.graph = torch.cuda.CUDAGraph()
        self.graph2 = None
        if not self.distributed:
            with torch.cuda.graph(self.graph):
                self.static_loss = self._eager_step()
        else:
            with torch.cuda.graph(self.graph):
                self.static_loss, self.flat = self._fwd_bwd_flat()
            self._grads_from(self.flat)
            self.graph2 = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph2, pool=self.graph.pool()):
                self.opt.step()

    def _grads_from(self, flat):
        off = 0
        for p in self.params:
            p.grad = flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    def _fwd_bwd_flat(self):
        b = self.static.fresh()
        for p in self.params:
            p.grad = None
        out = self.model(b)
        loss = loss_fn(self.loss_kind, out[:self.real], b.y)
        loss.backward()
        grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in self.params]
        return loss.detach(), torch.cat([g.reshape(-1) for g in grads])

    def _eager_step(self):
        if self.distributed:
            loss, flat = self._fwd_bwd_flat()
            dist.all_reduce(flat, op=dist.ReduceOp.SUM)
            self._grads_from(flat)
            self.opt.step()
            return loss
        b = self.static.fresh()
        for p in self.params:
            p.grad = None
        out = self.model(b)
        loss = loss_fn(self.loss_kind, out[:self.real], b.y)
        loss.backward()
        self.opt.step()
        return loss.detach()

    def load(self, batch):
        """Copy a padded batch (device or pinned host, same shapes as the example) into the captured input buffers."""
        for k in self.fields:
            getattr(self.static, k).copy_(getattr(batch, k), non_blocking=True)

    def load_unpadded(self, b):
        """Write a DEVICE batch of any size that fits (N < captured nodes, E <= captured entries, same number of graphs) into the
        captured input buffers and fill the rest with the neutral padding of ``pad_batch`` (device-side, no host sync)."""
        st = self.static
        N, E = b.x.size(0), b.edge_index2.size(1)
        Np, Ep = st.x.size(0), st.edge_index2.size(1)
        if N >= Np or E > Ep or b.num_graphs != self.real:
            raise ValueError("batch (%d nodes, %d entries, %d graphs) does not fit the captured shapes (%d, %d, %d)"
                             % (N, E, b.num_graphs, Np, Ep, self.real))
        st.x[:N].copy_(b.x, non_blocking=True)
        st.x[N:].zero_()
        st.edge_index2[:, :E].copy_(b.edge_index2, non_blocking=True)
        if Ep > E:
            st.edge_index2[:, E:].copy_((N + torch.arange(Ep - E, device=st.x.device) % (Np - N)).unsqueeze(0))
        st.edge_attr2[:E].copy_(b.edge_attr2, non_blocking=True)
        st.edge_attr2[E:].zero_()
        st.batch[:N].copy_(b.batch, non_blocking=True)
        st.batch[N:].fill_(self.real)
        st.graph_ptr[:self.real + 1].copy_(b.graph_ptr, non_blocking=True)
        st.graph_ptr[self.real + 1:].fill_(Np)
        st.y.copy_(b.y, non_blocking=True)

    def load_compact(self, cb):
        """Device ``CompactBatch`` (wire format) -> the captured input buffers, padding included: two library launches
        (``gnnml3_collate``) + the label copy, instead of ~20 index kernels of ``expand`` and ~12 copies of ``load_unpadded``."""
        st = self.static
        N = (cb.xc if cb.xc is not None else cb.x).size(0)
        if N >= st.x.size(0) or cb.el.size(1) > st.edge_index2.size(1) or cb.num_graphs != self.real:
            raise ValueError("batch does not fit the captured shapes")
        cb.expand_into(st)

    def load_designed(self, rec):
        """Per-graph records of a batch whose supports were just designed on the GPU (``synthetic.design_raw``) -> the captured
        input buffers: collation, global ids and the neutral padding in two library launches."""
        from . import ops
        d = rec["design"]
        if rec["num_graphs"] != self.real:
            raise ValueError("batch does not fit the captured shapes")
        ops.collate_device(rec["n32"], d["counts"], d["edge_index2"], d["edge_attr2"], rec["num_graphs"], self.static, x=rec["x"])
        self.static.y.copy_(rec["y"].reshape(self.static.y.shape), non_blocking=True)

    def load_from_dataset(self, dataset, idx_host):
        """Graph ids -> the captured input buffers, collated on the device from an HBM-resident ``DeviceDataset``."""
        dataset.collate_into(idx_host, self.static)

    def step(self, batch=None):
        """One optimisation step; returns the (device, captured) loss tensor -- valid until the next step."""
        if batch is not None:
            self.load(batch)
        self.graph.replay()
        if self.graph2 is not None:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
            self.graph2.replay()
        return self.static_loss
