"""Batching of per-graph records into one disjoint-union batch -- the semantics of PyG's
``Batch.from_data_list`` for the attributes GNNML3 reads (``DataLoader`` at graph8c.py:18, Zinc12k.py:20-22,
exp_classify.py:19-21, counting.py:29-31).  Integer work, bit-exact with the reference:

* ``x``, ``edge_attr2``, ``y`` are concatenated along dim 0;
* ``edge_index2`` (name contains "index") is concatenated along the last dim and shifted by the running
  node count, so the batched edge list stays sorted by (src, dst);
* ``batch[n]`` = index of the graph that owns node n; graphs stay contiguous and in input order.

In addition the batch carries ``graph_ptr`` (int32 node offsets) so that the readout and the CSR build never
have to rediscover the block structure.
"""
import numpy as np
import torch


class Batch(object):
    """Attribute bag with the reference's field names (``x``, ``edge_index2``, ``edge_attr2``, ``batch``, ``y``)."""

    def __init__(self, **kw):
        self.__dict__.update(kw)

    def to(self, device, non_blocking=True):
        out = Batch()
        for k, v in self.__dict__.items():
            out.__dict__[k] = v.to(device, non_blocking=non_blocking) if isinstance(v, torch.Tensor) else v
        if getattr(out, "batch", None) is not None and getattr(out, "graph_ptr", None) is not None and out.batch.is_cuda:
            out.batch._gnnml3_ptr = (out.batch._version, out.graph_ptr)
        return out

    def fresh(self):
        """New tensor objects over the same storage: drops the per-tensor caches (graph plan, sorted edge
        features), i.e. what a training loop sees when every step brings a new batch."""
        out = Batch()
        for k, v in self.__dict__.items():
            out.__dict__[k] = v.detach() if isinstance(v, torch.Tensor) else v
        if getattr(out, "batch", None) is not None and getattr(out, "graph_ptr", None) is not None and out.batch.is_cuda:
            out.batch._gnnml3_ptr = (out.batch._version, out.graph_ptr)
        return out

    def pin_memory(self):
        out = Batch()
        for k, v in self.__dict__.items():
            out.__dict__[k] = v.pin_memory() if isinstance(v, torch.Tensor) else v
        return out

    def nbytes(self):
        return sum(v.numel() * v.element_size() for v in self.__dict__.values() if isinstance(v, torch.Tensor))


def collate(graphs):
    """``graphs``: sequence of dicts / objects with ``x [n,f]``, ``edge_index2 [2,e]``, ``edge_attr2 [e,K]``
    and optionally ``y``.  Returns a host ``Batch``."""
    def get(g, k):
        return g.get(k) if isinstance(g, dict) else getattr(g, k, None)

    ns = np.array([int(get(g, "x").shape[0]) for g in graphs], dtype=np.int64)
    off = np.zeros(len(graphs) + 1, dtype=np.int64)
    np.cumsum(ns, out=off[1:])
    x = torch.cat([torch.as_tensor(get(g, "x"), dtype=torch.float32) for g in graphs], 0)
    ei = torch.cat([torch.as_tensor(get(g, "edge_index2"), dtype=torch.int64) + int(o) for g, o in zip(graphs, off[:-1])], 1)
    ea = torch.cat([torch.as_tensor(get(g, "edge_attr2"), dtype=torch.float32) for g in graphs], 0)
    batch = torch.from_numpy(np.repeat(np.arange(len(graphs), dtype=np.int64), ns))
    out = Batch(x=x, edge_index2=ei, edge_attr2=ea, batch=batch, num_graphs=len(graphs),
                graph_ptr=torch.from_numpy(off.astype(np.int32)))
    ys = [get(g, "y") for g in graphs]
    if all(y is not None for y in ys) and len(ys) > 0:
        ys = [torch.as_tensor(y) for y in ys]
        out.y = torch.cat([y.reshape(1, -1) if y.dim() < 2 else y for y in ys], 0)
    return out
