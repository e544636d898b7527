"""Graph readout: drop-ins for PyG ``global_add_pool`` / ``global_mean_pool`` as used by the reference's
GNNML3 wrappers (graph8c.py:277, Zinc12k.py:343, counting.py:370, exp_classify.py:293).

The nodes of a graph are contiguous in a PyG batch, so the readout is a segmented sum over node ranges
(``graph_ptr``) -- one kernel, no atomics, deterministic -- instead of torch_scatter's atomic scatter."""
import torch

from . import ops


class _SegmentPoolFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, graph_ptr, mean):
        ctx.graph_ptr, ctx.mean, ctx.N = graph_ptr, mean, x.size(0)
        return ops.segment_pool_fwd(x, graph_ptr, mean)

    @staticmethod
    def backward(ctx, g):
        return ops.segment_pool_bwd(g.contiguous(), ctx.graph_ptr, ctx.mean, ctx.N), None, None


class _SegmentMaxFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, graph_ptr):
        out, arg = ops.segment_max_fwd(x, graph_ptr)
        ctx.graph_ptr, ctx.N = graph_ptr, x.size(0)
        ctx.save_for_backward(arg)
        return out

    @staticmethod
    def backward(ctx, g):
        return ops.segment_max_bwd(g.contiguous(), ctx.saved_tensors[0], ctx.graph_ptr, ctx.N), None


def graph_ptr_from_batch(batch, size=None):
    """``batch`` [N] (sorted graph id per node) -> int32 ``graph_ptr`` [B+1]; cached on the tensor object."""
    cached = getattr(batch, "_gnnml3_ptr", None)
    if cached is not None and cached[0] == batch._version and (size is None or cached[1].numel() == int(size) + 1):
        return cached[1]
    if not batch.is_cuda:
        raise RuntimeError("gnn_matlang_b200: batch must be a CUDA tensor (no CPU fallback)")
    if batch.numel() > 1 and bool((batch[1:] < batch[:-1]).any()):
        raise RuntimeError("global_*_pool: the batch vector must be sorted (PyG batches are)")
    B = int(size) if size is not None else (int(batch.max()) + 1 if batch.numel() else 0)
    ptr = torch.searchsorted(batch.contiguous(), torch.arange(B + 1, device=batch.device, dtype=batch.dtype)).to(torch.int32)
    try:
        batch._gnnml3_ptr = (batch._version, ptr)
    except Exception:  # pragma: no cover
        pass
    return ptr


def global_add_pool(x, batch, size=None):
    return _SegmentPoolFn.apply(x, graph_ptr_from_batch(batch, size), False)


def global_mean_pool(x, batch, size=None):
    return _SegmentPoolFn.apply(x, graph_ptr_from_batch(batch, size), True)


def global_max_pool(x, batch, size=None):
    """PyG ``global_max_pool`` (enzymes.py:340,384; ptc.py): per-graph maximum, gradient to the node that attains it."""
    return _SegmentMaxFn.apply(x, graph_ptr_from_batch(batch, size))
