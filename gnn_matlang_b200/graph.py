"""GraphPlan: the per-batch index structures shared by every layer of a model.

The reference re-derives the gather/scatter indexing K times per layer inside PyG's propagate
(libs/spect_conv.py:77).  Here the batched ``edge_index2`` is turned ONCE into a dst-sorted CSR plus the
transposed (src-sorted) CSR (gnnml3_csr_build); the plan is cached on the ``edge_index`` tensor object, so
the four or five ML3Layers of a model -- which all receive the same ``edge_index2`` -- share it, and the
(unlearned) ``edge_attr2`` is permuted into dst-sorted order once per batch as well.
"""
import torch

from . import ops


_RANGE_CHECK = [True]


def set_range_check(flag):
    """The index range check reads one flag back from the device (a host sync per new batch); trusted
    pipelines (our own collation) switch it off."""
    _RANGE_CHECK[0] = bool(flag)


class GraphPlan(object):
    __slots__ = ("N", "E", "rowptr", "col", "perm", "rowptrT", "colT", "permT", "win", "winT", "device")

    def __init__(self, edge_index, num_nodes, check_range=True):
        d = ops.csr_build(edge_index, num_nodes, check_range=check_range)
        self.N, self.E = int(num_nodes), int(edge_index.size(1))
        self.device = edge_index.device
        for k, v in d.items():
            setattr(self, k, v)


def get_plan(edge_index, num_nodes):
    """Plan for ``edge_index`` (cached on the tensor object; invalidated by in-place modification)."""
    cached = getattr(edge_index, "_gnnml3_plan", None)
    if cached is not None and cached[0] == edge_index._version and cached[1].N == int(num_nodes):
        return cached[1]
    plan = GraphPlan(edge_index, num_nodes, check_range=_RANGE_CHECK[0])
    try:
        edge_index._gnnml3_plan = (edge_index._version, plan)
    except Exception:  # pragma: no cover - tensors that refuse attributes just rebuild each call
        pass
    return plan


class _SortEdgeAttr(torch.autograd.Function):
    """edge_attr [E,K] (original edge order) -> dst-sorted order; backward scatters the gradient back."""

    @staticmethod
    def forward(ctx, edge_attr, plan):
        ctx.plan = plan
        return ops.gather_rows(edge_attr, plan.perm)

    @staticmethod
    def backward(ctx, g):
        return ops.scatter_rows(g.contiguous(), ctx.plan.perm), None


def sorted_edge_attr(edge_attr, plan):
    """``edge_attr[plan.perm]``; cached on the tensor object when no gradient is required (the model feeds
    the same ``edge_attr2`` to every layer -- Zinc12k.py:338-341)."""
    if edge_attr.dim() != 2 or edge_attr.size(0) != plan.E:
        raise RuntimeError("edge_attr must be [E, K] with E == edge_index.size(1) (got %s, E=%d)"
                           % (tuple(edge_attr.shape), plan.E))
    if edge_attr.requires_grad and torch.is_grad_enabled():
        return _SortEdgeAttr.apply(edge_attr, plan)
    cached = getattr(edge_attr, "_gnnml3_sorted", None)
    if cached is not None and cached[0] == edge_attr._version and cached[1] is plan:
        return cached[2]
    out = ops.gather_rows(edge_attr.detach(), plan.perm)
    try:
        edge_attr._gnnml3_sorted = (edge_attr._version, plan, out)
    except Exception:  # pragma: no cover
        pass
    return out
