"""TEST INFRASTRUCTURE ONLY -- in-memory stand-in for the PyG 1.6.1 names the reference imports.

The reference (balcilar/gnn-matlang) pins torch_geometric==1.6.1 (README.md:9); PyG is not vendored,
not installed here and cannot be installed (no network).  This module registers fake
``torch_geometric.*`` modules so that ``/root/reference/libs/spect_conv.py`` and
``/root/reference/libs/utils.py`` can be imported *unmodified* in THIS container, which is how
``oracle/make_golden.py`` produces the committed fixtures under ``tests/golden/``.

It is never imported by the product package and never used on the GPU box (``/root/reference`` does
not exist there).  Semantics restated (PyG 1.6.1, ``flow='source_to_target'``, ``aggr='add'``):

* ``MessagePassing.propagate(edge_index, x=..., norm=..., size=None)``:
  ``x_j = x.index_select(0, edge_index[0])``; ``msg = self.message(x_j, norm)``;
  ``out = zeros(x.size(0), F).index_add_(0, edge_index[1], msg)``  (scatter-add over dim 0 with
  ``dim_size = x.size(0)``; sequential in edge order on CPU).
* ``Data``: attribute bag.  ``InMemoryDataset``: unused base class for the dataset parsers.
"""
import sys
import types

import torch


class MessagePassing(torch.nn.Module):
    def __init__(self, aggr="add", flow="source_to_target", node_dim=0, **kwargs):
        super().__init__()
        assert aggr == "add" and flow == "source_to_target"
        self.aggr = aggr

    def propagate(self, edge_index, size=None, **kwargs):
        x = kwargs["x"]
        norm = kwargs["norm"]
        x_j = x.index_select(0, edge_index[0])
        msg = self.message(x_j, norm)
        out = torch.zeros(x.size(0), msg.size(1), dtype=msg.dtype, device=msg.device)
        return out.index_add(0, edge_index[1], msg)


class Data(object):
    def __init__(self, **kwargs):
        for k, v in kwargs.items():
            setattr(self, k, v)


class InMemoryDataset(object):
    pass


def _unused(*a, **k):
    raise NotImplementedError("not needed on the GNNML3 hot path")


def install():
    """Register the fake modules (idempotent)."""
    if "torch_geometric" in sys.modules and getattr(sys.modules["torch_geometric"], "_gnnml3_stub", False):
        return
    names = ["torch_geometric", "torch_geometric.data", "torch_geometric.data.data", "torch_geometric.utils",
             "torch_geometric.typing", "torch_geometric.nn", "torch_geometric.nn.conv"]
    mods = {n: types.ModuleType(n) for n in names}
    mods["torch_geometric"]._gnnml3_stub = True
    mods["torch_geometric.data"].InMemoryDataset = InMemoryDataset
    mods["torch_geometric.data"].Data = Data
    mods["torch_geometric.data"].data = mods["torch_geometric.data.data"]
    mods["torch_geometric.data.data"].Data = Data
    for n in ("to_networkx", "to_undirected", "remove_self_loops", "add_self_loops", "get_laplacian"):
        setattr(mods["torch_geometric.utils"], n, _unused)
    mods["torch_geometric.typing"].OptTensor = None
    mods["torch_geometric.nn"].conv = mods["torch_geometric.nn.conv"]
    mods["torch_geometric.nn.conv"].MessagePassing = MessagePassing
    mods["torch_geometric"].data = mods["torch_geometric.data"]
    mods["torch_geometric"].utils = mods["torch_geometric.utils"]
    mods["torch_geometric"].typing = mods["torch_geometric.typing"]
    mods["torch_geometric"].nn = mods["torch_geometric.nn"]
    sys.modules.update(mods)


def import_reference(ref_root="/root/reference"):
    """Import the reference's two library files verbatim; returns (spect_conv_module, utils_module)."""
    import importlib.util
    import os
    install()
    out = []
    for name in ("spect_conv", "utils"):
        path = os.path.join(ref_root, "libs", name + ".py")
        spec = importlib.util.spec_from_file_location("_gnnml3_ref_" + name, path)
        mod = importlib.util.module_from_spec(spec)
        sys.dont_write_bytecode = True
        spec.loader.exec_module(mod)
        out.append(mod)
    return tuple(out)
