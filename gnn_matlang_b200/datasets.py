"""PyG-free readers for the raw dataset files the four configured reference scripts start from:

* ``read_graph6``   -- ``dataset/graph8c/raw/graph8c.g6`` and ``dataset/sr25/raw/sr251256.g6``
  (reference: ``libs/utils.py:473-478`` -- ``nx.read_graph6`` + ``to_undirected`` + ``x = ones(n, 1)``, ``y = 0``)
* ``read_exp_pickle`` -- ``dataset/EXP/raw/GRAPHSAT.pkl`` (reference: ``libs/utils.py:440-442``: a pickled list of PyG
  ``Data`` objects; read here without torch_geometric by mapping the pickled classes to a plain record)

Both return a list of ``dict(x, edge_index, y)`` records (numpy / torch on the host), the input ``SpectralDesign`` and
``batch.collate`` take.  Host-side parsing only: integer work, bit-exact by construction; nothing here touches the GPU.
The ``.mat`` schemas of ZINC / counting (``libs/utils.py:240-261, 386-413``) are not implemented: neither file ships with the
reference.
"""
import io
import pickle

import numpy as np
import torch

__all__ = ["read_graph6", "read_exp_pickle"]


def _g6_graph(line):
    """One graph6 record -> (n, dense 0/1 adjacency).  Format (B. McKay, formats.txt): N(n) then the upper triangle
    x(0,1) x(0,2) x(1,2) x(0,3) ... packed big-endian, 6 bits per byte, every byte offset by 63."""
    v = np.frombuffer(line, dtype=np.uint8).astype(np.int64) - 63
    if v.size == 0 or v.min() < 0 or v.max() > 63:
        raise ValueError("not a graph6 record")
    if v[0] < 63:
        n, pos = int(v[0]), 1
    elif v.size >= 4 and v[1] < 63:
        n, pos = int((v[1] << 12) | (v[2] << 6) | v[3]), 4
    elif v.size >= 8:
        n, pos = 0, 8
        for b in v[2:8]:
            n = (n << 6) | int(b)
    else:
        raise ValueError("truncated graph6 size field")
    nbits = n * (n - 1) // 2
    if (v.size - pos) * 6 < nbits:
        raise ValueError("truncated graph6 record: %d bits for n = %d" % ((v.size - pos) * 6, n))
    bits = ((v[pos:, None] >> np.arange(5, -1, -1)[None, :]) & 1).reshape(-1)[:nbits]
    A = np.zeros((n, n), dtype=np.uint8)
    cols, rows = np.triu_indices(n, 1)[::-1] if n > 1 else (np.zeros(0, np.int64), np.zeros(0, np.int64))
    # column-major order of the strict upper triangle: (0,1) (0,2) (1,2) (0,3) ...  == sort pairs (i < j) by (j, i)
    order = np.lexsort((rows, cols)) if n > 1 else np.zeros(0, np.int64)
    i, j = rows[order], cols[order]
    A[i, j] = bits
    A[j, i] = bits
    return n, A


def read_graph6(path_or_bytes, feature_dim=1):
    """List of ``dict(x=ones[n, feature_dim] float32, edge_index=[2, 2m] int64, y=0)``; ``edge_index`` holds both directions
    of every edge, sorted by (source, target) -- what ``to_undirected`` (coalesce) returns in the reference."""
    data = open(path_or_bytes, "rb").read() if isinstance(path_or_bytes, str) else bytes(path_or_bytes)
    out = []
    for line in data.split(b"\n"):
        line = line.strip()
        if line.startswith(b">>graph6<<"):
            line = line[len(b">>graph6<<"):]
        if not line:
            continue
        n, A = _g6_graph(line)
        r, c = np.nonzero(A)                                           # row-major: sorted by (source, target)
        out.append(dict(x=np.ones((n, feature_dim), np.float32), edge_index=np.vstack((r, c)).astype(np.int64), y=0))
    return out


class _Record(object):
    """Stand-in for the pickled ``torch_geometric.data.Data``: keeps whatever attributes the pickle restores."""

    def __setstate__(self, state):
        self.__dict__.update(state if isinstance(state, dict) else {})


class _NoPyGUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module.split(".")[0] == "torch_geometric":
            return _Record
        return super().find_class(module, name)


def read_exp_pickle(path_or_bytes):
    """``GRAPHSAT.pkl`` (EXP, 1200 planar SAT pair graphs) -> list of ``dict(x=[n, 1] int64 node type, edge_index=[2, e] int64,
    y=int)`` in file order; no torch_geometric import (the pickled ``Data`` class is mapped to a plain record).  Only unpickle
    files you trust -- like the reference, this executes the pickle stream."""
    f = open(path_or_bytes, "rb") if isinstance(path_or_bytes, str) else io.BytesIO(bytes(path_or_bytes))
    with f:
        items = _NoPyGUnpickler(f).load()
    out = []
    for d in items:
        a = d.__dict__
        if "x" not in a and "_store" in a:                            # newer PyG layouts keep the tensors in a storage object
            a = getattr(a["_store"], "__dict__", {}).get("_mapping", a)
        x, ei, y = a["x"], a["edge_index"], a.get("y", 0)
        x = x if isinstance(x, torch.Tensor) else torch.as_tensor(x)
        ei = ei if isinstance(ei, torch.Tensor) else torch.as_tensor(ei)
        out.append(dict(x=x.reshape(x.shape[0], -1).to(torch.int64).numpy(), edge_index=ei.to(torch.int64).numpy(),
                        y=int(torch.as_tensor(y).reshape(-1)[0]) if y is not None else 0))
    return out
