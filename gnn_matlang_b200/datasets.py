"""PyG-free readers for the raw dataset files the four configured reference scripts start from:

* ``read_graph6``   -- ``dataset/graph8c/raw/graph8c.g6`` and ``dataset/sr25/raw/sr251256.g6``
  (reference: ``libs/utils.py:473-478`` -- ``nx.read_graph6`` + ``to_undirected`` + ``x = ones(n, 1)``, ``y = 0``)
* ``read_exp_pickle`` -- ``dataset/EXP/raw/GRAPHSAT.pkl`` (reference: ``libs/utils.py:440-442``: a pickled list of PyG
  ``Data`` objects; read here without torch_geometric by mapping the pickled classes to a plain record)

* ``read_mat``      -- the MATLAB files of the other GNNML3 scripts: ``Zinc.mat`` (``libs/utils.py:240-261``: atom type |
  degree code one-hot features, nmax 37), ``randomgraph.mat`` of the counting task (``:386-413``: the five substructure
  counts recomputed from the adjacency exactly as the reference does), and the TU files that do ship with the reference,
  ``enzymes.mat`` / ``proteins.mat`` (``:93-112, 146-165``: first three feature columns unless ``contfeat``), ``ptc.mat``
  (``:46-61``), ``mutag.mat`` (``:195-211``: labels mapped from {-1, 1} to {0, 1})

All return a list of ``dict(x, edge_index, y)`` records (numpy on the host), the input ``SpectralDesign`` and
``batch.collate`` take.  Host-side parsing only: integer work, bit-exact by construction; nothing here touches the GPU.
"""
import io
import pickle

import numpy as np
import torch

__all__ = ["read_graph6", "read_exp_pickle", "read_mat"]


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


def _edges_of(A):
    """``np.where(A > 0)`` of a dense (or scipy-sparse) adjacency: row-major order = sorted by (source, target)."""
    A = A.toarray() if hasattr(A, "toarray") else np.asarray(A)
    r, c = np.where(A > 0)
    return A, np.vstack((r, c)).astype(np.int64)


def _comb3(d):
    d = int(d)
    return d * (d - 1) * (d - 2) // 6 if d >= 3 else 0


def read_mat(path, kind, contfeat=False):
    """Records of one of the reference's MATLAB datasets (``scipy.io.loadmat``); ``kind`` in
    {"zinc", "counting", "enzymes", "proteins", "ptc", "mutag"}.  Follows the ``process()`` of the reference's dataset class line
    by line (see the module docstring for the line ranges); ``y`` keeps the reference's dtype (int64 class ids, float32 for
    ZINC / mutag, float64 counts for the counting task)."""
    import scipy.io as sio
    a = sio.loadmat(path)
    out = []
    if kind == "zinc":                                   # Zinc12KDataset.process
        F, E, Y = a["F"][0], a["E"][0], a["Y"]
        ntype, maxdeg = 21, 4
        for i in range(len(E)):
            A, ei = _edges_of(E[i])
            n = A.shape[0]
            x = np.zeros((n, ntype + maxdeg), np.float32)
            deg = (A > 0).sum(1)
            codes = np.asarray(F[i][0]).reshape(-1)
            for j in range(codes.shape[0]):
                x[j, int(codes[j])] = 1                  # atom code
                x[j, -int(deg[j])] = 1                   # degree code (from the end; degree 0 lands on column 0 as in the reference)
            out.append(dict(x=x, edge_index=ei, y=np.asarray(Y[i, :])))
    elif kind == "counting":                             # GraphCountDataset.process
        As = a["A"][0]
        for i in range(len(As)):
            A, ei = _edges_of(As[i])
            A = A.astype(np.float64)
            A2 = A.dot(A)
            A3 = A2.dot(A)
            tri = np.trace(A3) / 6
            tailed = ((np.diag(A3) / 2) * (A.sum(0) - 2)).sum()
            cyc4 = 1 / 8 * (np.trace(A3.dot(A)) + np.trace(A2) - 2 * A2.sum())
            cus = A.dot(np.diag(np.exp(-A.dot(A).sum(1)))).dot(A).sum()
            star = sum(_comb3(d) for d in A.sum(0))
            out.append(dict(x=np.ones((A.shape[0], 1), np.float32), edge_index=ei, y=np.array([[tri, tailed, star, cyc4, cus]])))
    elif kind in ("enzymes", "proteins", "ptc"):
        As, F = a["A"][0], a["F"][0]
        Y = a["Y"][0].astype(np.int64) if kind == "enzymes" else a["Y"].astype(np.int64)[:, 0]
        for i in range(len(As)):
            _, ei = _edges_of(As[i])
            f = np.asarray(F[i])
            x = (f if (contfeat or kind == "ptc") else f[:, 0:3]).astype(np.float32)
            out.append(dict(x=x, edge_index=ei, y=np.array([Y[i]], np.int64)))
    elif kind == "mutag":
        As, F = a["A"][0], a["F"][0]
        Y = ((a["y"] + 1) // 2).astype(np.float32)
        for i in range(len(As)):
            _, ei = _edges_of(As[i])
            out.append(dict(x=np.asarray(F[i]).astype(np.float32), edge_index=ei, y=Y[i].astype(np.float32)))
    else:
        raise ValueError("read_mat: unknown kind %r" % (kind,))
    return out
