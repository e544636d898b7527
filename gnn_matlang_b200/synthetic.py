"""Synthetic graph generators shaped like the reference's datasets -- the real ZINC /
subgraph-counting .mat files are not shipped with the reference (``.MISSING_LARGE_BLOBS``), so bench.py and
the tests train on graphs of the same shape:

* ``zinc``      -- molecule-like: random tree with max degree 4 plus 1-3 ring closures, n ~ clipped
                   N(23.2, 4.3) in [9, 37]; x = 21-way atom one-hot + 4-way degree one-hot
                   (libs/utils.py:253-259); supports over the 2-hop mask (Zinc12k.py:12: recfield=2, nfreq=7 -> K=8)
* ``counting``  -- random regular graphs (n, d) in {(10,6), (15,6), (20,5), (30,5)}; x = [1, deg/max];
                   1-hop mask, K = 12 (counting.py:16: nfreq=10, addadj)
* ``sweep``     -- G(n, 4/(n-1)), n ~ U{30..100}; 1-hop or 2-hop mask; K = 10

A pool of distinct graphs is generated once; batches are drawn from the pool and collated with vectorised
numpy (same result as ``batch.collate`` on the drawn records).
"""
import numpy as np
import torch

from .batch import Batch


def _sym_sorted(u, v):
    ei = np.concatenate([np.vstack((u, v)), np.vstack((v, u))], 1)
    order = np.lexsort((ei[1], ei[0]))
    return ei[:, order].astype(np.int64)


def zinc_graph(rng):
    n = int(np.clip(round(rng.normal(23.2, 4.3)), 9, 37))
    deg = np.zeros(n, dtype=np.int64)
    us, vs = [], []
    for v in range(1, n):
        lo = max(0, v - 6)
        cand = np.arange(lo, v)
        cand = cand[deg[cand] < 3]
        if len(cand) == 0:
            cand = np.arange(0, v)
            cand = cand[deg[cand] < 4]
        u = int(cand[rng.integers(len(cand))])
        us.append(u)
        vs.append(v)
        deg[u] += 1
        deg[v] += 1
    have = set(zip(us, vs))
    for _ in range(int(rng.integers(1, 4))):
        for _try in range(10):
            u, v = sorted(int(t) for t in rng.integers(0, n, 2))
            if 2 <= v - u <= 6 and (u, v) not in have and deg[u] < 4 and deg[v] < 4:
                have.add((u, v))
                us.append(u)
                vs.append(v)
                deg[u] += 1
                deg[v] += 1
                break
    ei = _sym_sorted(np.array(us), np.array(vs))
    x = np.zeros((n, 25), np.float32)
    x[np.arange(n), rng.integers(0, 21, n)] = 1
    x[np.arange(n), 25 - np.clip(deg, 1, 4)] = 1
    return n, ei, x


def regular_graph(rng, n, d):
    while True:
        stubs = np.repeat(np.arange(n), d)
        rng.shuffle(stubs)
        a, b = stubs[0::2], stubs[1::2]
        if np.any(a == b):
            continue
        key = np.minimum(a, b) * n + np.maximum(a, b)
        if len(np.unique(key)) != len(key):
            continue
        return _sym_sorted(a, b)


def counting_graph(rng):
    n, d = [(10, 6), (15, 6), (20, 5), (30, 5)][int(rng.integers(4))]
    ei = regular_graph(rng, n, d)
    x = np.ones((n, 2), np.float32)
    x[:, 1] = np.bincount(ei[0], minlength=n) / 6.0
    return n, ei, x


def sweep_graph(rng, nfeat):
    n = int(rng.integers(30, 101))
    up = np.triu(rng.random((n, n)) < 4.0 / (n - 1), 1)
    u, v = np.where(up)
    ei = _sym_sorted(u, v)
    return n, ei, rng.standard_normal((n, nfeat)).astype(np.float32)


def mask_edges(n, ei, recfield):
    """Row-major coordinates of the receptive-field mask (libs/utils.py:566-573, :608): recfield 0 -> A,
    r >= 1 -> (A + I)^(2^(r-1)) > 0."""
    A = np.zeros((n, n), dtype=bool)
    A[ei[0], ei[1]] = True
    if recfield == 0:
        M = A
    else:
        M = A | np.eye(n, dtype=bool)
        for _ in range(1, recfield):
            M = (M.astype(np.float32) @ M.astype(np.float32)) > 0
    r, c = np.where(M)
    return np.vstack((r, c)).astype(np.int64)


class GraphPool(object):
    """``count`` distinct graphs stored as flat arrays; ``draw(rng, B)`` collates a host Batch of B graphs."""

    def __init__(self, kind, count, seed=0, K=None, nfeat=None, recfield=None, supports="normal"):
        rng = np.random.default_rng(seed)
        self.kind = kind
        defaults = dict(zinc=(8, 25, 2), counting=(12, 2, 1), sweep=(10, 64, 1))[kind]
        self.K = K or defaults[0]
        self.F = nfeat or defaults[1]
        self.recfield = defaults[2] if recfield is None else recfield
        ns, xs, eis, eas, ys = [], [], [], [], []
        for _ in range(count):
            if kind == "zinc":
                n, ei, x = zinc_graph(rng)
            elif kind == "counting":
                n, ei, x = counting_graph(rng)
            else:
                n, ei, x = sweep_graph(rng, self.F)
            ei2 = mask_edges(n, ei, self.recfield)
            ns.append(n)
            xs.append(x)
            eis.append(ei2)
            ys.append(rng.standard_normal())
        self.n = np.array(ns, dtype=np.int64)
        self.e = np.array([e.shape[1] for e in eis], dtype=np.int64)
        self.node_off = np.concatenate([[0], np.cumsum(self.n)])
        self.edge_off = np.concatenate([[0], np.cumsum(self.e)])
        self.x = np.concatenate(xs, 0)
        self.ei2 = np.concatenate(eis, 1)
        self.y = np.array(ys, dtype=np.float32)
        # edge features: N(0,1) placeholders unless real supports are attached with set_supports()
        self.ea2 = rng.standard_normal((self.ei2.shape[1], self.K)).astype(np.float32)
        self.supports = supports

    def set_supports(self, ea2):
        assert ea2.shape == self.ea2.shape
        self.ea2 = np.ascontiguousarray(ea2, dtype=np.float32)
        self.supports = "spectral_design"

    def draw(self, rng, B):
        idx = rng.integers(0, len(self.n), B)
        return self.collate(idx)

    def collate(self, idx):
        idx = np.asarray(idx)
        n, e = self.n[idx], self.e[idx]
        goff = np.concatenate([[0], np.cumsum(n)])
        eoff = np.concatenate([[0], np.cumsum(e)])
        # vectorised range gather: node rows
        nsel = np.repeat(self.node_off[idx] - goff[:-1], n) + np.arange(goff[-1])
        esel = np.repeat(self.edge_off[idx] - eoff[:-1], e) + np.arange(eoff[-1])
        x = torch.from_numpy(self.x[nsel])
        ei = torch.from_numpy(self.ei2[:, esel] + np.repeat(goff[:-1], e)[None, :])
        ea = torch.from_numpy(self.ea2[esel])
        batch = torch.from_numpy(np.repeat(np.arange(len(idx), dtype=np.int64), n))
        return Batch(x=x, edge_index2=ei, edge_attr2=ea, batch=batch, num_graphs=len(idx),
                     graph_ptr=torch.from_numpy(goff.astype(np.int32)), y=torch.from_numpy(self.y[idx]).reshape(-1, 1))
