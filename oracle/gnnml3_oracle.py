"""TEST INFRASTRUCTURE ONLY -- CPU restatement ("oracle") of the GNNML3 hot path of balcilar/gnn-matlang.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import this file, and only as the checker / the timed CPU baseline -- never as part of the product path.
The product package ``gnn_matlang_b200`` must not import anything from ``oracle/``.

Parity pinning: the reference ships no tests or golden vectors, so this restatement is
pinned against the reference's OWN code run in the build container: ``oracle/make_golden.py`` imports
``/root/reference/libs/spect_conv.py`` and ``/root/reference/libs/utils.py`` verbatim (behind
``oracle/pyg_stub.py``), runs them on seeded inputs and commits inputs+outputs under ``tests/golden/``;
``tests/test_oracle_golden.py`` checks every function below against those fixtures.  The PyG 1.6.1 /
torch_scatter semantics at the boundary (``propagate``, ``Batch.from_data_list``, ``global_*_pool``) are
third-party code absent from ``/root/reference`` and are restated from their published behaviour; for
those three the fixtures pin this file against the stand-in, not against PyG itself ("parity unpinned"
for the PyG collation/pool semantics proper; pinned for everything in libs/).

All citations are relative to /root/reference.  Arithmetic is FP32 (torch CPU) for the layer path and
float64/float32 numpy for SpectralDesign, exactly as in the reference.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------------------
# PyG boundary semantics (third-party, torch_geometric==1.6.1 pinned by README.md:9; not vendored)
# --------------------------------------------------------------------------------------------------


def propagate_add(x, edge_index, norm):
    """``MessagePassing.propagate`` with ``aggr='add'``, ``flow='source_to_target'`` and the reference's
    ``message`` (libs/spect_conv.py:98-99): out[t] = sum_{e: edge_index[1,e]==t} norm[e] * x[edge_index[0,e]].
    CPU index_add_ accumulates sequentially in edge order (SURVEY.md hard part 8)."""
    x_j = x.index_select(0, edge_index[0])
    msg = norm.view(-1, 1) * x_j
    out = torch.zeros(x.size(0), x.size(1), dtype=x.dtype)
    return out.index_add(0, edge_index[1], msg)


def global_add_pool(x, batch, num_graphs):
    """PyG ``global_add_pool`` (graph8c.py:277, Zinc12k.py:343, counting.py:370): per-graph sum."""
    out = torch.zeros(num_graphs, x.size(1), dtype=x.dtype)
    return out.index_add(0, batch, x)


def global_mean_pool(x, batch, num_graphs):
    """PyG ``global_mean_pool`` (exp_classify.py:293): per-graph mean (count clamped to >= 1)."""
    s = global_add_pool(x, batch, num_graphs)
    cnt = torch.bincount(batch, minlength=num_graphs).clamp(min=1).to(x.dtype)
    return s / cnt.view(-1, 1)


def global_max_pool(x, batch, num_graphs):
    """PyG global_max_pool (torch_scatter 'max', third-party -- restated): per-graph maximum; an empty graph gives 0."""
    out = torch.zeros(num_graphs, x.size(1), dtype=x.dtype)
    for b in range(num_graphs):
        m = batch == b
        if bool(m.any()):
            out[b] = x[m].max(0).values
    return out


class OracleGNNML3Variant(torch.nn.Module):
    """The GNNML3 variants of the TU-dataset scripts: ``enzymes`` (enzymes.py:345-386: 4 x ML3Layer(64||0, learnedge=False),
    dropout 0.1 before every layer, add||max read-out, BatchNorm1d, log_softmax(fc2)) and ``mutag`` (mutag.py:268-307:
    3 x ML3Layer(24||24, learnedge=False) + BatchNorm1d each, mean read-out, fc2(relu(fc1)))."""

    def __init__(self, config, ne, ninp):
        super().__init__()
        self.config = config
        if config == "enzymes":
            nout1, nout2, nl = 64, 0, 4
        elif config == "filtering":                    # filtering.py:252-281: 3 x ML3Layer(32||16, learnedge=False), fc2 per node
            nout1, nout2, nl = 32, 16, 3
        else:
            nout1, nout2, nl = 24, 24, 3
        nin = nout1 + nout2
        self.nl = nl
        for l in range(nl):
            setattr(self, "conv%d" % (l + 1), OracleML3Layer(False, ne, ne, ninp if l == 0 else nin, nout1, nout2))
        if config == "enzymes":
            self.bn4 = torch.nn.BatchNorm1d(2 * nin)
            self.fc2 = torch.nn.Linear(2 * nin, 6)
        elif config == "filtering":
            self.fc2 = torch.nn.Linear(nin, 1)
        else:
            for l in range(nl):
                setattr(self, "bn%d" % (l + 1), torch.nn.BatchNorm1d(nin))
            self.fc1 = torch.nn.Linear(nin, 32)
            self.fc2 = torch.nn.Linear(32, 1)

    def forward(self, b):
        x, ei, ea = b["x"], b["edge_index2"], b["edge_attr2"]
        for l in range(self.nl):
            if self.config == "enzymes":
                x = F.dropout(x, p=0.1, training=self.training)
            x = getattr(self, "conv%d" % (l + 1))(x, ei, ea)
            if self.config == "mutag":
                x = getattr(self, "bn%d" % (l + 1))(x)
        if self.config == "filtering":
            return self.fc2(x)
        if self.config == "enzymes":
            x = torch.cat([global_add_pool(x, b["batch"], b["num_graphs"]), global_max_pool(x, b["batch"], b["num_graphs"])], 1)
            return F.log_softmax(self.fc2(self.bn4(x)), dim=1)
        x = global_mean_pool(x, b["batch"], b["num_graphs"])
        return self.fc2(F.relu(self.fc1(x)))


def collate(graphs):
    """PyG ``Batch.from_data_list`` restricted to the attributes GNNML3 reads.

    ``graphs``: list of dicts with ``x [n,f]``, ``edge_index2 [2,e]`` int64, ``edge_attr2 [e,K]`` and
    optionally ``y``.  Attributes whose name contains ``index`` are concatenated along the last dim and
    shifted by the running node count; everything else along dim 0; ``batch`` = graph id per node."""
    xs, eis, eas, bs, ys = [], [], [], [], []
    off = 0
    for g, d in enumerate(graphs):
        n = d["x"].shape[0]
        xs.append(torch.as_tensor(d["x"], dtype=torch.float32))
        eis.append(torch.as_tensor(d["edge_index2"], dtype=torch.int64) + off)
        eas.append(torch.as_tensor(d["edge_attr2"], dtype=torch.float32))
        bs.append(torch.full((n,), g, dtype=torch.int64))
        if "y" in d and d["y"] is not None:
            ys.append(torch.as_tensor(d["y"]).reshape(1, -1) if torch.as_tensor(d["y"]).dim() < 2
                      else torch.as_tensor(d["y"]))
        off += n
    out = dict(x=torch.cat(xs, 0), edge_index2=torch.cat(eis, 1), edge_attr2=torch.cat(eas, 0),
               batch=torch.cat(bs, 0), num_graphs=len(graphs))
    if ys:
        out["y"] = torch.cat(ys, 0)
    return out


# --------------------------------------------------------------------------------------------------
# libs/spect_conv.py
# --------------------------------------------------------------------------------------------------


def glorot_(t):
    """libs/spect_conv.py:13-16 -- U(-s, s), s = sqrt(6 / (size(-2) + size(-1)))."""
    s = math.sqrt(6.0 / (t.size(-2) + t.size(-1)))
    with torch.no_grad():
        t.uniform_(-s, s)
    return t


def spectconv_forward(x, edge_index, edge_attr, weight, bias=None, selfconn=False, depthwise=False,
                      DSweight=None):
    """libs/spect_conv.py:64-96.

    default:   out = [x @ W[-1] if selfconn] + sum_{k<K} P_k(x) @ W[k]  (+ bias)        (:70-80, :93-94)
    depthwise: out = ([x * DS[-1] if selfconn] + (1 + DS[0]) * P_0(x) + sum_{k>=1} DS[k] * P_k(x)) @ W[0]
               (+ bias)                                                                  (:81-91)
    with P_k(x) = propagate_add(x, edge_index, edge_attr[:, k])."""
    out = 0
    if not depthwise:
        nk = weight.size(0)
        if selfconn:
            out = x @ weight[-1]
            nk -= 1
        for k in range(nk):
            out = out + propagate_add(x, edge_index, edge_attr[:, k]) @ weight[k]
    else:
        nk = DSweight.size(0)
        if selfconn:
            out = x * DSweight[-1]
            nk -= 1
        out = out + (1 + DSweight[0:1, :]) * propagate_add(x, edge_index, edge_attr[:, 0])
        for k in range(1, nk):
            out = out + DSweight[k:k + 1, :] * propagate_add(x, edge_index, edge_attr[:, k])
        out = out @ weight[0]
    if bias is not None:
        out = out + bias
    return out


def spectconcatconv_forward(x, edge_index, edge_attr, weight, bias=None, selfconn=True):
    """libs/spect_conv.py:137-158: ``[x W_last (if selfconn) || P_0(x) W_0 || .. ] + bias``  (bias has K' * out entries)."""
    out = []
    enditr = weight.size(0)
    if selfconn:
        out.append(torch.matmul(x, weight[-1]))
        enditr -= 1
    for i in range(enditr):
        out.append(torch.matmul(propagate_add(x, edge_index, edge_attr[:, i]), weight[i]))
    out = torch.cat(out, 1)
    if bias is not None:
        out = out + bias
    return out


def spectconv_forward_dense(x, edge_index, edge_attr, weight, bias=None):
    """Independent dense statement of the default branch, after libs/layers_tf.py:222-245
    (``sum_k (S_k @ X) @ W_k`` with dense supports): S_k[t, s] = sum of edge_attr[e, k] over edges s->t."""
    n = x.size(0)
    out = torch.zeros(n, weight.size(2), dtype=torch.float64)
    for k in range(edge_attr.size(1)):
        S = torch.zeros(n, n, dtype=torch.float64)
        S.index_put_((edge_index[1], edge_index[0]), edge_attr[:, k].double(), accumulate=True)
        out += (S @ x.double()) @ weight[k].double()
    if bias is not None:
        out += bias.double()
    return out


def edge_mlp_forward(edge_attr, w1, w2, w3, w4):
    """libs/spect_conv.py:206-207: relu(fc1_4([relu(fc1_1 ea) || tanh(fc1_2 ea) * tanh(fc1_3 ea)])),
    all four Linear layers bias-free (:191-194); ``w*`` are the ``nn.Linear.weight`` tensors [out, in]."""
    tmp = torch.cat([F.relu(edge_attr @ w1.t()), torch.tanh(edge_attr @ w2.t()) * torch.tanh(edge_attr @ w3.t())], 1)
    return F.relu(tmp @ w4.t())


def ml3layer_forward(x, edge_index, edge_attr, p, learnedge=True, nout2=1):
    """libs/spect_conv.py:204-212.  ``p`` maps the reference's state_dict keys to tensors
    (``fc1_1.weight`` .. ``fc1_4.weight``, ``conv1.weight``, ``conv1.bias``, ``fc11.weight``, ``fc11.bias``,
    ``fc12.weight``, ``fc12.bias``)."""
    if learnedge:
        edge_attr = edge_mlp_forward(edge_attr, p["fc1_1.weight"], p["fc1_2.weight"], p["fc1_3.weight"],
                                     p["fc1_4.weight"])
    conv = F.relu(spectconv_forward(x, edge_index, edge_attr, p["conv1.weight"], p.get("conv1.bias")))
    if nout2 > 0:
        gate = torch.tanh(x @ p["fc11.weight"].t() + p["fc11.bias"]) * torch.tanh(x @ p["fc12.weight"].t() + p["fc12.bias"])
        return torch.cat([conv, gate], 1)
    return conv


class OracleML3Layer(torch.nn.Module):
    """Module form of :func:`ml3layer_forward` with the reference's creation (= RNG) order
    (libs/spect_conv.py:184-202) and state_dict keys, used for the timed CPU baseline."""

    def __init__(self, learnedge, nedgeinput, nedgeoutput, ninp, nout1, nout2):
        super().__init__()
        self.learnedge, self.nout2 = learnedge, nout2
        if learnedge:
            self.fc1_1 = torch.nn.Linear(nedgeinput, 2 * nedgeinput, bias=False)
            self.fc1_2 = torch.nn.Linear(nedgeinput, 2 * nedgeinput, bias=False)
            self.fc1_3 = torch.nn.Linear(nedgeinput, 2 * nedgeinput, bias=False)
            self.fc1_4 = torch.nn.Linear(4 * nedgeinput, nedgeoutput, bias=False)
        else:
            nedgeoutput = nedgeinput
        self.conv1 = torch.nn.Module()
        self.conv1.weight = torch.nn.Parameter(glorot_(torch.empty(nedgeoutput, ninp, nout1)))
        self.conv1.bias = torch.nn.Parameter(torch.zeros(nout1))
        if nout2 > 0:
            self.fc11 = torch.nn.Linear(ninp, nout2)
            self.fc12 = torch.nn.Linear(ninp, nout2)

    def forward(self, x, edge_index, edge_attr):
        return ml3layer_forward(x, edge_index, edge_attr, dict(self.named_parameters()), self.learnedge, self.nout2)


# model wrappers: graph8c.py:249-279, Zinc12k.py:310-345, exp_classify.py:264-295, counting.py:335-372
MODEL_CONFIGS = {
    #            layers nout1 nout2 head            pool    final activation
    "graph8c": dict(nlayer=3, nout1=32, nout2=16, head=(10,), pool="add", final="tanh"),
    "zinc": dict(nlayer=4, nout1=30, nout2=2, head=(32, 1), pool="add", final=None),
    "exp": dict(nlayer=3, nout1=32, nout2=16, head=(10, 1), pool="mean", final=None),
    "counting": dict(nlayer=5, nout1=16, nout2=16, head=(32, 1), pool="add", final=None),
}


class OracleGNNML3(torch.nn.Module):
    """The four GNNML3 wrappers of the reference scripts as one parameterised module (same sub-module
    names ``conv1..convL``, ``fc1``, ``fc2`` and creation order)."""

    def __init__(self, config, ne, ninp):
        super().__init__()
        c = MODEL_CONFIGS[config] if isinstance(config, str) else config
        self.c = c
        nin = c["nout1"] + c["nout2"]
        for l in range(c["nlayer"]):
            setattr(self, "conv%d" % (l + 1), OracleML3Layer(True, ne, ne, ninp if l == 0 else nin, c["nout1"], c["nout2"]))
        self.fc1 = torch.nn.Linear(nin, c["head"][0])
        if len(c["head"]) > 1:
            self.fc2 = torch.nn.Linear(c["head"][0], c["head"][1])

    def forward(self, b):
        x = b["x"]
        for l in range(self.c["nlayer"]):
            x = getattr(self, "conv%d" % (l + 1))(x, b["edge_index2"], b["edge_attr2"])
        pool = global_add_pool if self.c["pool"] == "add" else global_mean_pool
        x = pool(x, b["batch"], b["num_graphs"])
        if len(self.c["head"]) == 1:
            x = self.fc1(x)
            return torch.tanh(x) if self.c["final"] == "tanh" else x
        return self.fc2(F.relu(self.fc1(x)))


# --------------------------------------------------------------------------------------------------
# libs/utils.py -- SpectralDesign
# --------------------------------------------------------------------------------------------------


def spectral_design(edge_index, x, recfield=1, dv=5, nfreq=5, adddegree=False, laplacien=True, addadj=False,
                    vmax=None):
    """libs/utils.py:546-610 (the PPGN tensors of :613-624 are not on the GNNML3 path).

    edge_index: int array [2, e]; x: float array [n, f].
    Returns dict(x [n, f(+1)] f32, lmax f32, edge_index2 [2, E] int64 (row-major over the mask),
    edge_attr2 [E, nfreq+1(+1)] f32)."""
    edge_index = np.asarray(edge_index)
    x = np.asarray(x, dtype=np.float32)
    n = x.shape[0]
    A = np.zeros((n, n), dtype=np.float32)
    A[edge_index[0], edge_index[1]] = 1                                     # :558-560
    if adddegree:
        x = np.concatenate([x, A.sum(0)[:, None]], 1)                        # :562-563 (column sums)
    if recfield == 0:                                                       # :566-573
        M = A > 0
    else:
        M = A + np.eye(n)
        for _ in range(1, recfield):
            M = M.dot(M)
        M = M > 0
    d = A.sum(axis=0)                                                       # :576-582
    with np.errstate(divide="ignore", invalid="ignore"):
        dis = 1 / np.sqrt(d)
    dis[~np.isfinite(dis)] = 0
    D = np.diag(dis)
    nL = np.eye(n) - (A.dot(D)).T.dot(D)                                    # f32 products, f64 result
    V, U = np.linalg.eigh(nL)                                               # :583-584
    V[V < 0] = 0
    lmax = V.max().astype(np.float32)                                       # :586
    if not laplacien:                                                       # :588-589 (float32 eigh, unclamped)
        V, U = np.linalg.eigh(A)
    top = V.max() if vmax is None else vmax                                 # :592-596
    centers = np.linspace(V.min(), top, nfreq)
    nsup = nfreq + 1 + (1 if addadj else 0)
    SP = np.zeros((nsup, n, n), dtype=np.float32)
    for i in range(nfreq):                                                  # :599-600
        SP[i] = M * ((U * np.exp(-(dv * (V - centers[i]) ** 2))[None, :]).dot(U.T))
    SP[nfreq] = np.eye(n)                                                   # :602
    if addadj:
        SP[nfreq + 1] = A                                                   # :604-605
    r, c = np.where(M > 0)                                                  # :608-610
    return dict(x=x, lmax=lmax, edge_index2=np.vstack((r, c)).astype(np.int64),
                edge_attr2=np.ascontiguousarray(SP[:, r, c].T).astype(np.float32))


def supports_dense(edge_index2, edge_attr2, n):
    """Scatter ``edge_attr2`` back to dense [K, n, n] supports: the representation in which supports are
    compared (invariant to eigenvector sign / rotation inside degenerate eigenspaces)."""
    K = edge_attr2.shape[1]
    S = np.zeros((K, n, n), dtype=np.float64)
    S[:, edge_index2[0], edge_index2[1]] = np.asarray(edge_attr2, dtype=np.float64).T
    return S


# --------------------------------------------------------------------------------------------------
# input helpers (graph6 parsing replaces networkx.read_graph6 used at libs/utils.py:473; synthetic shapes
# follow SURVEY.md section 8d)
# --------------------------------------------------------------------------------------------------


def parse_graph6(path_or_bytes):
    """Parse a .g6 file (n <= 62) into a list of (n, edge_index[2, 2m] int64, both directions, sorted
    row-major) -- same edge set as ``to_undirected(list(nx.read_graph6(...).edges()))`` (libs/utils.py:473-478)."""
    data = open(path_or_bytes, "rb").read() if isinstance(path_or_bytes, str) else path_or_bytes
    out = []
    for line in data.split(b"\n"):
        line = line.strip()
        if not line:
            continue
        v = np.frombuffer(line, dtype=np.uint8).astype(np.int64) - 63
        n = int(v[0])
        assert n <= 62
        bits = ((v[1:, None] >> np.arange(5, -1, -1)[None, :]) & 1).reshape(-1)
        A = np.zeros((n, n), dtype=np.int64)
        k = 0
        for j in range(1, n):
            for i in range(j):
                A[i, j] = A[j, i] = bits[k]
                k += 1
        r, c = np.where(A > 0)
        out.append((n, np.vstack((r, c)).astype(np.int64)))
    return out
