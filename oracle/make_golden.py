"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz by running the UNMODIFIED reference.

Run in the build container (needs /root/reference):  ``python oracle/make_golden.py``
The reference's libs/spect_conv.py and libs/utils.py are imported verbatim behind oracle/pyg_stub.py;
the model wrapper for the graph8c fixture is instantiated from the reference's ML3Layer with the layer
sizes of graph8c.py:249-279.  Inputs and outputs are stored together so that neither the GPU box nor the
CPU test-suite ever needs /root/reference.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import pyg_stub  # noqa: E402
from oracle.gnnml3_oracle import parse_graph6  # noqa: E402

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def rand_tree_with_rings(rng, n, maxdeg=4, rings=2):
    """ZINC-shaped molecule-like graph: random tree with bounded degree plus a few ring closures."""
    deg = np.zeros(n, dtype=np.int64)
    edges = set()
    for v in range(1, n):
        while True:
            u = int(rng.integers(max(0, v - 6), v))
            if deg[u] < maxdeg - 1:
                break
            u = int(rng.integers(0, v))
            if deg[u] < maxdeg:
                break
        edges.add((u, v))
        deg[u] += 1
        deg[v] += 1
    for _ in range(rings):
        for _try in range(20):
            u, v = (int(t) for t in rng.integers(0, n, 2))
            if u != v and (min(u, v), max(u, v)) not in edges and deg[u] < maxdeg and deg[v] < maxdeg and 2 <= abs(u - v) <= 6:
                edges.add((min(u, v), max(u, v)))
                deg[u] += 1
                deg[v] += 1
                break
    e = np.array(sorted(edges), dtype=np.int64).T
    ei = np.concatenate([e, e[::-1]], 1)
    order = np.lexsort((ei[1], ei[0]))
    return ei[:, order]


def rand_regular(rng, n, d):
    """Random d-regular simple graph by repeated pairing (counting-shaped)."""
    while True:
        stubs = np.repeat(np.arange(n), d)
        rng.shuffle(stubs)
        a, b = stubs[0::2], stubs[1::2]
        if np.any(a == b):
            continue
        key = np.minimum(a, b) * n + np.maximum(a, b)
        if len(np.unique(key)) != len(key):
            continue
        ei = np.concatenate([np.vstack((a, b)), np.vstack((b, a))], 1)
        order = np.lexsort((ei[1], ei[0]))
        return ei[:, order].astype(np.int64)


def main():
    os.makedirs(OUT, exist_ok=True)
    ref_conv, ref_utils = pyg_stub.import_reference(REF)
    Data = pyg_stub.Data
    rng = np.random.default_rng(0)

    # ------------------------------------------------------------------ inputs: real files -> compact npz
    g8 = parse_graph6(os.path.join(REF, "dataset/graph8c/raw/graph8c.g6"))
    assert len(g8) == 11117
    with open(os.path.join(REF, "dataset/graph8c/raw/graph8c.g6"), "rb") as f:
        raw = f.read()
    with open(os.path.join(OUT, "graph8c.g6"), "wb") as f:       # public data file (McKay), 78 KB, not source
        f.write(raw)

    import pickle
    exp_list = pickle.load(open(os.path.join(REF, "dataset/EXP/raw/GRAPHSAT.pkl"), "rb"))
    nexp = 200                                                   # first 100 pairs
    np.savez_compressed(
        os.path.join(OUT, "exp_first200.npz"),
        n=np.array([d.x.shape[0] for d in exp_list[:nexp]], dtype=np.int64),
        x=np.concatenate([d.x.numpy().reshape(-1) for d in exp_list[:nexp]]).astype(np.int8),
        y=np.array([int(d.y) for d in exp_list[:nexp]], dtype=np.int8),
        ne=np.array([d.edge_index.shape[1] for d in exp_list[:nexp]], dtype=np.int64),
        edge_index=np.concatenate([d.edge_index.numpy() for d in exp_list[:nexp]], 1).astype(np.int16))

    # ------------------------------------------------------------------ SpectralDesign fixtures
    cases = []

    def add_case(name, ei, x, **kw):
        d = Data(edge_index=torch.as_tensor(ei, dtype=torch.int64), x=torch.as_tensor(x))
        with np.errstate(all="ignore"):
            o = ref_utils.SpectralDesign(nmax=0, **kw)(d)
        cases.append(dict(name=name, kw=kw, ei=np.asarray(ei, dtype=np.int64), x=np.asarray(x, dtype=np.float32),
                          ox=o.x.numpy(), lmax=np.float32(o.lmax), ei2=o.edge_index2.numpy(), ea2=o.edge_attr2.numpy()))

    kw_g8 = dict(recfield=1, dv=2, nfreq=5, adddegree=True)                     # graph8c.py:16, exp_classify.py:16
    kw_zinc = dict(recfield=2, dv=2, nfreq=7)                                   # Zinc12k.py:12
    kw_cnt = dict(recfield=1, dv=1, nfreq=10, adddegree=True, laplacien=False, addadj=True)   # counting.py:16
    for i in list(range(0, 8)) + [100, 5000, 11116]:
        n, ei = g8[i]
        add_case("graph8c_%d" % i, ei, np.ones((n, 1), np.float32), **kw_g8)
    for i in (0, 1, 2, 3):
        d = exp_list[i]
        add_case("exp_%d" % i, d.edge_index.numpy(), d.x.numpy().astype(np.float32), **kw_g8)
    for i in range(4):
        n = int(np.clip(round(rng.normal(23.2, 4.3)), 9, 37))
        ei = rand_tree_with_rings(rng, n)
        x = np.zeros((n, 25), np.float32)
        x[np.arange(n), rng.integers(0, 21, n)] = 1
        deg = np.bincount(ei[0], minlength=n)
        x[np.arange(n), 25 - np.clip(deg, 1, 4)] = 1
        add_case("zinc_%d" % i, ei, x, **kw_zinc)
    for i, (n, dg) in enumerate([(10, 6), (15, 6), (20, 5), (30, 5)]):
        add_case("count_%d" % i, rand_regular(rng, n, dg), np.ones((n, 1), np.float32), **kw_cnt)
    # sweep-shaped: G(n, 4/(n-1)) with 1-hop and 2-hop masks, K = 10 (nfreq = 9)
    for i, n in enumerate([30, 64, 100]):
        up = np.triu(rng.random((n, n)) < 4.0 / (n - 1), 1)
        r, c = np.where(up | up.T)
        add_case("sweep1_%d" % i, np.vstack((r, c)), rng.standard_normal((n, 3)).astype(np.float32), recfield=1, dv=5, nfreq=9)
        add_case("sweep2_%d" % i, np.vstack((r, c)), rng.standard_normal((n, 3)).astype(np.float32), recfield=2, dv=5, nfreq=9)
    # edge cases: isolated nodes (libs/utils.py:578-580), recfield 0 / 3, vmax given, directed input, single node
    n, ei = g8[3]
    ei_iso = ei[:, (ei[0] != 0) & (ei[1] != 0)]
    add_case("isolated", ei_iso, np.ones((n, 1), np.float32), **kw_g8)
    add_case("recfield0", ei, np.ones((n, 1), np.float32), recfield=0, dv=2, nfreq=4)
    add_case("recfield3", g8[200][1], np.ones((8, 1), np.float32), recfield=3, dv=1, nfreq=3, addadj=True)
    add_case("vmax", ei, np.ones((n, 2), np.float32), recfield=1, dv=5, nfreq=5, vmax=2.0)
    add_case("adj_noclamp", g8[77][1], np.ones((8, 1), np.float32), recfield=2, dv=0.5, nfreq=6, laplacien=False)
    ei_dir = exp_list[5].edge_index.numpy()
    ei_dir = ei_dir[:, ei_dir[0] < ei_dir[1]]
    add_case("directed", ei_dir, exp_list[5].x.numpy().astype(np.float32), **kw_g8)
    add_case("no_edges", np.zeros((2, 0), np.int64), np.ones((3, 1), np.float32), **kw_g8)
    add_case("single_node", np.zeros((2, 0), np.int64), np.ones((1, 1), np.float32), **kw_g8)
    # graphs beyond the shared-memory eigensolver of the CUDA path (2-D grids as in filtering.py:17, small enough for a fixture):
    # the product designs them on its dense device path.  (No draws from `rng`: the fixtures above and below stay byte-identical.)

    def grid(side):
        idx = np.arange(side * side).reshape(side, side)
        e = np.concatenate([np.stack([idx[:, :-1].ravel(), idx[:, 1:].ravel()]), np.stack([idx[:-1, :].ravel(), idx[1:, :].ravel()])], 1)
        return side * side, np.concatenate([e, e[::-1]], 1)
    n, ei = grid(12)
    add_case("grid12_filtering", ei, np.ones((n, 1), np.float32), recfield=3, dv=10, nfreq=10, adddegree=False)   # filtering.py:17 (recfield 5 there)
    n, ei = grid(14)
    add_case("grid14_degree_adj", ei, np.ones((n, 1), np.float32), recfield=2, dv=5, nfreq=5, adddegree=True, addadj=True)
    n, ei = grid(13)
    add_case("grid13_adjacency_spectrum", ei, np.ones((n, 2), np.float32), recfield=1, dv=2, nfreq=4, laplacien=False, vmax=3.0)

    blob = {}
    import json
    blob["meta"] = np.frombuffer(json.dumps([dict(name=c["name"], kw=c["kw"]) for c in cases]).encode(), dtype=np.uint8)
    for c in cases:
        for k in ("ei", "x", "ox", "lmax", "ei2", "ea2"):
            blob[c["name"] + "/" + k] = c[k]
    np.savez_compressed(os.path.join(OUT, "spectral_design.npz"), **blob)

    # ------------------------------------------------------------------ SpectConv / ML3Layer fixtures
    torch.manual_seed(0)
    blob = {}
    meta = []

    def rand_graph(N, E, K, Fi):
        ei = torch.randint(0, N, (2, E))
        return torch.randn(N, Fi), ei, torch.randn(E, K)

    conv_cases = [
        dict(name="conv_default", N=40, E=300, K=3, Fi=5, Fo=7, kw=dict(selfconn=False)),
        dict(name="conv_selfconn", N=33, E=200, K=2, Fi=4, Fo=6, kw=dict(selfconn=True)),
        dict(name="conv_nobias", N=17, E=90, K=4, Fi=8, Fo=3, kw=dict(selfconn=False, bias=False)),
        dict(name="conv_depthwise", N=25, E=160, K=3, Fi=6, Fo=5, kw=dict(selfconn=True, depthwise=True)),
        dict(name="conv_depthwise_noself", N=25, E=160, K=3, Fi=6, Fo=5, kw=dict(selfconn=False, depthwise=True)),
        dict(name="conv_k1", N=50, E=260, K=1, Fi=16, Fo=16, kw=dict(selfconn=False)),
        dict(name="conv_wide", N=64, E=700, K=10, Fi=64, Fo=64, kw=dict(selfconn=False)),
        dict(name="conv_empty_rows", N=30, E=20, K=2, Fi=3, Fo=4, kw=dict(selfconn=False)),
    ]
    for c in conv_cases:
        x, ei, ea = rand_graph(c["N"], c["E"], c["K"], c["Fi"])
        x.requires_grad_(True)
        ea.requires_grad_(True)
        m = ref_conv.SpectConv(c["Fi"], c["Fo"], c["K"], **c["kw"])
        if c["kw"].get("depthwise"):
            with torch.no_grad():
                m.DSweight.normal_(0, 0.5)
        if m.bias is not None:
            with torch.no_grad():
                m.bias.normal_(0, 0.3)
        out = m(x, ei, ea)
        gout = torch.randn_like(out)
        out.backward(gout)
        n = c["name"]
        blob[n + "/x"], blob[n + "/ei"], blob[n + "/ea"] = x.detach().numpy(), ei.numpy(), ea.detach().numpy()
        blob[n + "/out"], blob[n + "/gout"] = out.detach().numpy(), gout.numpy()
        blob[n + "/gx"], blob[n + "/gea"] = x.grad.numpy(), ea.grad.numpy()
        for pn, p in m.named_parameters():
            blob[n + "/p/" + pn] = p.detach().numpy()
            blob[n + "/g/" + pn] = p.grad.numpy()
        meta.append(dict(name=n, kind="conv", Fi=c["Fi"], Fo=c["Fo"], K=c["K"], kw=c["kw"], repr=repr(m)))

    layer_cases = [
        dict(name="layer_learn", N=36, E=280, args=(True, 6, 6, 5, 8, 4)),
        dict(name="layer_learn_kout", N=36, E=280, args=(True, 4, 3, 7, 6, 2)),
        dict(name="layer_nolearn", N=30, E=200, args=(False, 5, 5, 6, 7, 3)),
        dict(name="layer_nogate", N=30, E=200, args=(True, 8, 8, 10, 12, 0)),
        dict(name="layer_zinc", N=70, E=450, args=(True, 8, 8, 32, 30, 2)),
        dict(name="layer_count", N=60, E=420, args=(True, 12, 12, 2, 16, 16)),
    ]
    for c in layer_cases:
        learnedge, kin, kout, ninp, nout1, nout2 = c["args"]
        x, ei, ea = rand_graph(c["N"], c["E"], kin, ninp)
        x.requires_grad_(True)
        ea.requires_grad_(True)
        m = ref_conv.ML3Layer(*c["args"])
        with torch.no_grad():
            m.conv1.bias.normal_(0, 0.3)
        out = m(x, ei, ea)
        gout = torch.randn_like(out)
        out.backward(gout)
        n = c["name"]
        blob[n + "/x"], blob[n + "/ei"], blob[n + "/ea"] = x.detach().numpy(), ei.numpy(), ea.detach().numpy()
        blob[n + "/out"], blob[n + "/gout"] = out.detach().numpy(), gout.numpy()
        blob[n + "/gx"], blob[n + "/gea"] = x.grad.numpy(), ea.grad.numpy()
        for pn, p in m.named_parameters():
            blob[n + "/p/" + pn] = p.detach().numpy()
            blob[n + "/g/" + pn] = p.grad.numpy()
        meta.append(dict(name=n, kind="layer", args=list(c["args"]), keys=[k for k, _ in m.named_parameters()]))
    blob["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(OUT, "spect_conv.npz"), **blob)

    # ------------------------------------------------------------------ graph8c GNNML3 model fixture
    # graph8c.py:249-279 with torch.manual_seed(0) (graph8c.py:284); first 300 graphs, batch 100 (:18)
    sd = ref_utils.SpectralDesign(nmax=0, **kw_g8)
    graphs = []
    for n, ei in g8[:300]:
        d = sd(Data(edge_index=torch.as_tensor(ei), x=torch.ones(n, 1)))
        graphs.append(d)
    ne, ninp = graphs[0].edge_attr2.shape[1], graphs[0].x.shape[1]

    class RefGNNML3(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.conv1 = ref_conv.ML3Layer(learnedge=True, nedgeinput=ne, nedgeoutput=ne, ninp=ninp, nout1=32, nout2=16)
            self.conv2 = ref_conv.ML3Layer(learnedge=True, nedgeinput=ne, nedgeoutput=ne, ninp=48, nout1=32, nout2=16)
            self.conv3 = ref_conv.ML3Layer(learnedge=True, nedgeinput=ne, nedgeoutput=ne, ninp=48, nout1=32, nout2=16)
            self.fc1 = torch.nn.Linear(48, 10)

    torch.manual_seed(0)
    model = RefGNNML3().eval()
    emb = []
    with torch.no_grad():
        for b0 in range(0, 300, 100):
            gs = graphs[b0:b0 + 100]
            off = np.cumsum([0] + [g.x.shape[0] for g in gs])
            x = torch.cat([g.x for g in gs])
            ei = torch.cat([g.edge_index2 + int(o) for g, o in zip(gs, off)], 1)
            ea = torch.cat([g.edge_attr2 for g in gs])
            bt = torch.cat([torch.full((g.x.shape[0],), i, dtype=torch.int64) for i, g in enumerate(gs)])
            x = model.conv3(model.conv2(model.conv1(x, ei, ea), ei, ea), ei, ea)
            pooled = torch.zeros(len(gs), x.shape[1]).index_add(0, bt, x)           # global_add_pool
            emb.append(torch.tanh(model.fc1(pooled)))
    blob = {"emb": torch.cat(emb).numpy()}
    for k, v in model.state_dict().items():
        blob["p/" + k] = v.numpy()
    np.savez_compressed(os.path.join(OUT, "graph8c_model.npz"), **blob)
    # ------------------------------------------------------------------ SpectConCatConv fixtures (libs/spect_conv.py:105-165)
    # (own file and own seed: the fixtures above keep their RNG stream and regenerate bit-identically)
    torch.manual_seed(1)
    blob, meta = {}, []
    for c in [dict(name="concat_selfconn", N=40, E=300, K=3, Fi=5, Fo=7, kw=dict(selfconn=True)),
              dict(name="concat_noself", N=33, E=250, K=4, Fi=6, Fo=3, kw=dict(selfconn=False)),
              dict(name="concat_nobias", N=20, E=120, K=2, Fi=8, Fo=8, kw=dict(selfconn=True, bias=False)),
              dict(name="concat_wide", N=64, E=600, K=6, Fi=40, Fo=12, kw=dict(selfconn=False))]:
        x, ei, ea = rand_graph(c["N"], c["E"], c["K"], c["Fi"])
        x.requires_grad_(True)
        ea.requires_grad_(True)
        m = ref_conv.SpectConCatConv(c["Fi"], c["Fo"], c["K"], **c["kw"])
        if m.bias is not None:
            with torch.no_grad():
                m.bias.normal_(0, 0.3)
        out = m(x, ei, ea)
        gout = torch.randn_like(out)
        out.backward(gout)
        n = c["name"]
        blob[n + "/x"], blob[n + "/ei"], blob[n + "/ea"] = x.detach().numpy(), ei.numpy(), ea.detach().numpy()
        blob[n + "/out"], blob[n + "/gout"] = out.detach().numpy(), gout.numpy()
        blob[n + "/gx"], blob[n + "/gea"] = x.grad.numpy(), ea.grad.numpy()
        for pn, p in m.named_parameters():
            blob[n + "/p/" + pn] = p.detach().numpy()
            blob[n + "/g/" + pn] = p.grad.numpy()
        meta.append(dict(name=n, kind="concat", Fi=c["Fi"], Fo=c["Fo"], K=c["K"], kw=c["kw"], repr=repr(m)))
    blob["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(OUT, "spect_concat.npz"), **blob)

    # ------------------------------------------------------------------ filtering.py GNNML3 (filtering.py:252-281) on a grid
    # 3 x ML3Layer(learnedge=False, 32||16) + fc2 per node, supports = the grid12_filtering case above (own file, own seed)
    c12 = [c for c in cases if c["name"] == "grid12_filtering"][0]
    ne_f = c12["ea2"].shape[1]

    class RefFiltering(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.conv1 = ref_conv.ML3Layer(learnedge=False, nedgeinput=ne_f, nedgeoutput=ne_f, ninp=1, nout1=32, nout2=16)
            self.conv2 = ref_conv.ML3Layer(learnedge=False, nedgeinput=ne_f, nedgeoutput=ne_f, ninp=48, nout1=32, nout2=16)
            self.conv3 = ref_conv.ML3Layer(learnedge=False, nedgeinput=ne_f, nedgeoutput=ne_f, ninp=48, nout1=32, nout2=16)
            self.fc2 = torch.nn.Linear(48, 1)

    torch.manual_seed(2)
    fm = RefFiltering()
    xf = torch.randn(c12["ox"].shape[0], 1)
    ei_f, ea_f = torch.as_tensor(c12["ei2"]), torch.as_tensor(c12["ea2"])
    out_f = fm.fc2(fm.conv3(fm.conv2(fm.conv1(xf, ei_f, ea_f), ei_f, ea_f), ei_f, ea_f))
    gout_f = torch.randn_like(out_f)
    out_f.backward(gout_f)
    blob = {"x": xf.numpy(), "out": out_f.detach().numpy(), "gout": gout_f.numpy()}
    for k, v in fm.named_parameters():
        blob["p/" + k] = v.detach().numpy()
        blob["g/" + k] = v.grad.numpy()
    np.savez_compressed(os.path.join(OUT, "filtering_model.npz"), **blob)

    print("params:", sum(p.numel() for p in model.parameters()))
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
