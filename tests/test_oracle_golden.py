"""CPU: the oracle restatement (oracle/gnnml3_oracle.py) against fixtures produced by the UNMODIFIED
reference (oracle/make_golden.py).  This is what pins the oracle (task section 3)."""
import os

import numpy as np
import pytest
import torch

from oracle import gnnml3_oracle as O
from conftest import GOLDEN, load_npz

SD_Z, SD_META = load_npz("spectral_design.npz")
CV_Z, CV_META = load_npz("spect_conv.npz")


@pytest.mark.parametrize("case", SD_META, ids=[c["name"] for c in SD_META])
def test_spectral_design_matches_reference(case):
    n = case["name"]
    with np.errstate(all="ignore"):
        o = O.spectral_design(SD_Z[n + "/ei"], SD_Z[n + "/x"], **case["kw"])
    # edge indexing: bit-exact
    assert np.array_equal(o["edge_index2"], SD_Z[n + "/ei2"])
    assert o["edge_index2"].dtype == np.int64
    assert np.array_equal(o["x"], SD_Z[n + "/ox"])
    # same LAPACK, same numpy: values agree to float32 round-off (compared as matrices)
    nn = SD_Z[n + "/x"].shape[0]
    a = O.supports_dense(o["edge_index2"], o["edge_attr2"], nn)
    b = O.supports_dense(SD_Z[n + "/ei2"], SD_Z[n + "/ea2"], nn)
    np.testing.assert_allclose(a, b, rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(o["lmax"], SD_Z[n + "/lmax"], rtol=1e-6)


def _params(z, n):
    pre = n + "/p/"
    return {k[len(pre):]: torch.tensor(z[k]) for k in z.files if k.startswith(pre)}


@pytest.mark.parametrize("case", [c for c in CV_META if c["kind"] == "conv"], ids=lambda c: c["name"])
def test_spectconv_matches_reference(case):
    n = case["name"]
    p = {k: v.requires_grad_(True) for k, v in _params(CV_Z, n).items()}
    x = torch.tensor(CV_Z[n + "/x"], requires_grad=True)
    ea = torch.tensor(CV_Z[n + "/ea"], requires_grad=True)
    ei = torch.tensor(CV_Z[n + "/ei"])
    kw = case["kw"]
    out = O.spectconv_forward(x, ei, ea, p["weight"], p.get("bias"), selfconn=kw.get("selfconn", True),
                              depthwise=kw.get("depthwise", False), DSweight=p.get("DSweight"))
    np.testing.assert_array_equal(out.detach().numpy(), CV_Z[n + "/out"])      # same ops, same order: bit-equal
    out.backward(torch.tensor(CV_Z[n + "/gout"]))
    np.testing.assert_allclose(x.grad.numpy(), CV_Z[n + "/gx"], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(ea.grad.numpy(), CV_Z[n + "/gea"], rtol=1e-6, atol=1e-6)
    for k, v in p.items():
        np.testing.assert_allclose(v.grad.numpy(), CV_Z[n + "/g/" + k], rtol=1e-6, atol=1e-5)


@pytest.mark.parametrize("case", [c for c in CV_META if c["kind"] == "conv" and not c["kw"].get("depthwise")
                                  and not c["kw"].get("selfconn", True)], ids=lambda c: c["name"])
def test_dense_support_identity(case):
    """sum_k S_k X W_k (libs/layers_tf.py:231-236) equals the edge-list form (libs/spect_conv.py:76-80)."""
    n = case["name"]
    p = _params(CV_Z, n)
    out = O.spectconv_forward_dense(torch.tensor(CV_Z[n + "/x"]), torch.tensor(CV_Z[n + "/ei"]),
                                    torch.tensor(CV_Z[n + "/ea"]), p["weight"], p.get("bias"))
    ref = CV_Z[n + "/out"]
    np.testing.assert_allclose(out.numpy(), ref, rtol=1e-4, atol=1e-4 * np.abs(ref).max())


@pytest.mark.parametrize("case", [c for c in CV_META if c["kind"] == "layer"], ids=lambda c: c["name"])
def test_ml3layer_matches_reference(case):
    n = case["name"]
    learnedge, kin, kout, ninp, nout1, nout2 = case["args"]
    p = {k: v.requires_grad_(True) for k, v in _params(CV_Z, n).items()}
    assert list(p.keys()) == case["keys"]
    x = torch.tensor(CV_Z[n + "/x"], requires_grad=True)
    ea = torch.tensor(CV_Z[n + "/ea"], requires_grad=True)
    out = O.ml3layer_forward(x, torch.tensor(CV_Z[n + "/ei"]), ea, p, learnedge, nout2)
    np.testing.assert_allclose(out.detach().numpy(), CV_Z[n + "/out"], rtol=1e-6, atol=1e-6)
    out.backward(torch.tensor(CV_Z[n + "/gout"]))
    np.testing.assert_allclose(x.grad.numpy(), CV_Z[n + "/gx"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(ea.grad.numpy(), CV_Z[n + "/gea"], rtol=1e-5, atol=1e-5)
    for k, v in p.items():
        np.testing.assert_allclose(v.grad.numpy(), CV_Z[n + "/g/" + k], rtol=1e-5, atol=1e-4)


def test_oracle_module_keys_and_init():
    m = O.OracleML3Layer(True, 6, 6, 2, 32, 16)
    assert [k for k, _ in m.named_parameters()] == ["fc1_1.weight", "fc1_2.weight", "fc1_3.weight", "fc1_4.weight",
                                                    "conv1.weight", "conv1.bias", "fc11.weight", "fc11.bias",
                                                    "fc12.weight", "fc12.bias"]
    s = (6.0 / (2 + 32)) ** 0.5
    assert m.conv1.weight.abs().max() <= s and m.conv1.bias.abs().max() == 0


def test_graph8c_model_fixture():
    """graph8c.py:249-279 forward on the first 300 graphs with the reference's seed-0 weights."""
    z, _ = load_npz("graph8c_model.npz")
    g8 = O.parse_graph6(os.path.join(GOLDEN, "graph8c.g6"))
    assert len(g8) == 11117 and all(n == 8 for n, _ in g8)
    graphs = [O.spectral_design(ei, np.ones((n, 1), np.float32), recfield=1, dv=2, nfreq=5, adddegree=True)
              for n, ei in g8[:300]]
    model = O.OracleGNNML3("graph8c", ne=6, ninp=2)
    model.load_state_dict({k[2:]: torch.tensor(z[k]) for k in z.files if k.startswith("p/")})
    assert sum(p.numel() for p in model.parameters()) == 23714
    with torch.no_grad():
        emb = torch.cat([model(O.collate(graphs[i:i + 100])) for i in range(0, 300, 100)])
    np.testing.assert_allclose(emb.numpy(), z["emb"], rtol=1e-5, atol=1e-6)


def test_filtering_model_fixture():
    """filtering.py:252-281 built from the UNMODIFIED reference's `ML3Layer(learnedge=False)` on a 12x12 grid (supports: the
    grid12_filtering SpectralDesign fixture): the oracle model reproduces the per-node output and every parameter gradient."""
    z, _ = load_npz("filtering_model.npz")
    ei2, ea2 = SD_Z["grid12_filtering/ei2"], SD_Z["grid12_filtering/ea2"]
    model = O.OracleGNNML3Variant("filtering", ea2.shape[1], 1)
    assert [k for k, _ in model.named_parameters()] == [k[2:] for k in z.files if k.startswith("p/")]       # names and creation order
    model.load_state_dict({k[2:]: torch.tensor(z[k]) for k in z.files if k.startswith("p/")})
    out = model(dict(x=torch.tensor(z["x"]), edge_index2=torch.tensor(ei2), edge_attr2=torch.tensor(ea2)))
    np.testing.assert_allclose(out.detach().numpy(), z["out"], rtol=1e-5, atol=1e-6)
    out.backward(torch.tensor(z["gout"]))
    for k, p in model.named_parameters():
        g = z["g/" + k]
        np.testing.assert_allclose(p.grad.numpy(), g, rtol=1e-4, atol=1e-5 * np.abs(g).max())


def test_graph8c_isomorphism_kat():
    """The reference's own known-answer run (graph8c.py:281-302): embed all 11,117 graphs with freshly initialised models
    (``torch.manual_seed(iter)``), mark a pair distinguished when its L1 embedding distance exceeds 1e-3 under any seed so
    far, print the number of never-distinguished pairs.  The unmodified reference (imported behind the PyG stand-in at
    survey time) gives 1 after seed 0 and 0 after seeds 0-1 -- the paper's Table-1 result for GNNML3;
    the oracle (same init order, same arithmetic) must reproduce both counts."""
    g8 = O.parse_graph6(os.path.join(GOLDEN, "graph8c.g6"))
    graphs = [O.spectral_design(ei, np.ones((n, 1), np.float32), recfield=1, dv=2, nfreq=5, adddegree=True) for n, ei in g8]
    batches = [O.collate(graphs[i:i + 100]) for i in range(0, len(graphs), 100)]       # graph8c.py:18, batch_size=100
    n = len(graphs)
    seen = torch.zeros(n, n, dtype=torch.bool)
    similar = []
    for seed in range(2):
        torch.manual_seed(seed)
        model = O.OracleGNNML3("graph8c", ne=6, ninp=2).eval()
        with torch.no_grad():
            emb = torch.cat([model(b) for b in batches])
        for r in range(0, n, 2048):
            seen[r:r + 2048] |= torch.cdist(emb[r:r + 2048], emb, p=1) > 0.001
        similar.append((int((~seen).sum()) - n) // 2)
    assert similar == [1, 0]


def test_sr25_known_answer():
    """sr25.py:248-301: the 15 strongly regular graphs SR(25,12,5,6) are 3-WL equivalent, so GNNML3 must NOT distinguish any
    of the 105 pairs (paper Table 1) -- a check that the spectral supports are permutation-equivariant to rounding: the
    embeddings of all 15 graphs agree to ~1e-6 under every seed.  (tests/golden/sr251256.g6 is the public data file.)"""
    from gnn_matlang_b200.datasets import read_graph6
    g = read_graph6(os.path.join(GOLDEN, "sr251256.g6"))
    assert len(g) == 15 and all(d["x"].shape == (25, 1) and d["edge_index"].shape == (2, 300) for d in g)
    batch = O.collate([O.spectral_design(d["edge_index"], d["x"], recfield=1, dv=2, nfreq=5, adddegree=True) for d in g])
    seen = torch.zeros(15, 15, dtype=torch.bool)
    for seed in range(2):
        torch.manual_seed(seed)
        model = O.OracleGNNML3("graph8c", ne=6, ninp=2).eval()              # same architecture as graph8c.py
        with torch.no_grad():
            emb = model(batch)
        dist = torch.cdist(emb, emb, p=1)
        assert float(dist.max()) < 1e-5
        seen |= dist > 0.001
    assert (int((~seen).sum()) - 15) // 2 == 105


def test_exp_pairs_known_answer():
    """exp_iso.py:283-305 on the first 100 EXP pairs (tests/golden/exp_first200.npz): every pair of 1-WL-equivalent,
    non-isomorphic graphs is told apart by a freshly initialised GNNML3 (paper: 0 similar pairs; the full 600-pair file gives
    0 as well with the reference checkout present)."""
    z = np.load(os.path.join(GOLDEN, "exp_first200.npz"))
    xo, eo = np.cumsum(np.r_[0, z["n"]]), np.cumsum(np.r_[0, z["ne"]])
    with np.errstate(all="ignore"):
        graphs = [O.spectral_design(z["edge_index"][:, eo[i]:eo[i + 1]].astype(np.int64),
                                    z["x"][xo[i]:xo[i + 1]].reshape(-1, 1).astype(np.float32),
                                    recfield=1, dv=2, nfreq=5, adddegree=True) for i in range(200)]
    batches = [O.collate(graphs[i:i + 100]) for i in range(0, 200, 100)]
    torch.manual_seed(0)
    model = O.OracleGNNML3("graph8c", ne=6, ninp=2).eval()                  # exp_iso.py:248-280 = the graph8c architecture
    with torch.no_grad():
        emb = torch.cat([model(b) for b in batches]).numpy()
    assert int((np.abs(emb[0::2] - emb[1::2]).sum(1) <= 0.001).sum()) == 0


def test_collate_and_pool_semantics():
    g = [dict(x=np.ones((3, 2), np.float32), edge_index2=np.array([[0, 1, 2], [1, 2, 0]]), edge_attr2=np.ones((3, 4), np.float32), y=1.0),
         dict(x=2 * np.ones((2, 2), np.float32), edge_index2=np.array([[0, 1], [1, 0]]), edge_attr2=np.zeros((2, 4), np.float32), y=0.0)]
    b = O.collate(g)
    assert b["edge_index2"].tolist() == [[0, 1, 2, 3, 4], [1, 2, 0, 4, 3]]
    assert b["batch"].tolist() == [0, 0, 0, 1, 1] and b["y"].shape == (2, 1)
    assert O.global_add_pool(b["x"], b["batch"], 2).tolist() == [[3, 3], [4, 4]]
    assert O.global_mean_pool(b["x"], b["batch"], 2).tolist() == [[1, 1], [2, 2]]


def test_spectral_design_invariants():
    """identity column == [src == dst]; adjacency column == [src != dst] for recfield 1;
    edge_index2 strictly increasing in src * n + dst."""
    for c in SD_META:
        n = c["name"]
        ei2, ea2 = SD_Z[n + "/ei2"], SD_Z[n + "/ea2"]
        nn = SD_Z[n + "/x"].shape[0]
        nf = c["kw"]["nfreq"]
        assert np.all(np.diff(ei2[0] * nn + ei2[1]) > 0)
        assert np.array_equal(ea2[:, nf], (ei2[0] == ei2[1]).astype(np.float32))
        if c["kw"].get("addadj") and c["kw"]["recfield"] == 1:
            assert np.array_equal(ea2[:, nf + 1], (ei2[0] != ei2[1]).astype(np.float32))


CC_Z, CC_META = load_npz("spect_concat.npz")


@pytest.mark.parametrize("case", CC_META, ids=lambda c: c["name"])
def test_spectconcatconv_matches_reference(case):
    """oracle restatement of libs/spect_conv.py:105-165 against the fixture generated by the unmodified reference"""
    n = case["name"]
    p = {k: v.requires_grad_(True) for k, v in _params(CC_Z, n).items()}
    x = torch.tensor(CC_Z[n + "/x"], requires_grad=True)
    ea = torch.tensor(CC_Z[n + "/ea"], requires_grad=True)
    ei = torch.tensor(CC_Z[n + "/ei"])
    out = O.spectconcatconv_forward(x, ei, ea, p["weight"], p.get("bias"), selfconn=case["kw"].get("selfconn", True))
    np.testing.assert_array_equal(out.detach().numpy(), CC_Z[n + "/out"])      # same ops, same order: bit-equal
    out.backward(torch.tensor(CC_Z[n + "/gout"]))
    np.testing.assert_allclose(x.grad.numpy(), CC_Z[n + "/gx"], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(ea.grad.numpy(), CC_Z[n + "/gea"], rtol=1e-6, atol=1e-6)
    for k, v in p.items():
        np.testing.assert_allclose(v.grad.numpy(), CC_Z[n + "/g/" + k], rtol=1e-6, atol=1e-6)
