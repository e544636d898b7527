"""GPU tests written after the round-1 GPU budget was spent: collected last so that, should one of them
disagree with the hardware, the 224 tests that did run on a B200 are not cut off by `pytest -x`."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from conftest import GOLDEN  # noqa: E402
from oracle import gnnml3_oracle as O  # noqa: E402
from test_gpu_kernels import assert_close, dev  # noqa: E402


def test_gemm_tn_column_block_views():
    """Operands that are column-block views of wider buffers (row stride > width, as the layer entry points pass them): the
    columns between the width and the row stride hold other data and must not leak into the contraction -- on the tcgen05
    path they are outside the TMA tensor map (zero-filled), on the other paths they are masked."""
    from gnn_matlang_b200 import ops
    g = torch.Generator().manual_seed(9)
    M = 20000
    for Ka, lda, Nb, ldb in [(25, 28, 256, 260), (32, 36, 100, 128), (25, 28, 4, 8)]:
        abuf = torch.full((M, lda), 3.0)
        bbuf = torch.full((M, ldb), -5.0)
        abuf[:, :Ka] = torch.randn(M, Ka, generator=g)
        bbuf[:, :Nb] = torch.randn(M, Nb, generator=g)
        A, B = abuf.to(dev())[:, :Ka], bbuf.to(dev())[:, :Nb]
        assert A.stride(0) == lda and B.stride(0) == ldb
        out = ops.gemm_tn(A, B)
        assert_close(out, abuf[:, :Ka].double().t() @ bbuf[:, :Nb].double(), name="gemm_tn views %d/%d x %d/%d" % (Ka, lda, Nb, ldb))


def _embed_all(model, graphs, bs=100):
    from gnn_matlang_b200.batch import collate
    with torch.no_grad():
        return torch.cat([model(collate(graphs[i:i + bs]).to(dev())) for i in range(0, len(graphs), bs)])


def test_isomorphism_known_answers_on_the_gpu():
    """The reference's own known-answer runs through the CUDA path (fresh `torch.manual_seed(iter)` models, supports from the
    oracle's SpectralDesign, batch_size 100): graph8c.py:281-302 -> 1 undistinguished pair after seed 0, 0 after seeds 0-1;
    sr25.py -> all 105 pairs undistinguished; exp_iso.py on the first 100 pairs -> 0.  The oracle gives the same counts
    (tests/test_oracle_golden.py) with margins around the 1e-3 threshold (1.5e-4 / 1e-3 / 8e-3) far above the FP32 bar."""
    from gnn_matlang_b200.models import GNNML3
    kw = dict(recfield=1, dv=2, nfreq=5, adddegree=True)
    g8 = O.parse_graph6(os.path.join(GOLDEN, "graph8c.g6"))
    graphs8 = [O.spectral_design(ei, np.ones((n, 1), np.float32), **kw) for n, ei in g8]
    sr = O.parse_graph6(os.path.join(GOLDEN, "sr251256.g6"))
    graphs_sr = [O.spectral_design(ei, np.ones((n, 1), np.float32), **kw) for n, ei in sr]
    z = np.load(os.path.join(GOLDEN, "exp_first200.npz"))
    xo, eo = np.cumsum(np.r_[0, z["n"]]), np.cumsum(np.r_[0, z["ne"]])
    with np.errstate(all="ignore"):
        graphs_exp = [O.spectral_design(z["edge_index"][:, eo[i]:eo[i + 1]].astype(np.int64),
                                        z["x"][xo[i]:xo[i + 1]].reshape(-1, 1).astype(np.float32), **kw) for i in range(200)]
    n8 = len(graphs8)
    seen8 = torch.zeros(n8, n8, dtype=torch.bool, device=dev())
    similar8 = []
    for seed in range(2):
        torch.manual_seed(seed)
        model = GNNML3("graph8c", ne=6, ninp=2).to(dev()).eval()        # graph8c.py, sr25.py and exp_iso.py share the architecture
        emb = _embed_all(model, graphs8)
        for r in range(0, n8, 2048):
            seen8[r:r + 2048] |= torch.cdist(emb[r:r + 2048], emb, p=1) > 0.001
        similar8.append((int((~seen8).sum()) - n8) // 2)
        esr = _embed_all(model, graphs_sr)
        assert float(torch.cdist(esr, esr, p=1).max()) < 1e-3            # all 105 pairs of SR(25,12,5,6) graphs undistinguished
        if seed == 0:
            eexp = _embed_all(model, graphs_exp)
            assert int(((eexp[0::2] - eexp[1::2]).abs().sum(1) <= 0.001).sum()) == 0
    assert similar8 == [1, 0]


def test_two_streams_do_not_share_scratch():
    """scratch buffers are keyed by (device, stream, tag) -- two streams driving the library at once must
    not overwrite each other's partial sums.  Edge-MLP backward (per-group partials in the workspace) with different weights on
    two streams, enqueued back to back, against the same calls made one after the other."""
    from gnn_matlang_b200 import ops
    d = torch.device("cuda:0")
    g = torch.Generator().manual_seed(5)
    E, K = 300000, 8
    ea = torch.randn(E, K, generator=g).to(d)
    go = torch.randn(E, K, generator=g).to(d)
    wsets = [[(torch.randn(2 * K, K, generator=g) * 0.5).to(d) for _ in range(3)] + [(torch.randn(K, 4 * K, generator=g) * 0.3).to(d)]
             for _ in range(2)]
    ref = [ops.edge_mlp_bwd(ea, None, go, *w, need_dea=False)[1] for w in wsets]
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    for _ in range(3):
        out = []
        for s, w in zip(streams, wsets):
            with torch.cuda.stream(s):
                out.append(ops.edge_mlp_bwd(ea, None, go, *w, need_dea=False)[1])
        torch.cuda.synchronize()
        for o, r in zip(out, ref):
            for a, b in zip(o, r):
                assert torch.equal(a, b)
    keys = [k for k in ops._workspaces if k[0] == 0]
    assert len({k[1] for k in keys}) >= 3          # default stream + the two side streams


def test_ml3layer_without_learned_edges_ignores_extra_channels():
    """with learnedge=False the reference reads edge_attr[:, :K] only (libs/spect_conv.py:76-80)."""
    from gnn_matlang_b200.libs.spect_conv import ML3Layer
    d = torch.device("cuda:0")
    torch.manual_seed(0)
    N, E, K = 300, 2000, 4
    layer = ML3Layer(False, K, K, 10, 16, 4).to(d)
    x = torch.randn(N, 10, device=d)
    ei = torch.randint(0, N, (2, E), device=d)
    ea = torch.randn(E, K + 3, device=d)
    a = layer(x, ei, ea)
    b = layer(x, ei, ea[:, :K].contiguous())
    assert torch.equal(a, b)


def _dot(a, b):
    return float((a.double() * b.double()).sum())


@pytest.mark.parametrize("case", ["zinc_step_fused", "sweep_1M_nodes_f64"])
def test_full_size_adjoint_and_linearity_properties(case):
    """Parity at BASELINE.json's full sizes, where the CPU oracle would take minutes: size-independent properties of SpectConv.
    The layer (without bias) is bilinear in (x, edge_attr), so for random g
        <conv(x, ea), g> = <x, dL/dx> = <ea, dL/dea>            (L = <conv, g>: adjointness of the forward kernel and the two
                                                                 backward kernels -- fused dx over the transposed CSR, fused /
                                                                 two-kernel SDDMM)
        conv(a x1 + b x2, ea) = a conv(x1, ea) + b conv(x2, ea) (linearity)
        <W_k, dL/dW_k> summed over k = <conv, g>                (weight gradient, Euler's identity for a linear map)
    must hold to FP32 accuracy.  zinc_step_fused: the 8192-graph ZINC batch (K = 8, 32 -> 30: tensor-memory fused kernels);
    sweep_1M_nodes_f64: BASELINE configs[4] (1 M nodes, K = 10, F = 64: shared-memory plane kernel + two-kernel backward)."""
    from gnn_matlang_b200.libs.spect_conv import SpectConv
    from gnn_matlang_b200.synthetic import GraphPool
    d = torch.device("cuda:0")
    rng = np.random.default_rng(7)
    if case == "zinc_step_fused":
        pool = GraphPool("zinc", 512, seed=2)
        hb = pool.draw(rng, 8192)
        K, Fi, Fo = pool.K, 32, 30
    else:
        pool = GraphPool("sweep", 256, seed=1, K=10, nfeat=4, recfield=1)
        hb = pool.draw(rng, int(1000000 / float(pool.n.mean())))
        K, Fi, Fo = 10, 64, 64
    N = hb.x.shape[0]
    ei = hb.edge_index2.to(d)
    torch.manual_seed(1)
    layer = SpectConv(Fi, Fo, K, selfconn=False, bias=False).to(d)
    g = torch.Generator(device="cpu").manual_seed(3)
    ea = (torch.randn(ei.size(1), K, generator=g) * 0.3).to(d).requires_grad_(True)
    x1 = torch.randn(N, Fi, generator=g).to(d).requires_grad_(True)
    x2 = torch.randn(N, Fi, generator=g).to(d)
    gout = torch.randn(N, Fo, generator=g).to(d)
    y1 = layer(x1, ei, ea)
    y1.backward(gout)
    L = _dot(y1.detach(), gout)
    scale = float((y1.double().abs() * gout.double().abs()).sum())         # sum of |terms|: the rounding scale of the dot product
    assert abs(_dot(x1.detach(), x1.grad) - L) <= 2e-6 * scale, ("dx", _dot(x1.detach(), x1.grad), L, scale)
    assert abs(_dot(ea.detach(), ea.grad) - L) <= 2e-6 * scale, ("dea", _dot(ea.detach(), ea.grad), L, scale)
    assert abs(_dot(layer.weight.detach(), layer.weight.grad) - L) <= 2e-6 * scale, ("dW", _dot(layer.weight.detach(), layer.weight.grad), L)
    with torch.no_grad():
        y2 = layer(x2, ei, ea.detach())
        y3 = layer(0.75 * x1.detach() - 1.5 * x2, ei, ea.detach())
    ref = 0.75 * y1.detach().double() - 1.5 * y2.double()
    err = (y3.double() - ref).abs().max().item()
    assert err <= 2e-5 * ref.abs().max().item(), (err, ref.abs().max().item())


def test_aligned_rows_copies_foreign_column_blocks():
    """ADVICE (round 1): a caller's column-block view with F % 4 != 0 whose neighbouring columns hold NaN must not reach the
    128-bit gathers as it is (the kernels multiply the padding columns by zero weights, and 0 * NaN is NaN); tensors the
    library allocated itself (zeroed padding) still pass through without a copy."""
    from gnn_matlang_b200 import ops
    g = torch.Generator().manual_seed(5)
    N, F, K, Nc, deg = 3000, 30, 8, 16, 5
    wide = torch.full((N, 32), float("nan"))
    wide[:, :F] = torch.randn(N, F, generator=g)
    xv = wide.to(dev())[:, :F]
    assert xv.stride(0) == 32 and xv.data_ptr() % 16 == 0
    xa = ops.aligned_rows(xv)
    assert xa.data_ptr() != xv.data_ptr(), "a foreign view with F % 4 != 0 must be copied"
    src = torch.randint(0, N, (N * deg,), generator=g)
    dst = torch.randint(0, N, (N * deg,), generator=g)
    ei = torch.stack([src, dst])
    ea = torch.randn(N * deg, K, generator=g)
    W = torch.randn(K, F, Nc, generator=g) / np.sqrt(K * F)
    plan = ops.csr_build(ei.to(dev()), N)
    ea_s = ea.to(dev())[plan["perm"].long()].contiguous()
    out, _ = ops.fused_agg_proj(plan["rowptr"], plan["col"], None, ea_s, xa, W.view(K * F, Nc).to(dev()), epilogue=0)   # global gathers
    ref = torch.zeros(N, Nc, dtype=torch.float64)
    x64 = wide[:, :F].double()
    for k in range(K):
        ref += torch.zeros(N, F, dtype=torch.float64).index_add_(0, dst, ea[:, k:k + 1].double() * x64[src]) @ W[k].double()
    assert bool(torch.isfinite(out).all())
    assert_close(out, ref, name="conv on a copied column block")
    # library-allocated padded rows: no copy, and the same tensor goes straight into the next gather
    y, _ = ops.fused_agg_proj(plan["rowptr"], plan["col"], None, ea_s, xa, torch.randn(K * F, 30, generator=g).to(dev()), epilogue=0)
    assert y.size(1) == 30 and y.stride(0) == 32
    assert ops.aligned_rows(y).data_ptr() == y.data_ptr()
    assert float(y._base[:, 30:].abs().max()) == 0.0


@pytest.mark.parametrize("N,Fo,G", [(1000, 30, 2), (129, 32, 16), (5000, 16, 16), (777, 64, 0), (300, 13, 4), (2000, 32, 8), (640, 30, 0)])
def test_act_bwd_y_streaming_and_general_kernels(N, Fo, G):
    """d pre and the bias sums from the fused layer's outputs: the streaming kernel (aligned shapes: Fo4 in {16, 32, 64}, gate
    width 2 or a power-of-two multiple of 4 with Fo % 4 == 0) and the general kernel (everything else) against float64 autograd."""
    from gnn_matlang_b200 import ops
    g = torch.Generator().manual_seed(N + Fo + G)
    pre = torch.randn(N, Fo + 2 * G, generator=g)
    gy = torch.randn(N, Fo + G, generator=g)
    y = torch.cat([torch.relu(pre[:, :Fo]), torch.tanh(pre[:, Fo:Fo + G]) * torch.tanh(pre[:, Fo + G:])], 1)
    aux = torch.tanh(pre[:, Fo:])
    d = dev()
    W = Fo + G
    ld = (W + 3) // 4 * 4
    yb = torch.zeros(N, ld); yb[:, :W] = y
    gb = torch.zeros(N, ld); gb[:, :W] = gy
    gpre, csum = ops.ml3_act_bwd_y(yb.to(d)[:, :W], aux.to(d) if G else None, gb.to(d)[:, :W], Fo, G)
    pr = pre.double().requires_grad_(True)
    yr = torch.cat([torch.relu(pr[:, :Fo]), torch.tanh(pr[:, Fo:Fo + G]) * torch.tanh(pr[:, Fo + G:])], 1)
    yr.backward(gy.double())
    Fo4 = (Fo + 3) // 4 * 4
    assert_close(gpre[:, :Fo], pr.grad[:, :Fo], name="gc")
    if G:
        assert_close(gpre[:, Fo4:Fo4 + 2 * G], pr.grad[:, Fo:], name="gates")
    if Fo4 > Fo:
        assert float(gpre[:, Fo:Fo4].abs().max()) == 0.0
    assert_close(csum, pr.grad.sum(0), name="bias sums")
