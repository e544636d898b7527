"""GPU parity of the tcgen05 / TMEM GEMM (csrc/gemm_tc.cu) against float64 matmul -- FP32 bar (rtol 1e-5 with the
max|ref| companion), every tail case (rows, columns, contraction), both tile widths, bias / ReLU epilogues."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from test_gpu_kernels import assert_close, dev  # noqa: E402

TC_CASES = [(128, 32, 32), (1024, 32, 32), (1000, 64, 16), (5000, 256, 30), (4096, 640, 64), (3000, 244, 32), (2111, 36, 256),
            (777, 100, 7), (20000, 256, 30), (1500, 2560, 200), (129, 8, 1), (100000, 200, 30)]


@pytest.mark.parametrize("M,Kc,Nc", TC_CASES)
def test_gemm_nn_tc(M, Kc, Nc):
    from gnn_matlang_b200 import ops, _lib
    g = torch.Generator().manual_seed(M + Kc + Nc)
    A = torch.randn(M, Kc, generator=g)
    B = torch.randn(Kc, Nc, generator=g)
    bias = torch.randn(Nc, generator=g)
    ref = A.double() @ B.double() + bias.double()
    out = ops.gemm_nn_tc(A.to(dev()), B.to(dev()), bias.to(dev()))
    assert_close(out, ref, name="gemm_nn_tc")
    out_relu = ops.gemm_nn_tc(A.to(dev()), B.to(dev()), bias.to(dev()), epilogue=_lib.EPI_RELU)
    assert torch.equal(out_relu, out.clamp(min=0))
    assert torch.equal(out, ops.gemm_nn_tc(A.to(dev()), B.to(dev()), bias.to(dev())))       # deterministic


@pytest.mark.parametrize("chunk", [1, 2, 3, 8, 1000])
def test_gemm_nn_tc_chunking(chunk):
    """How many k-blocks accumulate inside tensor memory before the FP32 register fold must not change the result
    beyond rounding (and long in-TMEM accumulation stays within the bar at Kc = 640)."""
    from gnn_matlang_b200 import ops
    g = torch.Generator().manual_seed(chunk)
    A = torch.randn(3000, 640, generator=g)
    B = torch.randn(640, 64, generator=g)
    out = ops.gemm_nn_tc(A.to(dev()), B.to(dev()), None, chunk_kblocks=chunk)
    assert_close(out, A.double() @ B.double(), rtol=1e-5 if chunk <= 8 else 3e-5, name="chunk %d" % chunk)


def test_gemm_nn_tc_strided_operands():
    """A as a column-block view (lda > Kc, Kc not a multiple of 4) and C as a column-block view, as the layer uses."""
    from gnn_matlang_b200 import ops
    g = torch.Generator().manual_seed(0)
    buf = torch.randn(4000, 36, generator=g)
    B = torch.randn(30, 256, generator=g)
    A = buf.to(dev())[:, :30]
    out = torch.zeros(4000, 260, device=dev())
    ops.gemm_nn_tc(A, B.to(dev()), None, out=out[:, 4:])
    assert_close(out[:, 4:], buf[:, :30].double() @ B.double(), name="strided")
    assert float(out[:, :4].abs().max()) == 0.0


def test_gemm_nn_dispatches_to_tcgen05():
    from gnn_matlang_b200 import ops
    g = torch.Generator().manual_seed(1)
    A, B = torch.randn(2048, 64, generator=g).to(dev()), torch.randn(64, 48, generator=g).to(dev())
    ops.profile_start()
    out = ops.gemm_nn(A, B)
    recs = ops.profile_stop()
    assert "gemm_nn_tc" in [r[0] for r in recs]
    assert_close(out, A.double().cpu() @ B.double().cpu(), name="dispatch")
