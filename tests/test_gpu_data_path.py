"""The PyG-free data path on the device: compact wire format expanded on the GPU and the HBM-resident
dataset with on-device collation -- both bit-exact against host collation (integer work; the floats are copied, not computed)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from test_gpu_kernels import dev  # noqa: E402

FIELDS = ("x", "edge_index2", "edge_attr2", "batch", "y", "graph_ptr")


@pytest.mark.parametrize("kind,widths", [("zinc", (21, 4)), ("counting", None)])
def test_compact_batch_expands_bit_exact_on_the_gpu(kind, widths):
    from gnn_matlang_b200.batch import CompactBatch
    from gnn_matlang_b200.synthetic import GraphPool
    pool = GraphPool(kind, 48, seed=3)
    hb = pool.draw(np.random.default_rng(1), 200)
    cb = CompactBatch.from_batch(hb, widths)
    assert cb.nbytes() < hb.nbytes()
    b = cb.pin_memory().to(dev()).expand()
    for k in FIELDS:
        assert torch.equal(getattr(b, k).cpu(), getattr(hb, k)), k
        assert getattr(b, k).dtype == getattr(hb, k).dtype, k


def test_device_dataset_collation_bit_exact_and_trains():
    from gnn_matlang_b200.models import GNNML3
    from gnn_matlang_b200.synthetic import DeviceDataset, GraphPool
    from gnn_matlang_b200.train import Trainer
    pool = GraphPool("zinc", 64, seed=4)
    idx = np.random.default_rng(2).integers(0, 64, 300)
    hb = pool.collate(idx)
    dds = DeviceDataset(pool, dev())
    b = dds.collate(torch.from_numpy(idx).pin_memory())
    for k in FIELDS:
        assert torch.equal(getattr(b, k).cpu(), getattr(hb, k)), k
    # the device-collated batch drives a training step to the same loss as the host-collated one
    torch.manual_seed(0)
    m1 = GNNML3("zinc", pool.K, pool.F).to(dev())
    torch.manual_seed(0)
    m2 = GNNML3("zinc", pool.K, pool.F).to(dev())
    l1 = Trainer(m1, loss="l1").step(b)
    l2 = Trainer(m2, loss="l1").step(hb.to(dev()))
    assert float(l1) == float(l2)


def test_captured_step_equals_eager_step_and_padding_is_neutral():
    """train.GraphedTrainer: the whole optimisation step captured in one CUDA graph and replayed on padded batches gives the
    same losses and parameters as the eager Trainer on the unpadded batches (padding = one dummy graph of isolated zero nodes +
    zero-weight self-loops: exactly neutral), through both input routes (padded device batch; unpadded batch written into
    the captured buffers on the device)."""
    from gnn_matlang_b200.models import GNNML3
    from gnn_matlang_b200.synthetic import GraphPool
    from gnn_matlang_b200.train import GraphedTrainer, Trainer, pad_batch
    pool = GraphPool("zinc", 64, seed=6)
    rng = np.random.default_rng(3)
    host = [pool.draw(rng, 48) for _ in range(4)]
    Np, Ep = max(h.x.shape[0] for h in host) + 5, max(h.edge_index2.shape[1] for h in host) + 17
    torch.manual_seed(0)
    m1 = GNNML3("zinc", pool.K, pool.F).to(dev())
    torch.manual_seed(0)
    m2 = GNNML3("zinc", pool.K, pool.F).to(dev())
    eager = Trainer(m1, loss="l1", lr=1e-3)
    gt = GraphedTrainer(m2, pad_batch(host[0], Np, Ep), loss="l1", lr=1e-3, warmup=1)
    # the warm-up step trained m2 and initialised Adam: rewind both IN PLACE (the captured graph holds these tensors' addresses)
    with torch.no_grad():
        for a, b in zip(m2.parameters(), m1.parameters()):
            a.copy_(b)
        for st in gt.opt.state.values():
            for v in st.values():
                if isinstance(v, torch.Tensor):
                    v.zero_()
    for i, hb in enumerate(host):
        l1 = float(eager.step(hb.to(dev())))
        if i % 2 == 0:
            l2 = float(gt.step(pad_batch(hb, Np, Ep).to(dev())))
        else:
            gt.load_unpadded(hb.to(dev()))
            l2 = float(gt.step())
        assert abs(l1 - l2) <= 2e-5 * abs(l1), (i, l1, l2)
    for (k, a), (_, b) in zip(m1.named_parameters(), m2.named_parameters()):
        assert torch.allclose(a, b, rtol=0, atol=2e-6), k


@pytest.mark.parametrize("kind,widths", [("zinc", (21, 4)), ("counting", None)])
def test_fused_collation_into_padded_buffers_is_bit_exact(kind, widths):
    """gnnml3_collate writing straight into the static (padded) buffers of a captured step -- from the wire format and from the
    HBM-resident pool -- equals train.pad_batch of the host-collated batch, field by field (integer work; floats are copied)."""
    from gnn_matlang_b200.batch import Batch, CompactBatch
    from gnn_matlang_b200.synthetic import DeviceDataset, GraphPool
    from gnn_matlang_b200.train import pad_batch
    pool = GraphPool(kind, 64, seed=9)
    idx = np.random.default_rng(5).integers(0, 64, 1500)          # more than 1024 graphs: two chunks of the offset scan
    hb = pool.collate(idx)
    N, E = hb.x.size(0), hb.edge_index2.size(1)
    Np, Ep = N + 37, E + 211
    ref = pad_batch(hb, Np, Ep)
    d = dev()

    def static():
        return Batch(x=torch.full((Np, hb.x.size(1)), 7.0, device=d), edge_index2=torch.full((2, Ep), -1, dtype=torch.int64, device=d),
                     edge_attr2=torch.full((Ep, hb.edge_attr2.size(1)), 7.0, device=d), batch=torch.full((Np,), -1, dtype=torch.int64, device=d),
                     graph_ptr=torch.full((len(idx) + 2,), -1, dtype=torch.int32, device=d), y=torch.zeros_like(hb.y, device=d),
                     num_graphs=len(idx) + 1)

    a = CompactBatch.from_batch(hb, widths).pin_memory().to(d).expand_into(static())
    b = DeviceDataset(pool, d).collate_into(torch.from_numpy(idx).pin_memory(), static())
    for out in (a, b):
        for k in FIELDS:
            assert torch.equal(getattr(out, k).cpu(), getattr(ref, k)), k
