"""Hardware data-parallel parity (SURVEY.md section 4 item iv): 2 ranks (NCCL, one per GPU) each training on HALF of every
global batch must end with the same parameters as 1 rank training on the full batches -- the CUDA model, the gradient SUM
all-reduce of train.Trainer and Adam, two optimisation steps.  Needs 2 GPUs (skipped otherwise; the committed log of a
2-GPU run is profiles/r02_dp_parity_2gpu.txt)."""
import os
import tempfile

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

B, STEPS = 64, 2


def _batches(pool):
    rng = np.random.default_rng(5)
    return [rng.integers(0, len(pool.n), B) for _ in range(STEPS)]


def _worker(rank, world, port, out_path):
    import torch.distributed as dist
    from gnn_matlang_b200.models import GNNML3
    from gnn_matlang_b200.synthetic import GraphPool
    from gnn_matlang_b200.train import Trainer
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    pool = GraphPool("zinc", 96, seed=9)
    torch.manual_seed(0)
    model = GNNML3("zinc", pool.K, pool.F).to(dev)
    tr = Trainer(model, loss="l1", lr=1e-3, distributed=True)
    per = B // world
    grads = None
    for idx in _batches(pool):
        tr.step(pool.collate(idx[rank * per:(rank + 1) * per]).to(dev))
        if grads is None:                  # all-reduced gradient of the first step (a view of the reduced bucket)
            grads = {k: p.grad.detach().cpu().clone() for k, p in model.named_parameters()}
    torch.cuda.synchronize()
    if rank == 0:
        torch.save(({k: v.detach().cpu() for k, v in model.state_dict().items()}, grads), out_path)
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_on_half_batches_equal_one_rank_on_full_batches():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from gnn_matlang_b200.models import GNNML3
    from gnn_matlang_b200.synthetic import GraphPool
    from gnn_matlang_b200.train import Trainer
    with tempfile.TemporaryDirectory() as d:
        out = os.path.join(d, "dp.pt")
        port = 29500 + os.getpid() % 2000
        mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
        dp, dp_grads = torch.load(out)
    dev = torch.device("cuda", 0)
    pool = GraphPool("zinc", 96, seed=9)
    torch.manual_seed(0)
    model = GNNML3("zinc", pool.K, pool.F).to(dev)
    init = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    tr = Trainer(model, loss="l1", lr=1e-3, distributed=False)
    gworst = 0.0
    for s, idx in enumerate(_batches(pool)):
        tr.step(pool.collate(idx).to(dev))
        if s == 0:
            # the all-reduced gradient of two half-batches IS the full-batch gradient (SUM losses): only the summation order
            # differs.  FP32 bar 1e-5 relative to the tensor's scale.
            for k, p in model.named_parameters():
                a, b = dp_grads[k].double(), p.grad.detach().cpu().double()
                err, scale = (a - b).abs().max().item(), b.abs().max().item()
                gworst = max(gworst, err / max(scale, 1e-30))
                assert err <= 1e-5 * scale + 1e-12, "grad %s: 2-rank vs 1-rank differ by %.3e (max|grad| %.3e)" % (k, err, scale)
    worst = 0.0
    for k, v in model.state_dict().items():
        a, b = dp[k].double(), v.detach().cpu().double()
        moved = (b - init[k].double()).abs().max().item()
        assert moved > 0, "parameter %s did not train" % k
        # Adam normalises every element by its own gradient history, so elements whose gradient nearly cancels amplify the
        # 1e-6 summation-order noise: the parameters are held to 10 % of the update size (2 steps x lr), the gradients above
        # to the FP32 bar
        err = (a - b).abs().max().item()
        worst = max(worst, err / moved)
        assert err <= 0.1 * moved, "%s: 2-rank vs 1-rank differ by %.3e (update size %.3e)" % (k, err, moved)
    print("2-rank vs 1-rank: first-step gradients worst |diff| / max|grad| = %.3e; parameters after %d steps worst |diff| / |update| "
          "= %.3e" % (gworst, STEPS, worst))
