"""cProfile of the host side of one training step (enqueue only)."""
import cProfile, pstats, sys, os, io
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gnn_matlang_b200.graph import set_range_check
from gnn_matlang_b200.models import GNNML3
from gnn_matlang_b200.synthetic import GraphPool
from gnn_matlang_b200.train import Trainer
dev = torch.device("cuda:0")
set_range_check(False)
pool = GraphPool("zinc", 1024, seed=1)
rng = np.random.default_rng(7)
ring = [pool.draw(rng, 8192).to(dev, non_blocking=False) for _ in range(3)]
torch.manual_seed(0)
model = GNNML3("zinc", pool.K, pool.F).to(dev)
tr = Trainer(model, loss="l1", lr=1e-3)
for i in range(5):
    tr.step(ring[i % 3].fresh())
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for i in range(20):
    tr.step(ring[i % 3].fresh())
pr.disable()
torch.cuda.synchronize()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(28)
print(s.getvalue()[:6000])
