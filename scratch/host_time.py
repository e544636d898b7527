"""Pure host enqueue time per training step: 3 steps after a sync (fits the launch queue), no sync inside."""
import sys, os, time
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
res = []
for rep in range(5):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(3):
        tr.step(ring[i % 3].fresh())
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    res.append(((t1 - t0) / 3 * 1e3, (t2 - t0) / 3 * 1e3))
print("host enqueue ms/step, total ms/step:", ["%.2f / %.2f" % r for r in res])
