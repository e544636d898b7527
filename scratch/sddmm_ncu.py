import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gnn_matlang_b200 import ops
from gnn_matlang_b200.synthetic import GraphPool
pool = GraphPool("zinc", 2048, seed=0)
hb = pool.draw(np.random.default_rng(0), 8192)
d = torch.device("cuda:0")
ei = hb.edge_index2.to(d)
N = hb.x.size(0)
plan = ops.csr_build(ei, N)
E = ei.size(1)
K, Fi, Fo = 8, 32, 30
g = torch.Generator().manual_seed(0)
x = torch.randn(N, Fi, generator=g).to(d)
gc = ops.aligned_rows(torch.randn(N, Fo, generator=g).to(d))
W = (torch.randn(K, Fi, Fo, generator=g) / 6).to(d)
for _ in range(4):
    ops.fused_sddmm(plan["rowptr"], plan["col"], x, gc, W, E, win=plan["win"])
torch.cuda.synchronize()
