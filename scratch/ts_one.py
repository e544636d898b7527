"""One forward launch of the fused layer kernel on the ZINC-shaped batch (for ncu): python scratch/ts_one.py gen [dx]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gnn_matlang_b200 import _lib, ops
from gnn_matlang_b200.synthetic import GraphPool
gen = int(sys.argv[1]); dx = len(sys.argv) > 2
pool = GraphPool("zinc", 2048, seed=0)
hb = pool.draw(np.random.default_rng(0), 8192)
d = torch.device("cuda:0")
ei = hb.edge_index2.to(d); N = hb.x.size(0)
plan = ops.csr_build(ei, N); E = ei.size(1); K = pool.K
Fi, Fo, G = 32, 30, 2
g = torch.Generator().manual_seed(0)
x = torch.randn(N, Fi, generator=g).to(d); ea = torch.randn(E, K, generator=g).to(d)
W = (torch.randn(K * Fi, Fo, generator=g) / 16).to(d); wg = (torch.randn(Fi, 2 * G, generator=g) / 6).to(d)
b = torch.zeros(Fo, device=d); bs = torch.zeros(2 * G, device=d)
_lib.load().gnnml3_fused_set_ts(gen)   # 1 tensor-memory kernel, 0 round-1 kernel
for _ in range(2):
    if dx:
        ops.fused_agg_proj(plan["rowptrT"], plan["colT"], plan["permT"], ea, x, W, epilogue=0, win=plan["winT"])
    else:
        ops.fused_agg_proj(plan["rowptr"], plan["col"], None, ea, x, W, bias=b, S=x, self_mode=1, Bself=wg, bias_s=bs, G=G, epilogue=1, win=plan["win"])
torch.cuda.synchronize()
