"""Fused dH + SDDMM kernel on one ZINC-shaped batch: staged source windows vs global gathers."""
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
flush = torch.empty(256 << 20, dtype=torch.uint8, device=d)
def timeit(fn, name, reps=10):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e) * 1e3)
    ts.sort(); print("%-34s median %8.1f us min %8.1f us" % (name, ts[len(ts) // 2], ts[0]), flush=True)
print("N=%d E=%d" % (N, E))
timeit(lambda: ops.fused_sddmm(plan["rowptr"], plan["col"], x, gc, W, E, win=plan["win"]), "fused_sddmm staged windows")
timeit(lambda: ops.fused_sddmm(plan["rowptr"], plan["col"], x, gc, W, E, win=None), "fused_sddmm global gathers")
