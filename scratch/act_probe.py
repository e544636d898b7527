"""Times gnnml3_ml3_act_bwd_y (+ the column-sum finish) on the ZINC step's shape, ten calls replayed from a CUDA graph
(GNNML3_ACT_GENERAL=1 selects the general kernel)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gnn_matlang_b200 import ops
d = torch.device("cuda:0")
N, Fo, G = 189413, 30, 2
g = torch.Generator().manual_seed(0)
ys = [torch.randn(N, 32, generator=g).to(d) for _ in range(10)]
aux = [torch.randn(N, 4, generator=g).to(d) for _ in range(10)]
gys = [torch.randn(N, 32, generator=g).to(d) for _ in range(10)]
ops.ml3_act_bwd_y(ys[0], aux[0], gys[0], Fo, G); torch.cuda.synchronize()
st = torch.cuda.Stream()
with torch.cuda.stream(st):
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr, stream=st):
        outs = [ops.ml3_act_bwd_y(ys[i], aux[i], gys[i], Fo, G) for i in range(10)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=d)
ts = []
for _ in range(12):
    flush.zero_()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); gr.replay(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e) * 1e2)
ts.sort(); print("ml3_act_bwd_y + colsum finish, per call inside a graph of 10: median %.1f us min %.1f us (83 MB per call)" % (ts[len(ts) // 2], ts[0]))
