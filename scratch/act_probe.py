import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gnn_matlang_b200 import ops
d = torch.device("cuda:0")
N, Fo, G = 189413, 30, 2
g = torch.Generator().manual_seed(0)
y = torch.randn(N, 32, generator=g).to(d); aux = torch.randn(N, 4, generator=g).to(d); gy = torch.randn(N, 32, generator=g).to(d)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=d)
ts = []
for _ in range(12):
    flush.zero_()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); ops.ml3_act_bwd_y(y, aux, gy, Fo, G); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e) * 1e3)
ts.sort(); print("ml3_act_bwd_y (+colsum finish): median %.1f us min %.1f us" % (ts[len(ts) // 2], ts[0]))
