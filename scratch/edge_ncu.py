"""One ZINC-step sized call of the tcgen05 edge-MLP kernels (forward, backward without and with d ea): the launches ncu captures."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gnn_matlang_b200 import ops
d = torch.device("cuda:0")
E, K = 1116845, int(sys.argv[1]) if len(sys.argv) > 1 else 8
g = torch.Generator().manual_seed(0)
ea = torch.randn(E, K, generator=g).to(d); go = torch.randn(E, K, generator=g).to(d)
w1, w2, w3 = [(torch.randn(2 * K, K, generator=g) / 3).to(d) for _ in range(3)]
w4 = (torch.randn(K, 4 * K, generator=g) / 6).to(d)
for _ in range(3):
    ops.edge_mlp_fwd(ea, None, w1, w2, w3, w4)
    ops.edge_mlp_bwd(ea, None, go, w1, w2, w3, w4, need_dea=False)
    ops.edge_mlp_bwd(ea, None, go, w1, w2, w3, w4, need_dea=True)
torch.cuda.synchronize()
