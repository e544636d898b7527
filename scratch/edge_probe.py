import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gnn_matlang_b200 import ops
d = torch.device("cuda:0")
E, K = 1116845, 8
g = torch.Generator().manual_seed(0)
ea = torch.randn(E, K, generator=g).to(d); go = torch.randn(E, K, generator=g).to(d)
w1, w2, w3 = [(torch.randn(2 * K, K, generator=g) / 3).to(d) for _ in range(3)]
w4 = (torch.randn(K, 4 * K, generator=g) / 6).to(d)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=d)
def timeit(fn, name, reps=10):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e) * 1e3)
    ts.sort(); print("%-30s median %8.1f us min %8.1f us" % (name, ts[len(ts) // 2], ts[0]), flush=True)
for gen in (True, False):
    ops.edge_mlp_set_tc(gen)
    tag = "tcgen05 " if gen else "fp32 fma "
    timeit(lambda: ops.edge_mlp_fwd(ea, None, w1, w2, w3, w4), tag + "edge_mlp_fwd")
    timeit(lambda: ops.edge_mlp_bwd(ea, None, go, w1, w2, w3, w4, need_dea=False), tag + "edge_mlp_bwd")
    timeit(lambda: ops.edge_mlp_bwd(ea, None, go, w1, w2, w3, w4, need_dea=True), tag + "edge_mlp_bwd+dea")
