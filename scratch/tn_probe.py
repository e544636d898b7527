import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gnn_matlang_b200 import ops
d = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
M = 189413
A = torch.randn(M, 32, generator=g).to(d); B = torch.randn(M, 256, generator=g).to(d)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=d)
ref = (A.double().t() @ B.double())
out = ops.gemm_tn(A, B)
err = (out.double() - ref).abs().max().item() / ref.abs().max().item()
print("max err / max|ref| = %.3e" % err)
def timeit(fn, name, reps=10):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e) * 1e3)
    ts.sort(); print("%-20s median %8.1f us min %8.1f us" % (name, ts[len(ts) // 2], ts[0]), flush=True)
timeit(lambda: ops.gemm_tn(A, B), "gemm_tn 32x256")
