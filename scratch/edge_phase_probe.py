"""Cycles per phase of the tcgen05 edge-MLP kernels (needs the -DEMT_PROFILE build: GNNML3_LIB=.../libgnnml3_b200_emtprof.so)."""
import ctypes, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gnn_matlang_b200 import ops, _lib
d = torch.device("cuda:0")
E, K = 1116845, int(sys.argv[1]) if len(sys.argv) > 1 else 8
g = torch.Generator().manual_seed(0)
ea = torch.randn(E, K, generator=g).to(d); go = torch.randn(E, K, generator=g).to(d)
w1, w2, w3 = [(torch.randn(2 * K, K, generator=g) / 3).to(d) for _ in range(3)]
w4 = (torch.randn(K, 4 * K, generator=g) / 6).to(d)
lib = _lib.load()
buf = (ctypes.c_ulonglong * 16)()
names = ["load+st", "round1", "activations", "round2", "dpre4+round3", "grads(+dea)", "bar", "phase2", "out"]
def run(fn, tag):
    for _ in range(3): fn()
    lib.gnnml3_emt_debug_fetch(buf, 1)
    fn()
    lib.gnnml3_emt_debug_fetch(buf, 1)
    n = max(int(buf[10]), 1)
    tot = buf[9] / n
    print("%-10s CTAs %d  cycles/CTA(group 0) %.0f :" % (tag, n, tot), "  ".join("%s %.0f (%.0f%%)" % (names[i], buf[i] / n, 100.0 * buf[i] / max(buf[9], 1)) for i in range(9) if buf[i]))
run(lambda: ops.edge_mlp_fwd(ea, None, w1, w2, w3, w4), "fwd")
run(lambda: ops.edge_mlp_bwd(ea, None, go, w1, w2, w3, w4, need_dea=False), "bwd")
run(lambda: ops.edge_mlp_bwd(ea, None, go, w1, w2, w3, w4, need_dea=True), "bwd+dea")
