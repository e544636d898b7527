"""Times the fused layer kernel on one ZINC-shaped batch (8192 graphs). Usage: python scratch/fused_probe.py [reps]"""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gnn_matlang_b200 import ops
from gnn_matlang_b200.synthetic import GraphPool

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
B = int(os.environ.get("PROBE_B", "8192"))
pool = GraphPool("zinc", 2048, seed=0)
hb = pool.draw(np.random.default_rng(0), B)
d = torch.device("cuda:0")
ei = hb.edge_index2.to(d)
N = hb.x.size(0)
plan = ops.csr_build(ei, N)
E = ei.size(1)
K, Fi, Fo, G = 8, 32, 30, 2
g = torch.Generator().manual_seed(0)
x = torch.randn(N, Fi, generator=g).to(d)
ea = torch.randn(E, K, generator=g).to(d)
W = (torch.randn(K * Fi, Fo, generator=g) / 16).to(d)
wg = (torch.randn(Fi, 2 * G, generator=g) / 6).to(d)
b = torch.zeros(Fo, device=d); bs = torch.zeros(2 * G, device=d)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=d)

def timeit(fn, name):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e3)
    ts.sort()
    print("%-28s median %8.1f us   min %8.1f us" % (name, ts[len(ts) // 2], ts[0]), flush=True)
    if os.environ.get("GNNML3_FUSED_DEBUG") == "1" and name.startswith("fused"):
        import ctypes
        from gnn_matlang_b200 import _lib
        buf = (ctypes.c_ulonglong * 16)()
        _lib.load().gnnml3_fused_debug_counters(buf, 1)
        c = [float(v) / (reps + 1) for v in buf]
        nagg = c[2] and round(c[2] / (c[5] or 1))
        print("   per launch (cycles summed over CTAs): agg gather %.3g wait %.3g total %.3g | mma wait_full %.3g wait_tempty %.3g total %.3g | epi wait %.3g total %.3g"
              % tuple(c[:8]))
        print("   MMA-thread loop cycles per CTA: mean %.3g  max %.3g  min %.3g (last launch)" % (c[5] / 148, buf[11], buf[12]))
        print("   mma issue+commit cycles per CTA %.3g (of total %.3g)" % (c[10] / 148, c[5] / 148))
        print("   fractions: agg gather %.2f wait %.2f | mma wait_full %.2f wait_tempty %.2f busy %.2f | epi wait %.2f"
              % (c[0] / c[2], c[1] / c[2], c[3] / c[5], c[4] / c[5], 1 - (c[3] + c[4]) / c[5], c[6] / c[7]))

print("N=%d E=%d" % (N, E))
timeit(lambda: ops.fused_agg_proj(plan["rowptr"], plan["col"], None, ea, x, W, bias=b, S=x, self_mode=1, Bself=wg, bias_s=bs, G=G, epilogue=1), "fused fwd (ml3)")
timeit(lambda: ops.fused_agg_proj(plan["rowptr"], plan["col"], None, ea, x, W, bias=b, epilogue=0), "fused fwd (plain)")
timeit(lambda: ops.fused_agg_proj(plan["rowptrT"], plan["colT"], plan["permT"], ea, x, W, epilogue=0), "fused transposed (permT)")
timeit(lambda: ops.spmm_k(plan["rowptr"], plan["col"], None, ea, x), "spmm_k")
H = ops.spmm_k(plan["rowptr"], plan["col"], None, ea, x)
timeit(lambda: ops.gemm_nn(H, W, b), "gemm_nn (tc)")
