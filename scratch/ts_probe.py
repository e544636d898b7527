"""Times the fused layer kernel variants on one ZINC-shaped batch (8192 graphs): tensor-memory kernel with staged windows /
global gathers, and the round-1 shared-memory-plane kernel.  Usage: python scratch/ts_probe.py [reps]
With GNNML3_LIB=gnn_matlang_b200/libgnnml3_b200_prof.so GNNML3_FUSED_DEBUG=1 it also prints the cycle counters."""
import ctypes, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gnn_matlang_b200 import _lib, ops
from gnn_matlang_b200.synthetic import GraphPool

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
B = int(os.environ.get("PROBE_B", "8192"))
kind = os.environ.get("PROBE_KIND", "zinc")
pool = GraphPool(kind, 2048, seed=0)
hb = pool.draw(np.random.default_rng(0), B)
d = torch.device("cuda:0")
ei = hb.edge_index2.to(d)
N = hb.x.size(0)
plan = ops.csr_build(ei, N)
E = ei.size(1)
K = pool.K
Fi, Fo, G = {"zinc": (32, 30, 2), "counting": (32, 16, 16)}.get(kind, (32, 30, 2))
g = torch.Generator().manual_seed(0)
x = torch.randn(N, Fi, generator=g).to(d)
ea = torch.randn(E, K, generator=g).to(d)
W = (torch.randn(K * Fi, Fo, generator=g) / 16).to(d)
wg = (torch.randn(Fi, 2 * G, generator=g) / 6).to(d)
b = torch.zeros(Fo, device=d); bs = torch.zeros(2 * G, device=d)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=d)
lib = _lib.load()
win = plan["win"].cpu().numpy()
wd = win[:, 1] - win[:, 0]
print("N=%d E=%d K=%d tiles=%d window rows: mean %.1f max %d, tiles over 256: %d" % (N, E, K, len(wd), wd.mean(), wd.max(), int((wd > 256).sum())))
bytes_fwd = 4.0 * (N * Fi + E * K + E + (N + 1) + K * Fi * Fo + Fi * 2 * G + Fo + N * (Fo + G) + N * 2 * G)

def counters(tag):
    if os.environ.get("GNNML3_FUSED_DEBUG") != "1":
        return
    buf = (ctypes.c_ulonglong * 16)()
    lib.gnnml3_fused_debug_counters(buf, 1)
    c = [float(v) / (reps + 1) for v in buf]
    if c[2] == 0 or c[5] == 0:
        return
    print("   [%s] per launch, summed over CTAs: agg gather %.3g slot-wait %.3g window-wait %.3g total %.3g | mma wait_full %.3g wait_tempty %.3g total %.3g | epi wait %.3g total %.3g"
          % (tag, c[0], c[1], c[8], c[2], c[3], c[4], c[5], c[6], c[7]))
    print("   per aggregator warp and tile (cycles): gather %.0f (inner loop %.0f for %.1f steps = %.0f per step), dump+slot waits %.0f, stage wait %.0f, total %.0f"
          % (c[0] / 1184 / 10, c[10] / 1184 / 10, c[13] / 1184 / 10, c[10] / max(c[13], 1), c[9] / 1184 / 10, c[8] / 1184 / 10, c[2] / 1184 / 10))
    print("   fractions: agg gather %.2f slot-wait %.2f window-wait %.2f | mma wait_full %.2f wait_tempty %.2f busy %.2f | epi wait %.2f ; MMA loop cycles per CTA %.3g"
          % (c[0] / c[2], c[1] / c[2], c[8] / c[2], c[3] / c[5], c[4] / c[5], 1 - (c[3] + c[4]) / c[5], c[6] / max(c[7], 1), c[5] / 148))

def timeit(fn, name):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e3)
    ts.sort()
    print("%-44s median %8.1f us   min %8.1f us   (fwd bytes / median = %.0f GB/s)" % (name, ts[len(ts) // 2], ts[0], bytes_fwd / ts[len(ts) // 2] / 1e3), flush=True)
    counters(name)

def fwd(win):
    return ops.fused_agg_proj(plan["rowptr"], plan["col"], None, ea, x, W, bias=b, S=x, self_mode=1, Bself=wg, bias_s=bs, G=G, epilogue=1, win=win)
def dxp(win):
    return ops.fused_agg_proj(plan["rowptrT"], plan["colT"], plan["permT"], ea, x, W, epilogue=0, win=win)

for ts_on, label in ((1, "tensor-memory"), (0, "smem planes (round 1)")):
    lib.gnnml3_fused_set_ts(ts_on)
    if ts_on:
        timeit(lambda: fwd(plan["win"]), label + " fwd ml3, staged windows")
        timeit(lambda: dxp(plan["winT"]), label + " transposed+permT, staged")
    timeit(lambda: fwd(None), label + " fwd ml3, global gathers")
    timeit(lambda: dxp(None), label + " transposed+permT, global")
lib.gnnml3_fused_set_ts(1)
print("paths (ts, planes):", ops.fused_path_counts())
