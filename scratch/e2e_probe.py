"""Where does an end-to-end step spend its time: H2D of the wire format, expansion into the captured buffers, graph replay."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gnn_matlang_b200.models import GNNML3
from gnn_matlang_b200.synthetic import GraphPool
from gnn_matlang_b200.train import GraphedTrainer, HostFeeder, pad_batch, padded_shapes
d = torch.device("cuda:0")
pool = GraphPool("zinc", 2048, seed=0)
rng = np.random.default_rng(0)
host = [pool.draw(rng, 8192) for _ in range(4)]
Np, Ep = padded_shapes(host)
torch.manual_seed(0)
m = GNNML3("zinc", pool.K, pool.F).to(d)
gt = GraphedTrainer(m, pad_batch(host[0], Np, Ep), loss="l1")
feeder = HostFeeder(d, onehot_widths=(21, 4))
for h in host: feeder.compact(h)
def ev(): return torch.cuda.Event(enable_timing=True)
for rep in range(3):
    feeder.prefetch(host[rep % 4]); torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0, e1, e2, e3 = ev(), ev(), ev(), ev()
    e0.record(); cb = feeder.get_compact(); e1.record(); gt.load_compact(cb); e2.record(); gt.step(); e3.record()
    t1 = time.perf_counter(); torch.cuda.synchronize()
    print("get %.3f ms  load_compact %.3f ms  step %.3f ms | host enqueue %.3f ms" % (e0.elapsed_time(e1), e1.elapsed_time(e2), e2.elapsed_time(e3), (t1 - t0) * 1e3))
    e0, e1 = ev(), ev()
    e0.record(); feeder.prefetch(host[(rep + 1) % 4]); feeder.copy_stream.synchronize(); e1.record(); torch.cuda.synchronize()
    cbh = feeder.compact(host[0])
    print("   prefetch (H2D %.1f MB) wall %.3f ms; pinned: %s" % (cbh.nbytes() / 1e6, e0.elapsed_time(e1), {k: v.is_pinned() for k, v in cbh._tensors().items()}))

def loop(order, steps=20):
    feeder.prefetch(host[0])
    for i in range(3):
        b = feeder.get_compact(); feeder.prefetch(host[(i + 1) % 4]); gt.load_compact(b); float(gt.step().item())
    torch.cuda.synchronize()
    loss_host = [torch.empty(1, dtype=torch.float32).pin_memory() for _ in range(2)]
    loss_ev = [torch.cuda.Event(), torch.cuda.Event()]
    e0, e1 = ev(), ev()
    e0.record()
    for i in range(steps):
        b = feeder.get_compact()
        if order == "before":
            feeder.prefetch(host[(i + 4) % 4])
        gt.load_compact(b)
        lt = gt.step()
        if order == "after":
            feeder.prefetch(host[(i + 4) % 4])
        loss_host[i & 1].copy_(lt.reshape(1), non_blocking=True)
        loss_ev[i & 1].record()
        if i > 0:
            loss_ev[(i - 1) & 1].synchronize()
    loss_ev[(steps - 1) & 1].synchronize()
    e1.record(); torch.cuda.synchronize()
    print("loop, prefetch %s the step: %.3f ms per step" % (order, e0.elapsed_time(e1) / steps))
loop("before"); loop("after"); loop("before")
