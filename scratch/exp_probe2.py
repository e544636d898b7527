import os, sys, time, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gnn_matlang_b200.libs.utils import SpectralDesign
from gnn_matlang_b200.models import GNNML3
from gnn_matlang_b200.synthetic import ExpPool, design_and_collate
from gnn_matlang_b200.train import DesignFeeder, GraphedTrainer, pad_batch
dev = torch.device("cuda:0")
pool = ExpPool(); sd = SpectralDesign(recfield=1, dv=2, nfreq=5, adddegree=True)
rng = np.random.default_rng(13); B = 50
raws = [pool.draw_raw(rng, B) for _ in range(8)]
n_tot = [int(r["node_ptr"][-1]) for r in raws]; e_tot = [int(r["node_ptr"][-1]) + int(r["edge_ptr"][-1]) for r in raws]
Np, Ep = int(max(n_tot) * 1.15) + 64, int(max(e_tot) * 1.15) + 64
raws = [{k: (v.to(dev) if (isinstance(v, torch.Tensor) and k not in ("node_ptr",)) else v) for k, v in r.items()} for r in raws]
ex = design_and_collate(raws[0], sd, dev)
exh = type(ex)(**{k: (v.cpu() if isinstance(v, torch.Tensor) else v) for k, v in ex.__dict__.items()})
torch.manual_seed(0)
gt = GraphedTrainer(GNNML3("exp", pool.K, pool.F).to(dev), pad_batch(exh, Np, Ep), loss="bce")
feeder = DesignFeeder(sd, dev, depth=3)
for j in range(3): feeder.prefetch(raws[j])
T = {"get": 0.0, "prefetch": 0.0, "load": 0.0, "step": 0.0}
torch.cuda.synchronize(); t00 = time.perf_counter()
for i in range(50):
    t0 = time.perf_counter(); bt = feeder.get(); t1 = time.perf_counter()
    feeder.prefetch(raws[(i + 3) % 8]); t2 = time.perf_counter()
    gt.load_unpadded(bt); t3 = time.perf_counter()
    gt.step(); t4 = time.perf_counter()
    T["get"] += t1 - t0; T["prefetch"] += t2 - t1; T["load"] += t3 - t2; T["step"] += t4 - t3
torch.cuda.synchronize(); tot = time.perf_counter() - t00
print("per step: total %.3f ms |" % (tot / 50 * 1e3), {k: round(v / 50 * 1e3, 3) for k, v in T.items()})
# inside prefetch: time design_batch pieces
import gnn_matlang_b200.libs.utils as U
t0 = time.perf_counter()
for i in range(20):
    out = sd.design_batch(raws[i % 8]["edge_index"], raws[i % 8]["edge_ptr"], raws[i % 8]["node_ptr"], device=dev, global_ids=True)
torch.cuda.synchronize(); print("design_batch alone (sequential, incl. sync): %.3f ms" % ((time.perf_counter() - t0) / 20 * 1e3))
t0 = time.perf_counter()
for i in range(20):
    out = sd.design_batch(raws[i % 8]["edge_index"], raws[i % 8]["edge_ptr"], raws[i % 8]["node_ptr"], device=dev, global_ids=True, max_entries=Ep)
t1 = time.perf_counter(); torch.cuda.synchronize(); print("design_batch(max_entries) host enqueue: %.3f ms, with device: %.3f ms" % ((t1 - t0) / 20 * 1e3, (time.perf_counter() - t0) / 20 * 1e3))
