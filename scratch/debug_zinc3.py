import sys; sys.path.insert(0,'tests'); sys.path.insert(0,'.')
import torch, numpy as np
from oracle import gnnml3_oracle as O
import test_gpu_model as T
from gnn_matlang_b200.batch import collate
from gnn_matlang_b200.models import GNNML3
from gnn_matlang_b200.pool import global_add_pool
from gnn_matlang_b200.libs.spect_conv import _LinearFn
dev = torch.device('cuda:0')
cfg = "zinc"
g = torch.Generator().manual_seed(11)
graphs = T._random_graphs(cfg, 24, g)
ne, ninp = graphs[0]["edge_attr2"].shape[1], graphs[0]["x"].shape[1]
torch.manual_seed(5)
ref = O.OracleGNNML3(cfg, ne, ninp)
model = GNNML3(cfg, ne, ninp); model.load_state_dict(ref.state_dict()); model = model.to(dev)
ob = O.collate(graphs); hb = collate(graphs).to(dev)
xs = [ob["x"]]
for l in range(4):
    xo = getattr(ref, "conv%d" % (l + 1))(xs[-1], ob["edge_index2"], ob["edge_attr2"]); xo.retain_grad(); xs.append(xo)
pooled = O.global_add_pool(xs[-1], ob["batch"], 24); pooled.retain_grad()
out_r = ref.fc2(torch.relu(ref.fc1(pooled)))
y = ob["y"].float()
torch.nn.functional.l1_loss(out_r, y, reduction="sum").backward()
def cmp(name, a, b):
    a = a.detach().cpu().double(); b = b.detach().double()
    d = (a - b).abs(); print("%-28s err %.2e max %.2e rel %.1e  nbad(>1e-4rel) %d / %d" % (name, d.max(), b.abs().max(), d.max() / b.abs().max(), int((d > 1e-4 * b.abs().max()).sum()), d.numel()))
# chained GPU with retained intermediates
gx = [hb.x]
for l in range(4):
    o = getattr(model, "conv%d" % (l + 1))(gx[-1], hb.edge_index2, hb.edge_attr2); o.retain_grad(); gx.append(o)
pg = global_add_pool(gx[-1], hb.batch, 24); pg.retain_grad()
h = _LinearFn.apply(pg, model.fc1.weight.t(), model.fc1.bias, 0)
og = _LinearFn.apply(torch.relu(h), model.fc2.weight.t(), model.fc2.bias, 0)
torch.nn.functional.l1_loss(og, hb.y.float(), reduction="sum").backward()
for l in range(1, 5):
    cmp("x%d" % l, gx[l], xs[l])
    print("   zero-pattern equal:", bool(((gx[l].cpu() == 0) == (xs[l] == 0)).all()), " mismatches:", int(((gx[l].cpu() == 0) != (xs[l] == 0)).sum()))
cmp("pooled", pg, pooled); cmp("dpooled", pg.grad, pooled.grad)
for l in range(4, 0, -1):
    cmp("d x%d" % l, gx[l].grad, xs[l].grad)
    d = (gx[l].grad.cpu() - xs[l].grad).abs(); 
    bad = d > 1e-4 * xs[l].grad.abs().max()
    print("   bad cols:", sorted(set(bad.nonzero()[:, 1].tolist()))[:40], " bad rows:", len(set(bad.nonzero()[:, 0].tolist())))
print("y dtype", hb.y.dtype, ob["y"].dtype, "y equal", torch.equal(hb.y.cpu().float(), y))
