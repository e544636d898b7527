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
# oracle with intermediates
xs = [ob["x"]]
for l in range(4):
    xo = getattr(ref, "conv%d" % (l + 1))(xs[-1], ob["edge_index2"], ob["edge_attr2"]); xo.retain_grad(); xs.append(xo)
pooled = O.global_add_pool(xs[-1], ob["batch"], 24); pooled.retain_grad()
out_r = ref.fc2(torch.relu(ref.fc1(pooled)))
y = ob["y"].float()
torch.nn.functional.l1_loss(out_r, y, reduction="sum").backward()
def cmp(name, a, b):
    a = a.detach().cpu().double(); b = b.detach().double()
    d = (a - b).abs(); print("%-28s err %.2e max %.2e rel %.1e  nbad(>1e-4rel) %d" % (name, d.max(), b.abs().max(), d.max() / b.abs().max(), int((d > 1e-4 * b.abs().max()).sum())))
# head + pool on GPU from oracle x4
x4 = xs[4].detach().to(dev).requires_grad_(True)
pg = global_add_pool(x4, hb.batch, 24); pg.retain_grad()
h = _LinearFn.apply(pg, model.fc1.weight.t(), model.fc1.bias, 0)
og = _LinearFn.apply(torch.relu(h), model.fc2.weight.t(), model.fc2.bias, 0)
torch.nn.functional.l1_loss(og, hb.y.float(), reduction="sum").backward()
cmp("pooled", pg, pooled); cmp("out", og, out_r); cmp("d pooled", pg.grad, pooled.grad); cmp("d x4", x4.grad, xs[4].grad)
# each layer in isolation with oracle inputs / upstream grads
for l in range(4, 0, -1):
    layer = getattr(model, "conv%d" % l); rl = getattr(ref, "conv%d" % l)
    layer.zero_grad()
    xin = xs[l - 1].detach().to(dev).requires_grad_(l > 1)
    o = layer(xin, hb.edge_index2, hb.edge_attr2)
    cmp("L%d out" % l, o, xs[l])
    o.backward(xs[l].grad.to(dev))
    if l > 1: cmp("L%d dx" % l, xin.grad, xs[l - 1].grad)
    for (k, p), (_, pr) in zip(layer.named_parameters(), rl.named_parameters()):
        cmp("L%d %s" % (l, k), p.grad, pr.grad)
    # mask agreement
    from gnn_matlang_b200.graph import get_plan, sorted_edge_attr
print("x4 exact zeros frac", float((xs[4][:, :30] == 0).float().mean()))
