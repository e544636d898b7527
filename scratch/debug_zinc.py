import sys; sys.path.insert(0,'tests'); sys.path.insert(0,'.')
import torch, numpy as np
from oracle import gnnml3_oracle as O
import test_gpu_model as T
from gnn_matlang_b200.batch import collate
from gnn_matlang_b200.models import GNNML3
dev = torch.device('cuda:0')
for cfg in ["zinc"]:
    g = torch.Generator().manual_seed(11)
    graphs = T._random_graphs(cfg, 24, g)
    ne, ninp = graphs[0]["edge_attr2"].shape[1], graphs[0]["x"].shape[1]
    torch.manual_seed(5)
    ref = O.OracleGNNML3(cfg, ne, ninp)
    model = GNNML3(cfg, ne, ninp); model.load_state_dict(ref.state_dict())
    ob = O.collate(graphs)
    out_r = ref(ob); y = ob["y"].float()
    loss_r = torch.nn.functional.l1_loss(out_r, y.expand_as(out_r), reduction="sum"); loss_r.backward()
    model = model.to(dev); hb = collate(graphs).to(dev)
    out = model(hb)
    loss = torch.nn.functional.l1_loss(out, hb.y.float().expand_as(out), reduction="sum"); loss.backward()
    print("min |out-y|", (out_r - y).abs().min().item())
    print("sign agree", torch.equal(torch.sign(out.cpu()-y), torch.sign(out_r-y)))
    for (k, p), (_, pr) in zip(model.named_parameters(), ref.named_parameters()):
        d = (p.grad.cpu().double() - pr.grad.double()).abs().max().item(); m = pr.grad.abs().max().item()
        print("%-22s err %.2e max %.2e rel %.1e" % (k, d, m, d / max(m, 1e-30)))
    # same with MSE loss (smooth)
    model.zero_grad(); ref.zero_grad()
    out_r = ref(ob); ((out_r - y) ** 2).sum().backward()
    out = model(hb); ((out - hb.y.float()) ** 2).sum().backward()
    print("---- MSE loss")
    for (k, p), (_, pr) in zip(model.named_parameters(), ref.named_parameters()):
        d = (p.grad.cpu().double() - pr.grad.double()).abs().max().item(); m = pr.grad.abs().max().item()
        print("%-22s err %.2e max %.2e rel %.1e" % (k, d, m, d / max(m, 1e-30)))
