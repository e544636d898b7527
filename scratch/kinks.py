import sys; sys.path.insert(0,'tests'); sys.path.insert(0,'.')
import torch, numpy as np
import torch.nn.functional as F
from oracle import gnnml3_oracle as O
import test_gpu_model as T
cfg = "zinc"
g = torch.Generator().manual_seed(11)
graphs = T._random_graphs(cfg, 24, g)
ne, ninp = graphs[0]["edge_attr2"].shape[1], graphs[0]["x"].shape[1]
torch.manual_seed(5)
ref = O.OracleGNNML3(cfg, ne, ninp)
ob = O.collate(graphs)
x = ob["x"]; ei = ob["edge_index2"]; ea = ob["edge_attr2"]
for l in range(4):
    L = getattr(ref, "conv%d" % (l+1)); p = dict(L.named_parameters())
    pre1 = ea @ p["fc1_1.weight"].t()
    tmp = torch.cat([F.relu(pre1), torch.tanh(ea @ p["fc1_2.weight"].t()) * torch.tanh(ea @ p["fc1_3.weight"].t())], 1)
    pre4 = tmp @ p["fc1_4.weight"].t()
    ea2 = F.relu(pre4)
    c = O.spectconv_forward(x, ei, ea2, p["conv1.weight"], p["conv1.bias"])
    for name, v in (("pre1", pre1), ("pre4", pre4), ("c", c)):
        a = v.detach().abs(); m = a.max()
        print("L%d %-5s n=%7d  max %.2e  #<1e-6*max: %d  #<1e-5*max: %d  #==0: %d" % (l+1, name, a.numel(), m, int((a < 1e-6*m).sum()), int((a < 1e-5*m).sum()), int((a == 0).sum())))
    x = L(x, ei, ea)
