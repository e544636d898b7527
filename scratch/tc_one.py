import sys; sys.path.insert(0,'.')
import torch
from gnn_matlang_b200 import ops
dev=torch.device('cuda:0'); g=torch.Generator().manual_seed(0)
M=189413; Kc=256; Nc=32
A=torch.randn(M,Kc,generator=g).to(dev); B=torch.randn(Kc,Nc,generator=g).to(dev)
for _ in range(3): ops.gemm_nn_tc(A,B,None)
torch.cuda.synchronize()
