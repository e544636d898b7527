import sys, time; sys.path.insert(0,'.')
import torch
from gnn_matlang_b200 import ops, _lib
dev=torch.device('cuda:0')
g=torch.Generator().manual_seed(0)
def timeit(f, n=20):
    for _ in range(3): f()
    torch.cuda.synchronize(); s=torch.cuda.Event(enable_timing=True); e=torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): f()
    e.record(); torch.cuda.synchronize(); return s.elapsed_time(e)/n*1e3
M=189413
flush=torch.empty(300_000_000//4, device=dev)
for (Kc,Nc,name) in [(256,30,"H*W"),(244,32,"dx"),(32,256,"dH(Kc30->32)"),(640,64,"sweepF64")]:
    A=torch.randn(M,Kc,generator=g).to(dev); B=torch.randn(Kc,Nc,generator=g).to(dev)
    ref=(A.double()@B.double())
    for chunk in (1,2,4,8,1000):
        out=ops.gemm_nn_tc(A,B,None,chunk_kblocks=chunk)
        err=((out.double()-ref).abs().max()/ref.abs().max()).item()
        t=timeit(lambda: ops.gemm_nn_tc(A,B,None,chunk_kblocks=chunk))
        print("%-14s tc chunk=%-4d  %.1f us   rel-to-max err %.2e   (A stream %.0f GB/s)"%(name,chunk,t,err,(M*Kc*4+M*Nc*4)/t/1e3))
    import os
    out=torch.empty(M,Nc,device=dev)
    lib=_lib.load()
    def old():
        _lib.check(lib.gnnml3_gemm_nn(A.data_ptr(),Kc,B.data_ptr(),Nc,None,out.data_ptr(),Nc,M,Nc,Kc,0,0,torch.cuda.current_stream().cuda_stream),"x")
    t=timeit(old); err=((out.double()-ref).abs().max()/ref.abs().max()).item()
    print("%-14s mma.sync 3xTF32 %.1f us   err %.2e"%(name,t,err))
    t=timeit(lambda: torch.matmul(A,B)); print("%-14s torch fp32 matmul (cuBLAS) %.1f us"%(name,t))
