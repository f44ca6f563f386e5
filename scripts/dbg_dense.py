import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, torch
from common import flatquad_terminal, c_problem
from approximate_optimal_control_b200 import _capi, pontryagin_utils as pu
prob,_=c_problem("flatquad")
n=1<<20
y=torch.as_tensor(flatquad_terminal(n),dtype=torch.float32,device="cuda").contiguous()
for dbg in (0,1,2,3):
    opts=_capi.make_opts(dtype=_capi.F32, save_dense=1, flags=dbg<<8)
    out={}
    ts=[]
    for i in range(4):
        a=torch.cuda.Event(enable_timing=True); b=torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); a.record(); pu.solve_batch_soa(prob,opts,y,out); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    print("dbg",dbg,"(1=no padding, 2=no node flush)", min(ts))
