import torch, sys
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ark_b200 import ops
for (N,V) in [(4966,60943),(10333,24101),(20000,60943)]:
    ldv=(V+7)//8*8
    x=torch.randn(N,ldv,device="cuda").to(torch.bfloat16)
    tgt=torch.randint(1,V,(N,),device="cuda",dtype=torch.int32)
    loss=torch.zeros(1,device="cuda")
    for _ in range(3): ops.softmax_ce(x.clone(),V,tgt,1e-4,True,loss,None)
    xs=[x.clone() for _ in range(5)]
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for w in xs: ops.softmax_ce(w,V,tgt,1e-4,True,loss,None)
    e1.record(); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)/5
    print(N,V,f"{ms:.3f} ms  {2*N*V*2/ms/1e6:.0f} GB/s")
