import sys, torch, numpy as np
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import ntt_aie_b200 as nt
Q=469762049
def t(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e=[torch.cuda.Event(enable_timing=True) for _ in range(reps+1)]
    e[0].record()
    for k in range(reps):
        fn(); e[k+1].record()
    torch.cuda.synchronize()
    ms=sorted(e[k].elapsed_time(e[k+1]) for k in range(reps))
    return sum(ms[1:-1])/(len(ms)-2)
for logn,batch in ((16,4096),(14,16384),(13,32768),(18,1024)):
    n=1<<logn
    roots=nt.make_roots(n,Q,3)
    x=torch.randint(0,Q,(batch,n),dtype=torch.int32,device='cuda')
    y=torch.empty_like(x)
    with nt.Plan(logn,Q,roots) as p:
        full=t(lambda: p.gs(x,y,batch))
        col=t(lambda: p.gs_stage_range(x,y,batch,12,logn))
        gb=batch*n*8/1e9
        print(f"logn {logn} batch {batch}: full {full:.3f} ms, column pass(es) {col:.3f} ms ({gb/col:.0f} GB/s), tile pass ~{full-col:.3f} ms ({gb/(full-col):.0f} GB/s)")
