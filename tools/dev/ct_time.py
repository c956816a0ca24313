import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import ntt_aie_b200 as nt
Q=469762049
def t(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e=[torch.cuda.Event(enable_timing=True) for _ in range(reps+1)]
    e[0].record()
    for k in range(reps):
        fn(); e[k+1].record()
    torch.cuda.synchronize()
    ms=sorted(e[k].elapsed_time(e[k+1]) for k in range(reps))
    return sum(ms[1:-1])/(len(ms)-2)
for logn,batch in ((12,16384),(12,65536),(11,32768),(13,8192),(16,1024)):
    n=1<<logn
    fwd,inv=nt.negacyclic_tables(n,Q,3)
    x=torch.randint(0,Q,(batch,n),dtype=torch.int32,device='cuda'); y=torch.empty_like(x)
    with nt.Plan(logn,Q,fwd) as pf, nt.Plan(logn,Q,inv) as pi:
        g=t(lambda: pi.gs(x,y,batch)); c=t(lambda: pf.ct(x,y,batch))
        gb=batch*n*8/1e6
        print(f"logn {logn} batch {batch}: GS {g:.4f} ms ({gb/g/6539.5:.3f} of roofline) [{pi.last_path}]  CT {c:.4f} ms ({gb/c/6539.5:.3f}) [{pf.last_path}]")
