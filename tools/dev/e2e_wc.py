import os, sys, time, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import ntt_aie_b200 as nt
Q=469762049; N=4096; B=65536
plan=nt.Plan(12,Q,nt.make_roots(N,Q,3))
d_in=torch.randint(0,Q,(B,N),dtype=torch.int32,device='cuda')
for wc in (False, True):
    hin=nt.HostBuffer(B*N, write_combined=wc); hout=nt.HostBuffer(B*N)
    torch.from_numpy(hin.array).view(B,N).copy_(d_in)
    plan.gs_host(hin.ptr, hout.ptr, B)
    t0=time.perf_counter()
    for _ in range(5): plan.gs_host(hin.ptr, hout.ptr, B)
    dt=(time.perf_counter()-t0)/5
    print("write_combined" if wc else "plain pinned ", f"{B/dt:.4g} polys/s  {B*N*4/dt/1e9:.1f} GB/s each way")
    hin.free(); hout.free()
h_in = torch.empty((B, N), dtype=torch.int32, pin_memory=True); h_out = torch.empty_like(h_in).pin_memory()
h_in.copy_(d_in); plan.gs_host(h_in, h_out, B)
t0=time.perf_counter()
for _ in range(5): plan.gs_host(h_in, h_out, B)
dt=(time.perf_counter()-t0)/5
print("torch pinned  ", f"{B/dt:.4g} polys/s  {B*N*4/dt/1e9:.1f} GB/s each way")
