"""Launch each secondary kernel a few times at 2^24 coefficients (for ncu -k captures)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import ntt_aie_b200 as nt
Q = 469762049
for logn in (11, 12, 14, 16):
    n = 1 << logn
    batch = (1 << 24) // n
    fwd, inv = nt.negacyclic_tables(n, Q, 3)
    x = torch.randint(0, Q, (batch, n), dtype=torch.int32, device="cuda")
    y = torch.empty_like(x)
    with nt.Plan(logn, Q, fwd) as pf, nt.Plan(logn, Q, inv) as pi:
        for _ in range(3):
            pi.gs(x, y, batch)
            pf.ct(x, y, batch)
        torch.cuda.synchronize()
print("done")
