"""Repeat the persistent N = 2^16 kernels (GS, CT, product) and the staged forward kernels many
times on fresh random data and compare every run with the two-pass / small-batch paths: a rare
ordering bug (counters, TMA after generic stores) would show up as a mismatch."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import ntt_aie_b200 as nt
Q = 469762049
n = 1 << 16
rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 20
fwd, inv = nt.negacyclic_tables(n, Q, 3)
bad = 0
with nt.Plan(16, Q, fwd) as pf, nt.Plan(16, Q, inv) as pi:
    for r in range(rounds):
        batch = 896 + 8 * (r % 5)
        g = torch.Generator(device="cuda").manual_seed(1000 + r)
        a = torch.randint(0, Q, (batch, n), dtype=torch.int32, device="cuda", generator=g)
        big_ct = torch.empty_like(a); big_gs = torch.empty_like(a)
        pf.ct(a, big_ct, batch); p1 = pf.last_path
        pi.gs(a, big_gs, batch); p2 = pi.last_path
        ref_ct = torch.empty_like(a); ref_gs = torch.empty_like(a)
        for s in range(0, batch, 100):
            e = min(batch, s + 100)
            pf.ct(a[s:e], ref_ct[s:e], e - s)
            pi.gs(a[s:e], ref_gs[s:e], e - s)
        torch.cuda.synchronize()
        ok = bool(torch.equal(big_ct, ref_ct)) and bool(torch.equal(big_gs, ref_gs))
        bad += 0 if ok else 1
        if not ok or r == 0:
            print(r, batch, p1, p2, pf.last_path, ok, flush=True)
# staged forward kernels at N = 2048, 4096, 8192: batch vs one-by-one chunks
for logn in (11, 12, 13, 15):
    nn = 1 << logn
    f2, _ = nt.negacyclic_tables(nn, Q, 3)
    with nt.Plan(logn, Q, f2) as p:
        for r in range(rounds):
            batch = (1 << 22) // nn + 3 * r
            g = torch.Generator(device="cuda").manual_seed(5000 + r)
            a = torch.randint(0, Q, (batch, nn), dtype=torch.int32, device="cuda", generator=g)
            o1 = torch.empty_like(a); o2 = torch.empty_like(a)
            p.ct(a, o1, batch)
            half = batch // 2
            p.ct(a[:half], o2[:half], half); p.ct(a[half:], o2[half:], batch - half)
            torch.cuda.synchronize()
            ok = bool(torch.equal(o1, o2))
            bad += 0 if ok else 1
            if not ok:
                print("ct mismatch", logn, r, batch, flush=True)
print("stress done, mismatches:", bad)
