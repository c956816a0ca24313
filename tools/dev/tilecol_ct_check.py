"""Parity of the persistent forward kernel (N = 2^16) against the oracle rows and the two-pass path."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import ntt_aie_b200 as nt, oracle
n = 1 << 16
rng = np.random.default_rng(4242)
for q, batch in ((469762049, 1024), ((1 << 30) - 35, 900), ((1 << 29) - 3, 896)):
    table = rng.integers(0, q, n, dtype=np.int32)
    table[1::2] = q - 1
    a = rng.integers(0, q, (batch, n), dtype=np.int32)
    a[0] = q - 1
    d_a = torch.from_numpy(a).cuda()
    d_o = torch.empty_like(d_a)
    with nt.Plan(16, q, table) as p:
        p.ct(d_a, d_o, batch)
        torch.cuda.synchronize()
        path = p.last_path
        rows = [0, 1, 2, 511, batch - 1]
        got = d_o.cpu().numpy()
        want = oracle.ntt_ct(a[rows], table, q)
        ok_rows = np.array_equal(got[rows], want)
        # in place
        p.ct(d_a, d_a, batch)
        torch.cuda.synchronize()
        ok_inplace = bool(torch.equal(d_a, d_o))
        # a small batch takes the two-pass path: compare the whole big batch against it in chunks
        ok_all = True
        d_b = torch.from_numpy(a).cuda()
        d_c = torch.empty_like(d_b)
        for s in range(0, batch, 128):
            e = min(batch, s + 128)
            p.ct(d_b[s:e], d_c[s:e], e - s)
        torch.cuda.synchronize()
        ok_all = bool(torch.equal(d_c, d_o))
        print(q, batch, path, "rows", ok_rows, "inplace", ok_inplace, "vs two-pass", ok_all, "small path", p.last_path, flush=True)
