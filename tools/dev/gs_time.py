"""Device time of nttb200_gs_batch / ct_batch per N at 2^26 (or given) coefficients."""
import json, os, sys, statistics
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import ntt_aie_b200 as nt
Q = 469762049
peak = 6539.5
logns = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "12").split(",")]
kind = sys.argv[2] if len(sys.argv) > 2 else "gs"
logtotal = int(sys.argv[3]) if len(sys.argv) > 3 else 28
for logn in logns:
    n = 1 << logn
    batch = (1 << logtotal) // n
    roots = nt.make_roots(n, Q, 3)
    a = torch.randint(0, Q, (batch, n), dtype=torch.int32, device="cuda")
    c = torch.empty_like(a)
    with nt.Plan(logn, Q, roots) as p:
        fn = (lambda: p.gs(a, c, batch)) if kind == "gs" else (lambda: p.ct(a, c, batch))
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(21)]
        ev[0].record()
        for k in range(20):
            fn()
            ev[k + 1].record()
        torch.cuda.synchronize()
        ms = sorted(ev[k].elapsed_time(ev[k + 1]) for k in range(20))
        t = statistics.mean(ms[1:-1])
        print(json.dumps({"kind": kind, "logn": logn, "batch": batch, "ms": round(t, 4),
                          "frac": round(batch * n * 8 / (t * 1e-3) / 1e9 / peak, 4),
                          "path": p.last_path}), flush=True)
