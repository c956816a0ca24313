"""Launch each kernel family a few times (for `ncu -k regex:<name>` captures):
    python tools/dev/ncu_targets.py <which>
which: headline | polymul | poly15 | tilecol16 | ct4096 | ct15 | ct16 | fourstep_local"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import ntt_aie_b200 as nt
Q = 469762049
which = sys.argv[1]


def run(logn, total_log, fn_name):
    n = 1 << logn
    batch = (1 << total_log) // n
    fwd, inv = nt.negacyclic_tables(n, Q, 3)
    x = torch.randint(0, Q, (batch, n), dtype=torch.int32, device="cuda")
    y = torch.empty_like(x)
    z = torch.empty_like(x)
    with nt.Plan(logn, Q, fwd) as pf, nt.Plan(logn, Q, inv) as pi:
        for _ in range(5):
            if fn_name == "gs":
                pi.gs(x, y, batch)
            elif fn_name == "ct":
                pf.ct(x, y, batch)
            else:
                nt.polymul_negacyclic(pf, pi, x, y, z, batch)
        torch.cuda.synchronize()
        print(which, pi.last_path, pf.last_path)


if which == "headline":
    run(12, 28, "gs")
elif which == "polymul":
    run(12, 26, "mul")
elif which == "poly15":
    run(15, 26, "gs")
elif which == "tilecol16":
    run(16, 28, "gs")
elif which == "ct4096":
    run(12, 28, "ct")
elif which == "ct15":
    run(15, 26, "ct")
elif which == "ct16":
    run(16, 28, "ct")
elif which == "fourstep_local":
    run(23, 23, "gs")
