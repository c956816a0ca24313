"""Device time of the N=4096 golden kernel with plan flags (bitrev adapters etc.)."""
import json, os, sys, statistics
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import ntt_aie_b200 as nt
Q, n, batch = 469762049, 4096, 65536
roots = nt.make_roots(n, Q, 3)
a = torch.randint(0, Q, (batch, n), dtype=torch.int32, device="cuda")
c = torch.empty_like(a)
for flags in (0, nt.INPUT_BITREV, nt.OUTPUT_BITREV, nt.INPUT_BITREV | nt.OUTPUT_BITREV, nt.ORDER_AIE_DEVICE):
    with nt.Plan(12, Q, roots, flags=flags) as p:
        for _ in range(3):
            p.gs(a, c, batch)
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(21)]
        ev[0].record()
        for k in range(20):
            p.gs(a, c, batch)
            ev[k + 1].record()
        torch.cuda.synchronize()
        ms = sorted(ev[k].elapsed_time(ev[k + 1]) for k in range(20))
        t = statistics.mean(ms[1:-1])
        print(json.dumps({"flags": flags, "ms": round(t, 4), "frac": round(batch * n * 8 / (t * 1e-3) / 1e9 / 6539.5, 4),
                          "path": p.last_path}), flush=True)
