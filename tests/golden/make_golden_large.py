"""Digests of LARGE transforms from the reference's OWN golden code (oracle/_ref, compiled
from /root/reference/src/test.cpp:15-60): one N = 2^22, 2^24, 2^26 and 2^27 vector each
(q = 469762049, table w^i as make_roots builds it), and the reference's a[i] = i input at
the reference's larger sizes where a[i] >= p (N = 4096, 8192 with p = 3329).

    python tests/golden/make_golden_large.py        (build container only; ~1-2 minutes)

The outputs are far too large to commit (256 MiB at 2^26), so the fixture holds the
position-weighted digests of tools/digest.py plus the first/last 16 words.
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from tools.digest import digest_numpy  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
Q29 = 469762049
SEED = 0x5EED0026


def large_input(logn: int) -> np.ndarray:
    return np.random.default_rng(SEED + logn).integers(0, Q29, 1 << logn, dtype=np.int32)


def main() -> None:
    assert oracle.have_ref(), "build oracle/_ref first (make -C oracle)"
    out = {"q": Q29, "g": 3, "seed": SEED}
    for logn in (() if "--unreduced-only" in sys.argv else (22, 24, 26, 27)):
        n = 1 << logn
        t0 = time.perf_counter()
        roots = oracle.make_roots(n, Q29, 3)          # 64-bit restatement of make_roots
        a = large_input(logn)
        o = oracle.ref_ntt(a, roots, Q29, -1)          # the reference's own ntt()
        d = digest_numpy(o)
        out[f"digest_{logn}"] = np.array(d, dtype=np.uint64)
        out[f"head_{logn}"] = o[:16].copy()
        out[f"tail_{logn}"] = o[-16:].copy()
        out[f"in_digest_{logn}"] = np.array(digest_numpy(a), dtype=np.uint64)
        print(f"logn {logn}: {time.perf_counter() - t0:.1f} s", flush=True)
    if "--unreduced-only" not in sys.argv:
        np.savez_compressed(os.path.join(HERE, "large_digests.npz"), **out)

    # the reference's own harness input a[i] = i (src/test.cpp:141) at sizes where i >= p:
    # the golden reduces with % on first touch
    un = {}
    for n in (4096, 8192):
        roots = oracle.ref_make_roots(n, 3329, 3)      # verbatim make_roots (all ones: (p-1)/n = 0)
        a = np.arange(n, dtype=np.int32)
        un[f"roots_{n}"] = roots
        un[f"out_{n}"] = oracle.ref_ntt(a, roots, 3329, -1)
    # and with a table that is not all ones, 29-bit q, inputs up to 2^31-1 would overflow the
    # golden's int32 sums, so stay below 2^30 - q (the golden's v0 + v1 must fit int32)
    n = 4096
    roots = oracle.make_roots(n, Q29, 3)
    a = np.random.default_rng(77).integers(0, (1 << 30) - 1, n, dtype=np.int32)
    a[1::2] = np.minimum(a[1::2], Q29 - 1)             # keep v0 + p - v1 >= 0 at stage 0
    un["roots_q29"] = roots
    un["a_q29"] = a
    un["out_q29"] = oracle.ref_ntt(a, roots, Q29, -1)
    np.savez_compressed(os.path.join(HERE, "unreduced_inputs.npz"), **un)
    print("written")


if __name__ == "__main__":
    main()
