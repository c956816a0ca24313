"""Regenerates tests/golden/*.npz from the reference's OWN golden code.

Run in the build container (where /root/reference exists):
    python tests/golden/make_golden.py
It loads oracle/_ref/libntt_ref.so -- modPow/make_roots/ntt compiled by
oracle/Makefile from /root/reference/src/test.cpp:15-60 where they lie -- and
stores small input/output vectors.  The fixtures travel to the GPU box; the
reference tree does not.

Cases
  default      the reference's one test: N=2048, p=3329, g=3, a[i]=i, full depth
               (src/test.cpp:66-67,76-77,137-141,203-207), table from the VERBATIM
               make_roots, plus every partial depth 0..10 as a digest
  n4096_q29    N=4096, q=469762049 (7*2^26+1), table w^i with w=3^((q-1)/N)
               (64-bit make_roots restatement; the verbatim one overflows there),
               seeded random rows + edge rows (zeros, all q-1, a[i]=i)
  n4096_p3329  N=4096, p=3329: (p-1)/n == 0 so the reference's table is all ones
  small        N=2..64 at q=469762049 and p=3329
  n65536_q29   one N=2^16 row (digest + first/last words only)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
Q29 = 469762049


def digest(a: np.ndarray) -> np.uint64:
    return np.uint64(oracle.fnv1a64_words(a))


def main() -> None:
    assert oracle.have_ref(), "build oracle/_ref first (make -C oracle)"
    out = {}

    # ---- default: the reference's own test -------------------------------
    n, p, g = 2048, 3329, 3
    roots = oracle.ref_make_roots(n, p, g)           # verbatim make_roots
    a = np.arange(n, dtype=np.int32)
    full = oracle.ref_ntt(a, roots, p, 10)            # test_stage = n - 1 = 10
    np.savez_compressed(os.path.join(HERE, "default_n2048_p3329.npz"), n=n, p=p, g=g, roots=roots,
                        a=a, out=full,
                        stage_digest=np.array([digest(oracle.ref_ntt(a, roots, p, s))
                                               for s in range(11)], dtype=np.uint64))

    # ---- N=4096, 29-bit q ---------------------------------------------------
    n = 4096
    roots = oracle.make_roots(n, Q29, 3)              # widened restatement (ref overflows)
    rng = np.random.default_rng(0x5EED0001)
    rows = [rng.integers(0, Q29, n, dtype=np.int32) for _ in range(5)]
    rows += [np.zeros(n, np.int32), np.full(n, Q29 - 1, np.int32), np.arange(n, dtype=np.int32)]
    a = np.stack(rows)
    np.savez_compressed(os.path.join(HERE, "n4096_q29.npz"), n=n, p=Q29, g=3, roots=roots, a=a,
                        out=oracle.ref_ntt(a, roots, Q29, -1))

    # ---- N=4096, p=3329 (all-ones table) -------------------------------------
    roots = oracle.ref_make_roots(n, 3329, 3)
    assert (roots == 1).all()
    a = np.stack([np.arange(n, dtype=np.int32) % 3329,
                  np.random.default_rng(7).integers(0, 3329, n, dtype=np.int32)])
    np.savez_compressed(os.path.join(HERE, "n4096_p3329.npz"), n=n, p=3329, g=3, roots=roots, a=a,
                        out=oracle.ref_ntt(a, roots, 3329, -1))

    # ---- small sizes ------------------------------------------------------------
    small = {}
    for logn in range(1, 7):
        n = 1 << logn
        for p, tag in ((Q29, "q29"), (3329, "p3329")):
            roots = oracle.make_roots(n, p, 3)
            if p == 3329:
                assert (roots == oracle.ref_make_roots(n, p, 3)).all()
            a = np.random.default_rng(100 + logn).integers(0, p, (3, n), dtype=np.int32)
            small[f"roots_{tag}_{logn}"] = roots
            small[f"a_{tag}_{logn}"] = a
            small[f"out_{tag}_{logn}"] = oracle.ref_ntt(a, roots, p, -1)
    np.savez_compressed(os.path.join(HERE, "small.npz"), **small)

    # ---- N=2^16 digest -------------------------------------------------------------
    n = 1 << 16
    roots = oracle.make_roots(n, Q29, 3)
    a = np.random.default_rng(0x5EED0016).integers(0, Q29, n, dtype=np.int32)
    o = oracle.ref_ntt(a, roots, Q29, -1)
    np.savez_compressed(os.path.join(HERE, "n65536_q29_digest.npz"), n=n, p=Q29, g=3, seed=0x5EED0016,
                        digest=digest(o), head=o[:16], tail=o[-16:], total=np.int64(o.astype(np.int64).sum()))
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
