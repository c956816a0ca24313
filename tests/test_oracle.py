"""CPU tests: pin the oracle (oracle/ntt_oracle.c) to the reference's own golden
(oracle/_ref, compiled from /root/reference/src/test.cpp:15-60) and to the committed
fixtures generated from it."""
import numpy as np
import pytest

from conftest import Q29, Q30, load_golden


def test_default_config_known_answer(oracle_mod):
    """The reference's one test (src/test.cpp:66-67,76-77,137-141,203-207)."""
    g = load_golden("default_n2048_p3329.npz")
    n, p = int(g["n"]), int(g["p"])
    roots = oracle_mod.make_roots(n, p, int(g["g"]))
    assert np.array_equal(roots, g["roots"])          # verbatim make_roots output
    assert roots[1024] == 341 and roots[2047] == 3251  # SURVEY 4 known answers
    out = oracle_mod.ntt_gs(g["a"], roots, p, 10)
    assert np.array_equal(out, g["out"])
    assert out[:8].tolist() == [2187, 1952, 747, 1368, 1399, 3021, 3063, 854]
    assert out[1024:1028].tolist() == [177, 1607, 1745, 3137]
    assert out[2044:].tolist() == [1712, 1428, 3043, 1667]
    assert int(out.sum()) == 3333380


def test_stage_early_exit_matches_fixture(oracle_mod):
    """`stage` early exit (src/test.cpp:55-58) for every depth of the default config."""
    g = load_golden("default_n2048_p3329.npz")
    for s in range(11):
        out = oracle_mod.ntt_gs(g["a"], g["roots"], int(g["p"]), s)
        assert np.uint64(oracle_mod.fnv1a64_words(out)) == g["stage_digest"][s]
    # stage = -1 and stage >= logn-1 never hit the early return: full depth
    full = oracle_mod.ntt_gs(g["a"], g["roots"], int(g["p"]), -1)
    assert np.array_equal(full, g["out"])


@pytest.mark.parametrize("name", ["n4096_q29.npz", "n4096_p3329.npz"])
def test_n4096_fixtures(oracle_mod, name):
    g = load_golden(name)
    roots = oracle_mod.make_roots(int(g["n"]), int(g["p"]), int(g["g"]))
    assert np.array_equal(roots, g["roots"])
    assert np.array_equal(oracle_mod.ntt_gs(g["a"], roots, int(g["p"])), g["out"])


def test_small_fixtures(oracle_mod):
    g = load_golden("small.npz")
    for logn in range(1, 7):
        for p, tag in ((Q29, "q29"), (3329, "p3329")):
            roots = oracle_mod.make_roots(1 << logn, p, 3)
            assert np.array_equal(roots, g[f"roots_{tag}_{logn}"])
            assert np.array_equal(oracle_mod.ntt_gs(g[f"a_{tag}_{logn}"], roots, p),
                                  g[f"out_{tag}_{logn}"])


def test_n65536_digest(oracle_mod):
    g = load_golden("n65536_q29_digest.npz")
    n, p = int(g["n"]), int(g["p"])
    a = np.random.default_rng(int(g["seed"])).integers(0, p, n, dtype=np.int32)
    o = oracle_mod.ntt_gs(a, oracle_mod.make_roots(n, p, 3), p)
    assert np.uint64(oracle_mod.fnv1a64_words(o)) == g["digest"]
    assert np.array_equal(o[:16], g["head"]) and np.array_equal(o[-16:], g["tail"])


def test_against_reference_library(oracle_mod):
    """Live comparison with the reference's compiled golden where it is available."""
    if not oracle_mod.have_ref():
        pytest.skip("oracle/_ref/libntt_ref.so not built (no /root/reference here)")
    rng = np.random.default_rng(1)
    # verbatim modPow / make_roots on their valid domain (p < 46341)
    for p, g in ((3329, 3), (12289, 11), (40961, 3)):
        for n in (8, 256, 2048, 4096):
            assert np.array_equal(oracle_mod.make_roots(n, p, g), oracle_mod.ref_make_roots(n, p, g))
        for _ in range(50):
            x, e = int(rng.integers(0, p)), int(rng.integers(0, 1 << 15))
            assert oracle_mod.modpow(x, e, p) == oracle_mod.ref_modpow(x, e, p)
    # verbatim ntt() on 12/29/30-bit moduli, all depths
    for p in (3329, Q29, Q30, 1 << 30):
        for logn in (1, 2, 5, 9, 12):
            n = 1 << logn
            roots = rng.integers(0, p, n, dtype=np.int32)   # arbitrary table: table-driven
            a = rng.integers(0, p, (2, n), dtype=np.int32)
            for stage in (-1, 0, logn // 2, logn - 1, logn + 3):
                assert np.array_equal(oracle_mod.ntt_gs(a, roots, p, stage),
                                      oracle_mod.ref_ntt(a, roots, p, stage))


def test_threaded_batch_equals_serial(oracle_mod):
    rng = np.random.default_rng(3)
    n, p = 512, Q29
    roots = oracle_mod.make_roots(n, p, 3)
    a = rng.integers(0, p, (37, n), dtype=np.int32)
    want = oracle_mod.ntt_gs(a, roots, p)
    got = a.copy()
    oracle_mod.ntt_gs_batch_inplace(got, roots, p, nthreads=4)
    assert np.array_equal(got, want)
    if oracle_mod.have_ref():
        got = a.copy()
        oracle_mod.ref_ntt_batch_inplace(got, roots, p, nthreads=3)
        assert np.array_equal(got, want)


def test_ct_is_inverse_partner_of_golden(oracle_mod):
    """GS(psi^-bitrev) o CT(psi^bitrev) = n * identity -- validates the CT restatement
    against the golden network (parity unpinned upstream)."""
    rng = np.random.default_rng(4)
    for logn in (1, 3, 8, 11):
        n = 1 << logn
        psi = oracle_mod.powmod(3, (Q29 - 1) // (2 * n), Q29)
        tf = oracle_mod.make_bitrev_table(n, Q29, psi)
        ti = oracle_mod.make_bitrev_table(n, Q29, oracle_mod.powmod(psi, Q29 - 2, Q29))
        a = rng.integers(0, Q29, (2, n), dtype=np.int32)
        back = oracle_mod.ntt_gs(oracle_mod.ntt_ct(a, tf, Q29), ti, Q29)
        assert np.array_equal(back, oracle_mod.scale(a, n % Q29, Q29))


def test_negacyclic_product_vs_schoolbook(oracle_mod):
    rng = np.random.default_rng(5)
    for n in (2, 16, 256, 1024):
        psi = oracle_mod.powmod(3, (Q29 - 1) // (2 * n), Q29)
        tf = oracle_mod.make_bitrev_table(n, Q29, psi)
        ti = oracle_mod.make_bitrev_table(n, Q29, oracle_mod.powmod(psi, Q29 - 2, Q29))
        a = rng.integers(0, Q29, n, dtype=np.int32)
        b = rng.integers(0, Q29, n, dtype=np.int32)
        prod = oracle_mod.pointwise(oracle_mod.ntt_ct(a, tf, Q29), oracle_mod.ntt_ct(b, tf, Q29), Q29)
        c = oracle_mod.scale(oracle_mod.ntt_gs(prod, ti, Q29), oracle_mod.powmod(n, Q29 - 2, Q29), Q29)
        assert np.array_equal(c, oracle_mod.negacyclic_schoolbook(a, b, Q29))


def test_ans_order_is_pairwise_bit_swap(oracle_mod):
    """src/test.cpp:69-71,212-219: block i of the golden lands in block ans_order[i]."""
    n = 2048
    golden = np.arange(n, dtype=np.int32)
    got = oracle_mod.ans_order_permute(golden)
    order = [0, 2, 1, 3, 8, 10, 9, 11, 4, 6, 5, 7, 12, 14, 13, 15]
    blk = n // 16
    for i, o in enumerate(order):
        assert np.array_equal(got[o * blk:(o + 1) * blk], golden[i * blk:(i + 1) * blk])
        assert o == ((i & 5) << 1 | (i & 10) >> 1)


def test_device_barrett_agrees_with_golden_mod(oracle_mod):
    """barrett_2k (src/aie_core.cc:27-39) with w, u from src/aie2.py:18-19 equals `%`."""
    import math
    rng = np.random.default_rng(6)
    for p in (3329, Q29):
        w = math.ceil(math.log2(p))
        u = (1 << (2 * w)) // p
        for _ in range(2000):
            a, b = int(rng.integers(0, p)), int(rng.integers(0, p))
            assert oracle_mod.barrett_2k(a, b, p, w, u) == a * b % p
            assert oracle_mod.modadd(a, b, p) == (a + b) % p
            assert oracle_mod.modsub(a, b, p) == (a - b) % p


def test_position_weighted_digest_numpy_equals_torch():
    """tools/digest.py: the numpy form (fixtures) and the torch form (GPU tests, bench.py)
    agree, and the digest is additive over shards with their global offsets."""
    import torch
    from tools.digest import as_unsigned, digest_numpy, digest_torch
    rng = np.random.default_rng(5)
    x = rng.integers(0, 469762049, 100003, dtype=np.int32)
    want = digest_numpy(x)
    assert as_unsigned(digest_torch(torch.from_numpy(x))) == want
    parts = torch.zeros(2, dtype=torch.int64)
    for lo, hi in ((0, 17), (17, 50000), (50000, 100003)):
        parts += digest_torch(torch.from_numpy(x[lo:hi]), offset=lo)
    assert as_unsigned(parts) == want
    y = x.copy()
    y[[3, 99999]] = y[[99999, 3]]            # a swap changes it
    assert digest_numpy(y) != want


def test_large_digest_fixture_is_consistent():
    """tests/golden/large_digests.npz was written from the reference's own golden output;
    the N = 2^22 entry is cheap enough to recompute here with the restatement."""
    import os
    import oracle
    from conftest import GOLDEN, Q29
    from tools.digest import digest_numpy
    g = np.load(os.path.join(GOLDEN, "large_digests.npz"))
    logn = 22
    a = np.random.default_rng(int(g["seed"]) + logn).integers(0, Q29, 1 << logn, dtype=np.int32)
    out = oracle.ntt_gs(a, oracle.make_roots(1 << logn, Q29, 3), Q29)
    assert digest_numpy(out) == tuple(int(v) for v in g[f"digest_{logn}"])
    assert np.array_equal(out[:16], g[f"head_{logn}"]) and np.array_equal(out[-16:], g[f"tail_{logn}"])


def test_unreduced_fixture_matches_restatement():
    """a[i] = i with n > p (the reference harness input at its larger sizes): the golden's
    `%` on first touch == reduce, then transform."""
    import os
    import oracle
    from conftest import GOLDEN
    g = np.load(os.path.join(GOLDEN, "unreduced_inputs.npz"))
    for n in (4096, 8192):
        a = np.arange(n, dtype=np.int32) % 3329
        assert np.array_equal(oracle.ntt_gs(a, g[f"roots_{n}"], 3329), g[f"out_{n}"])
    a = (g["a_q29"].astype(np.int64) % 469762049).astype(np.int32)
    assert np.array_equal(oracle.ntt_gs(a, g["roots_q29"], 469762049), g["out_q29"])


def test_wide_modulus_restatement(oracle_mod):
    """2^30 < p < 2^31 is outside the golden's int32 domain (src/test.cpp:48-49 overflow):
    the widened restatement (a) equals the int32 form wherever that is defined, (b) equals
    Python big-integer arithmetic on the wide range, (c) GS(inv) o CT(fwd) = n * identity and
    the negacyclic product equals the schoolbook product at p = 2013265921 = 15*2^27 + 1."""
    rng = np.random.default_rng(77)
    for p in (3329, 469762049, 1 << 30):
        for n in (2, 64, 1024):
            table = rng.integers(0, p, n, dtype=np.int32)
            a = rng.integers(0, p, (2, n), dtype=np.int32)
            assert np.array_equal(oracle_mod.ntt_gs_wide(a, table, p), oracle_mod.ntt_gs(a, table, p))
            assert np.array_equal(oracle_mod.ntt_ct_wide(a, table, p), oracle_mod.ntt_ct(a, table, p))

    def gs_python(a, table, p):
        a = [int(x) for x in a]
        n, t, m = len(a), 1, len(a)
        while m > 1:
            h = m // 2
            for i in range(h):
                for j in range(2 * i * t, 2 * i * t + t):
                    v0, v1 = a[j], a[j + t]
                    a[j] = (v0 + v1) % p
                    a[j + t] = ((v0 + p - v1) % p) * int(table[h + i]) % p
            t, m = t * 2, h
        return np.array(a, dtype=np.int64)

    for p in (2013265921, 2147483647, (1 << 30) + 3):
        n = 256
        table = rng.integers(0, p, n, dtype=np.int64).astype(np.int32)
        a = rng.integers(0, p, n, dtype=np.int64).astype(np.int32)
        a[:4] = p - 1
        assert np.array_equal(oracle_mod.ntt_gs(a, table, p).astype(np.int64), gs_python(a, table, p))
    p, n, g = 2013265921, 512, 31
    psi = pow(g, (p - 1) // (2 * n), p)
    assert pow(psi, n, p) == p - 1
    fwd = oracle_mod.make_bitrev_table(n, p, psi)
    inv = oracle_mod.make_bitrev_table(n, p, pow(psi, p - 2, p))
    a = rng.integers(0, p, (2, n), dtype=np.int64).astype(np.int32)
    b = rng.integers(0, p, (2, n), dtype=np.int64).astype(np.int32)
    assert np.array_equal(oracle_mod.ntt_gs(oracle_mod.ntt_ct(a, fwd, p), inv, p),
                          oracle_mod.scale(a, n, p))
    prod = oracle_mod.pointwise(oracle_mod.ntt_ct(a, fwd, p), oracle_mod.ntt_ct(b, fwd, p), p)
    c = oracle_mod.scale(oracle_mod.ntt_gs(prod, inv, p), pow(n, p - 2, p), p)
    assert np.array_equal(c[0], oracle_mod.negacyclic_schoolbook(a[0], b[0], p))
