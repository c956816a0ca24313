"""CPU oracle for the ntt-aie golden model -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package; the product package
(``ntt-aie_b200``) never does.

Two ctypes-loaded libraries:

* ``liboracle.so``  -- the plain-C restatement in ``oracle/ntt_oracle.c`` (each
  function cites the reference lines it follows), built by ``oracle/Makefile``;
* ``_ref/libntt_ref.so`` -- the reference's OWN golden ``modPow`` / ``make_roots`` /
  ``ntt`` (``/root/reference/src/test.cpp:15-60``) compiled from where it lies.  It
  exists wherever ``/root/reference`` was present at build time; the prebuilt file
  travels to the GPU box.  ``have_ref()`` says whether it is loadable.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PORT_PATH = os.path.join(_HERE, "liboracle.so")
_REF_PATH = os.path.join(_HERE, "_ref", "libntt_ref.so")

_i32p = ctypes.POINTER(ctypes.c_int32)
_port = None
_ref = None


def build() -> None:
    """Compile the checker libraries (``make -C oracle``)."""
    subprocess.run(["make", "-C", _HERE, "--no-print-directory"], check=True,
                   stdout=subprocess.DEVNULL)


def _ptr(a: np.ndarray):
    assert a.dtype == np.int32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_i32p)


def _load_port():
    global _port
    if _port is None:
        if not os.path.exists(_PORT_PATH):
            build()
        lib = ctypes.CDLL(_PORT_PATH)
        lib.oracle_modpow.restype = ctypes.c_int32
        lib.oracle_modpow.argtypes = [ctypes.c_int32] * 3
        lib.oracle_powmod.restype = ctypes.c_int32
        lib.oracle_powmod.argtypes = [ctypes.c_int32, ctypes.c_int64, ctypes.c_int32]
        lib.oracle_make_roots.argtypes = [ctypes.c_int32, _i32p, ctypes.c_int32, ctypes.c_int32]
        for name in ("oracle_ntt_gs", "oracle_ntt_ct", "oracle_ntt_gs_wide", "oracle_ntt_ct_wide"):
            getattr(lib, name).argtypes = [_i32p, ctypes.c_int32, _i32p, ctypes.c_int32,
                                           ctypes.c_int32]
        lib.oracle_pointwise.argtypes = [_i32p, _i32p, _i32p, ctypes.c_int64, ctypes.c_int32]
        lib.oracle_scale.argtypes = [_i32p, _i32p, ctypes.c_int64, ctypes.c_int32, ctypes.c_int32]
        lib.oracle_negacyclic_schoolbook.argtypes = [_i32p, _i32p, _i32p, ctypes.c_int32,
                                                     ctypes.c_int32]
        lib.oracle_ans_order_permute.argtypes = [_i32p, _i32p, ctypes.c_int32]
        for name in ("oracle_modadd", "oracle_modsub"):
            getattr(lib, name).restype = ctypes.c_int32
            getattr(lib, name).argtypes = [ctypes.c_int32] * 3
        lib.oracle_barrett_2k.restype = ctypes.c_int32
        lib.oracle_barrett_2k.argtypes = [ctypes.c_int32] * 5
        lib.oracle_make_bitrev_table.argtypes = [ctypes.c_int32, _i32p, ctypes.c_int32,
                                                 ctypes.c_int32]
        lib.oracle_ntt_gs_batch.argtypes = [_i32p, ctypes.c_int32, ctypes.c_int64, _i32p,
                                            ctypes.c_int32, ctypes.c_int32, ctypes.c_int32]
        _port = lib
    return _port


def have_ref() -> bool:
    return os.path.exists(_REF_PATH)


def _load_ref():
    global _ref
    if _ref is None:
        lib = ctypes.CDLL(_REF_PATH)
        lib.ref_modpow.restype = ctypes.c_int32
        lib.ref_modpow.argtypes = [ctypes.c_int32] * 3
        lib.ref_make_roots.argtypes = [ctypes.c_int32, _i32p, ctypes.c_int32, ctypes.c_int32]
        lib.ref_ntt.argtypes = [_i32p, ctypes.c_int32, _i32p, ctypes.c_int32, ctypes.c_int32]
        lib.ref_ntt_batch.argtypes = [_i32p, ctypes.c_int32, ctypes.c_int64, _i32p, ctypes.c_int32,
                                      ctypes.c_int32, ctypes.c_int32]
        _ref = lib
    return _ref


# --------------------------------------------------------------------------- port
def modpow(x: int, n: int, mod: int) -> int:
    return int(_load_port().oracle_modpow(x, n, mod))


def powmod(b: int, e: int, m: int) -> int:
    return int(_load_port().oracle_powmod(b, e, m))


def make_roots(n: int, p: int, g: int) -> np.ndarray:
    """roots[0]=1; make_roots(n, roots, p, g)  (src/test.cpp:137-139)."""
    roots = np.zeros(n, dtype=np.int32)
    roots[0] = 1
    _load_port().oracle_make_roots(n, _ptr(roots), p, g)
    return roots


def ntt_gs(a: np.ndarray, roots: np.ndarray, p: int, stage: int = -1) -> np.ndarray:
    """Golden GS network (src/test.cpp:34-60) on a copy; 2-D input = batch of rows."""
    a = np.ascontiguousarray(a, dtype=np.int32).copy()
    roots = np.ascontiguousarray(roots, dtype=np.int32)
    n = a.shape[-1]
    flat = a.reshape(-1, n)
    lib = _load_port()
    # p > 2^30 is outside the golden's int32 domain: widened restatement (parity unpinned)
    fn = lib.oracle_ntt_gs_wide if p > (1 << 30) else lib.oracle_ntt_gs
    for row in flat:
        fn(_ptr(row), n, _ptr(roots), p, stage)
    return a


def ntt_gs_wide(a: np.ndarray, roots: np.ndarray, p: int, stage: int = -1) -> np.ndarray:
    """The widened (64-bit sums) form of the golden network for ANY p < 2^31."""
    a = np.ascontiguousarray(a, dtype=np.int32).copy()
    roots = np.ascontiguousarray(roots, dtype=np.int32)
    n = a.shape[-1]
    lib = _load_port()
    for row in a.reshape(-1, n):
        lib.oracle_ntt_gs_wide(_ptr(row), n, _ptr(roots), p, stage)
    return a


def ntt_ct_wide(a: np.ndarray, table: np.ndarray, p: int, stage: int = -1) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.int32).copy()
    table = np.ascontiguousarray(table, dtype=np.int32)
    n = a.shape[-1]
    lib = _load_port()
    for row in a.reshape(-1, n):
        lib.oracle_ntt_ct_wide(_ptr(row), n, _ptr(table), p, stage)
    return a


def ntt_ct(a: np.ndarray, table: np.ndarray, p: int, stage: int = -1) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.int32).copy()
    table = np.ascontiguousarray(table, dtype=np.int32)
    n = a.shape[-1]
    lib = _load_port()
    fn = lib.oracle_ntt_ct_wide if p > (1 << 30) else lib.oracle_ntt_ct
    for row in a.reshape(-1, n):
        fn(_ptr(row), n, _ptr(table), p, stage)
    return a


def pointwise(a: np.ndarray, b: np.ndarray, p: int) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.int32)
    b = np.ascontiguousarray(b, dtype=np.int32)
    c = np.empty_like(a)
    _load_port().oracle_pointwise(_ptr(a), _ptr(b), _ptr(c), a.size, p)
    return c


def scale(a: np.ndarray, s: int, p: int) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.int32)
    c = np.empty_like(a)
    _load_port().oracle_scale(_ptr(a), _ptr(c), a.size, s, p)
    return c


def negacyclic_schoolbook(a: np.ndarray, b: np.ndarray, p: int) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.int32)
    b = np.ascontiguousarray(b, dtype=np.int32)
    c = np.empty_like(a)
    _load_port().oracle_negacyclic_schoolbook(_ptr(a), _ptr(b), _ptr(c), a.size, p)
    return c


def ans_order_permute(golden: np.ndarray) -> np.ndarray:
    """answers[ans_order[i]*B + j] = golden[i*B + j]  (src/test.cpp:212-219)."""
    golden = np.ascontiguousarray(golden, dtype=np.int32)
    out = np.empty_like(golden)
    n = golden.shape[-1]
    lib = _load_port()
    for src, dst in zip(golden.reshape(-1, n), out.reshape(-1, n)):
        lib.oracle_ans_order_permute(_ptr(src), _ptr(dst), n)
    return out


def modadd(a: int, b: int, q: int) -> int:
    return int(_load_port().oracle_modadd(a, b, q))


def modsub(a: int, b: int, q: int) -> int:
    return int(_load_port().oracle_modsub(a, b, q))


def barrett_2k(a: int, b: int, q: int, w: int, u: int) -> int:
    return int(_load_port().oracle_barrett_2k(a, b, q, w, u))


def make_bitrev_table(n: int, p: int, base: int) -> np.ndarray:
    """table[k] = base^bitrev_logn(k) in the golden's table[h+i] index rule."""
    t = np.zeros(n, dtype=np.int32)
    _load_port().oracle_make_bitrev_table(n, _ptr(t), p, base)
    return t


def ntt_gs_batch_inplace(a: np.ndarray, roots: np.ndarray, p: int, nthreads: int,
                         stage: int = -1) -> None:
    """Threaded batched golden, in place (the "port" CPU baseline)."""
    assert a.dtype == np.int32 and a.flags["C_CONTIGUOUS"] and a.ndim == 2
    roots = np.ascontiguousarray(roots, dtype=np.int32)
    _load_port().oracle_ntt_gs_batch(_ptr(a), a.shape[1], a.shape[0], _ptr(roots), p, stage,
                                     nthreads)


# ---------------------------------------------------------------------- reference
def ref_modpow(x: int, n: int, mod: int) -> int:
    return int(_load_ref().ref_modpow(x, n, mod))


def ref_make_roots(n: int, p: int, g: int) -> np.ndarray:
    """The reference's verbatim make_roots (valid for p <= 65536 only)."""
    roots = np.zeros(n, dtype=np.int32)
    _load_ref().ref_make_roots(n, _ptr(roots), p, g)
    return roots


def ref_ntt(a: np.ndarray, roots: np.ndarray, p: int, stage: int = -1) -> np.ndarray:
    """The reference's verbatim ntt() on a copy (valid for p <= 2^30)."""
    a = np.ascontiguousarray(a, dtype=np.int32).copy()
    roots = np.ascontiguousarray(roots, dtype=np.int32)
    n = a.shape[-1]
    lib = _load_ref()
    for row in a.reshape(-1, n):
        lib.ref_ntt(_ptr(row), n, _ptr(roots), p, stage)
    return a


def ref_ntt_batch_inplace(a: np.ndarray, roots: np.ndarray, p: int, nthreads: int,
                          stage: int = -1) -> None:
    """Threaded batched verbatim golden, in place (the "reference" CPU baseline)."""
    assert a.dtype == np.int32 and a.flags["C_CONTIGUOUS"] and a.ndim == 2
    roots = np.ascontiguousarray(roots, dtype=np.int32)
    _load_ref().ref_ntt_batch(_ptr(a), a.shape[1], a.shape[0], _ptr(roots), p, stage, nthreads)


def fnv1a64_words(words: np.ndarray) -> int:
    """FNV-1a-64 over int32 words (xor word, multiply) -- the digest SURVEY 4 quotes."""
    h = 0xCBF29CE484222325
    for w in np.ascontiguousarray(words, dtype=np.int32).astype(np.uint32).tolist():
        h ^= w
        h = (h * 0x100000001B3) & 0xFFFFFFFFFFFFFFFF
    return h
