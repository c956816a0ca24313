"""ctypes binding of include/nttb200.h.  Device buffers are passed as raw pointers
(``tensor.data_ptr()`` of int32 CUDA tensors, or any integer address); torch is used
by callers for memory and streams only -- nothing here imports it."""
from __future__ import annotations

import ctypes
import os
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "lib", "libnttb200.so")

ORDER_GOLDEN = 0
ORDER_AIE_DEVICE = 1
FORCE_GENERIC = 2
REDUCE_INPUT = 4
INPUT_BITREV = 8
OUTPUT_BITREV = 16
GEN_POWERS = 0
GEN_BITREV = 1

_lib = None
_i32p = ctypes.POINTER(ctypes.c_int32)


class NttError(RuntimeError):
    def __init__(self, status: int, where: str):
        lib = load_library()
        msg = lib.nttb200_strerror(status).decode()
        detail = lib.nttb200_last_error().decode()
        super().__init__(f"{where}: {msg}" + (f" [{detail}]" if detail else ""))
        self.status = status


def lib_path() -> str:
    return _LIB_PATH


def load_library():
    """Load libnttb200.so; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise ImportError(
            f"{_LIB_PATH} is missing: build it with `python ntt-aie_b200/build.py` "
            "(there is no CPU fallback)")
    lib = ctypes.CDLL(_LIB_PATH)
    vp, sz, i32, u32, i64 = (ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int32, ctypes.c_uint32,
                             ctypes.c_int64)
    sig = {
        "nttb200_make_roots": (ctypes.c_int, [i32, _i32p, i32, i32]),
        "nttb200_make_bitrev_table": (ctypes.c_int, [i32, _i32p, i32, i32]),
        "nttb200_powmod": (i32, [i32, i64, i32]),
        "nttb200_plan_create": (ctypes.c_int, [ctypes.POINTER(vp), ctypes.c_int, u32, u32, _i32p,
                                               u32]),
        "nttb200_plan_destroy": (ctypes.c_int, [vp]),
        "nttb200_plan_create_generated": (ctypes.c_int, [ctypes.POINTER(vp), ctypes.c_int, u32, u32, u32,
                                                         u32, u32, u32, u32]),
        "nttb200_plan_table": (ctypes.c_int, [vp, _i32p]),
        "nttb200_reduce": (ctypes.c_int, [vp, vp, vp, sz, vp]),
        "nttb200_bitrev_permute": (ctypes.c_int, [vp, vp, vp, sz, vp]),
        "nttb200_transpose": (ctypes.c_int, [vp, vp, vp, sz, ctypes.c_int, vp]),
        "nttb200_gs_batch": (ctypes.c_int, [vp, vp, vp, sz, ctypes.c_int, vp]),
        "nttb200_ct_batch": (ctypes.c_int, [vp, vp, vp, sz, ctypes.c_int, vp]),
        "nttb200_gs_stage_range": (ctypes.c_int, [vp, vp, vp, sz, ctypes.c_int, ctypes.c_int, vp]),
        "nttb200_gs_host": (ctypes.c_int, [vp, vp, vp, sz, ctypes.c_int]),
        "nttb200_gs_stage_range_scatter": (ctypes.c_int, [vp, vp, ctypes.c_int, ctypes.c_int,
                                                          ctypes.POINTER(vp), ctypes.c_int,
                                                          ctypes.c_int, vp]),
        "nttb200_host_alloc": (ctypes.c_int, [ctypes.POINTER(vp), sz, ctypes.c_int]),
        "nttb200_host_free": (ctypes.c_int, [vp]),
        "nttb200_pointwise": (ctypes.c_int, [vp, vp, vp, vp, sz, vp]),
        "nttb200_scale": (ctypes.c_int, [vp, vp, vp, sz, i32, vp]),
        "nttb200_polymul_negacyclic": (ctypes.c_int, [vp, vp, vp, vp, vp, sz, vp]),
        "nttb200_rns_plan_create": (ctypes.c_int, [ctypes.POINTER(vp), ctypes.c_int, u32,
                                                   ctypes.POINTER(u32), ctypes.POINTER(_i32p), u32]),
        "nttb200_rns_plan_destroy": (ctypes.c_int, [vp]),
        "nttb200_rns_gs_batch": (ctypes.c_int, [vp, vp, vp, sz, vp]),
        "nttb200_rns_ct_batch": (ctypes.c_int, [vp, vp, vp, sz, vp]),
        "nttb200_rns_polymul_negacyclic": (ctypes.c_int, [vp, vp, vp, vp, vp, sz, vp]),
        "nttb200_strerror": (ctypes.c_char_p, [ctypes.c_int]),
        "nttb200_last_error": (ctypes.c_char_p, []),
        "nttb200_kernel_launches": (ctypes.c_uint64, []),
        "nttb200_plan_last_path": (ctypes.c_char_p, [vp]),
        "nttb200_plan_logn": (u32, [vp]),
        "nttb200_plan_modulus": (u32, [vp]),
        "nttb200_version": (ctypes.c_char_p, []),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


EXPORTED_SYMBOLS = (
    "nttb200_make_roots", "nttb200_make_bitrev_table", "nttb200_powmod", "nttb200_plan_create",
    "nttb200_plan_destroy", "nttb200_plan_create_generated", "nttb200_plan_table", "nttb200_reduce",
    "nttb200_bitrev_permute", "nttb200_transpose",
    "nttb200_gs_batch", "nttb200_ct_batch", "nttb200_gs_stage_range",
    "nttb200_gs_host", "nttb200_gs_stage_range_scatter", "nttb200_host_alloc", "nttb200_host_free",
    "nttb200_pointwise", "nttb200_scale", "nttb200_polymul_negacyclic",
    "nttb200_rns_plan_create", "nttb200_rns_plan_destroy", "nttb200_rns_gs_batch",
    "nttb200_rns_ct_batch", "nttb200_rns_polymul_negacyclic",
    "nttb200_strerror", "nttb200_last_error", "nttb200_kernel_launches", "nttb200_plan_last_path",
    "nttb200_plan_logn", "nttb200_plan_modulus", "nttb200_version",
)


def _check(status: int, where: str) -> None:
    if status != 0:
        raise NttError(status, where)


def version() -> str:
    return load_library().nttb200_version().decode()


def kernel_launches() -> int:
    return int(load_library().nttb200_kernel_launches())


# ------------------------------------------------------------------ host tables
def make_roots(n: int, p: int, g: int) -> np.ndarray:
    """``root[0] = 1; make_roots(n, root, p, g)`` of the reference host
    (src/test.cpp:27-32,137-139), 64-bit safe for every p <= 2^30."""
    roots = np.zeros(n, dtype=np.int32)
    _check(load_library().nttb200_make_roots(n, roots.ctypes.data_as(_i32p), p, g), "make_roots")
    return roots


def make_bitrev_table(n: int, p: int, base: int) -> np.ndarray:
    table = np.zeros(n, dtype=np.int32)
    _check(load_library().nttb200_make_bitrev_table(n, table.ctypes.data_as(_i32p), p, base),
           "make_bitrev_table")
    return table


def powmod(b: int, e: int, m: int) -> int:
    return int(load_library().nttb200_powmod(b, e, m))


def negacyclic_tables(n: int, q: int, g: int):
    """(forward psi^bitrev table, inverse psi^-bitrev table) for x^n + 1 mod q;
    q must be a prime with 2n | q-1 and g a primitive root."""
    if (q - 1) % (2 * n):
        raise ValueError("2n must divide q-1")
    psi = powmod(g, (q - 1) // (2 * n), q)
    psi_inv = powmod(psi, q - 2, q)
    return make_bitrev_table(n, q, psi), make_bitrev_table(n, q, psi_inv)


class HostBuffer:
    """Page-locked host memory from the library (successor of the reference's host-only
    buffer objects, src/test.cpp:115-134).  `.array` is a numpy int32 view."""

    def __init__(self, words: int, write_combined: bool = False):
        self._lib = load_library()
        p = ctypes.c_void_p()
        _check(self._lib.nttb200_host_alloc(ctypes.byref(p), words * 4, 1 if write_combined else 0),
               "host_alloc")
        self.ptr = int(p.value)
        self.words = words
        self.array = np.ctypeslib.as_array((ctypes.c_int32 * words).from_address(self.ptr))

    def free(self) -> None:
        if getattr(self, "ptr", 0):
            self.array = None
            self._lib.nttb200_host_free(ctypes.c_void_p(self.ptr))
            self.ptr = 0

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


# ------------------------------------------------------------------------- plan
def _addr(x) -> int:
    """Raw address of a device/host buffer: torch tensor, numpy array or int."""
    if isinstance(x, int):
        return x
    if hasattr(x, "data_ptr"):
        return int(x.data_ptr())
    if isinstance(x, np.ndarray):
        return int(x.ctypes.data)
    raise TypeError(f"cannot take the address of {type(x)!r}")


def _stream(stream) -> Optional[int]:
    if stream is None:
        return None
    if isinstance(stream, int):
        return stream
    return int(stream.cuda_stream)  # torch.cuda.Stream


class Plan:
    """One (N, q, twiddle table) on one GPU -- the successor of the compile-time
    constants + bo_root of the reference (src/aie2.py:14-19, src/test.cpp:119-144)."""

    def __init__(self, logn: int, q: int, table: np.ndarray, device: int = 0, flags: int = 0):
        table = np.ascontiguousarray(table, dtype=np.int32)
        if table.size != (1 << logn):
            raise ValueError("table must hold N = 2^logn words")
        self._lib = load_library()
        handle = ctypes.c_void_p()
        _check(self._lib.nttb200_plan_create(ctypes.byref(handle), device, logn, q,
                                             table.ctypes.data_as(_i32p), flags), "plan_create")
        self._h = handle
        self.logn, self.n, self.q, self.device, self.flags = logn, 1 << logn, q, device, flags

    @classmethod
    def generated(cls, logn: int, q: int, kind: int, base: int, gen_logn: Optional[int] = None,
                  block_mult: int = 1, device: int = 0, flags: int = 0) -> "Plan":
        """Plan whose table is generated ON THE DEVICE (nothing is shipped):
        table[h+i] = gen(h*block_mult + i), gen(e) = base^e (GEN_POWERS) or
        base^bitrev(e) over gen_logn bits (GEN_BITREV).  See nttb200_plan_create_generated."""
        self = cls.__new__(cls)
        self._lib = load_library()
        handle = ctypes.c_void_p()
        _check(self._lib.nttb200_plan_create_generated(
            ctypes.byref(handle), device, logn, q, kind, base,
            logn if gen_logn is None else gen_logn, block_mult, flags), "plan_create_generated")
        self._h = handle
        self.logn, self.n, self.q, self.device, self.flags = logn, 1 << logn, q, device, flags
        return self

    def close(self) -> None:
        if getattr(self, "_h", None):
            self._lib.nttb200_plan_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    @property
    def last_path(self) -> str:
        return self._lib.nttb200_plan_last_path(self._h).decode()

    def table(self) -> np.ndarray:
        """The plan's table as the reference host would hold it in bo_root."""
        out = np.empty(self.n, dtype=np.int32)
        _check(self._lib.nttb200_plan_table(self._h, out.ctypes.data_as(_i32p)), "plan_table")
        return out

    def bitrev_permute(self, d_in, d_out, batch: int, stream=None) -> None:
        """out[b][i] = in[b][bitrev(i)] (in place allowed)."""
        _check(self._lib.nttb200_bitrev_permute(self._h, _addr(d_in), _addr(d_out), batch,
                                                _stream(stream)), "bitrev_permute")

    def transpose(self, d_in, d_out, batch: int, to_batch_minor: bool, stream=None) -> None:
        """[batch][N] -> [N][batch] (to_batch_minor) or back; out of place."""
        _check(self._lib.nttb200_transpose(self._h, _addr(d_in), _addr(d_out), batch,
                                           1 if to_batch_minor else 0, _stream(stream)), "transpose")

    def reduce(self, d_in, d_out, count: int, stream=None) -> None:
        """out = in mod q for arbitrary int32 words (the golden's `%` on first touch)."""
        _check(self._lib.nttb200_reduce(self._h, _addr(d_in), _addr(d_out), count,
                                        _stream(stream)), "reduce")

    def gs(self, d_in, d_out, batch: int, stage: int = -1, stream=None) -> None:
        """Golden network ``ntt(a, n, roots, p, stage)`` (src/test.cpp:34-60) on device."""
        _check(self._lib.nttb200_gs_batch(self._h, _addr(d_in), _addr(d_out), batch, stage,
                                          _stream(stream)), "gs_batch")

    def ct(self, d_in, d_out, batch: int, stage: int = -1, stream=None) -> None:
        _check(self._lib.nttb200_ct_batch(self._h, _addr(d_in), _addr(d_out), batch, stage,
                                          _stream(stream)), "ct_batch")

    def gs_stage_range(self, d_in, d_out, batch: int, stage_begin: int, stage_end: int,
                       stream=None) -> None:
        _check(self._lib.nttb200_gs_stage_range(self._h, _addr(d_in), _addr(d_out), batch,
                                                stage_begin, stage_end, _stream(stream)),
               "gs_stage_range")

    def gs_stage_range_scatter(self, d_buf, stage_begin: int, stage_end: int, peer_ptrs, rank: int,
                               stream=None) -> None:
        """Stages [stage_begin, logn) of one vector with the results stored into the peers'
        receive buffers (the all-to-all fused into the last pass)."""
        world = len(peer_ptrs)
        arr = (ctypes.c_void_p * world)(*[int(x) for x in peer_ptrs])
        _check(self._lib.nttb200_gs_stage_range_scatter(self._h, _addr(d_buf), stage_begin,
                                                        stage_end, arr, world, rank,
                                                        _stream(stream)), "gs_stage_range_scatter")

    def gs_host(self, h_in, h_out, batch: int, stage: int = -1) -> None:
        """Host buffers in, host buffers out (the reference's BO sync + launch + sync,
        src/test.cpp:148-168)."""
        _check(self._lib.nttb200_gs_host(self._h, _addr(h_in), _addr(h_out), batch, stage),
               "gs_host")

    def pointwise(self, d_a, d_b, d_c, count: int, stream=None) -> None:
        _check(self._lib.nttb200_pointwise(self._h, _addr(d_a), _addr(d_b), _addr(d_c), count,
                                           _stream(stream)), "pointwise")

    def scale(self, d_a, d_c, count: int, scalar: int, stream=None) -> None:
        _check(self._lib.nttb200_scale(self._h, _addr(d_a), _addr(d_c), count, scalar,
                                       _stream(stream)), "scale")


class RnsPlan:
    """N = 4096 polynomials in RNS form: L channels with their own primes and tables,
    data laid out [batch][L][4096]; one launch per transform serves all channels."""

    def __init__(self, qs, tables, device: int = 0):
        self._lib = load_library()
        self.limbs = len(qs)
        tabs = [np.ascontiguousarray(t, dtype=np.int32) for t in tables]
        if len(tabs) != self.limbs or any(t.size != 4096 for t in tabs):
            raise ValueError("one 4096-word table per modulus")
        qarr = (ctypes.c_uint32 * self.limbs)(*[int(q) for q in qs])
        parr = (_i32p * self.limbs)(*[t.ctypes.data_as(_i32p) for t in tabs])
        handle = ctypes.c_void_p()
        _check(self._lib.nttb200_rns_plan_create(ctypes.byref(handle), device, self.limbs, qarr, parr,
                                                 0), "rns_plan_create")
        self._h = handle
        self.qs = [int(q) for q in qs]

    def close(self) -> None:
        if getattr(self, "_h", None):
            self._lib.nttb200_rns_plan_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def gs(self, d_in, d_out, batch: int, stream=None) -> None:
        _check(self._lib.nttb200_rns_gs_batch(self._h, _addr(d_in), _addr(d_out), batch,
                                              _stream(stream)), "rns_gs_batch")

    def ct(self, d_in, d_out, batch: int, stream=None) -> None:
        _check(self._lib.nttb200_rns_ct_batch(self._h, _addr(d_in), _addr(d_out), batch,
                                              _stream(stream)), "rns_ct_batch")


def rns_polymul_negacyclic(fwd: RnsPlan, inv: RnsPlan, d_a, d_b, d_c, batch: int, stream=None) -> None:
    _check(fwd._lib.nttb200_rns_polymul_negacyclic(fwd._h, inv._h, _addr(d_a), _addr(d_b),
                                                   _addr(d_c), batch, _stream(stream)),
           "rns_polymul_negacyclic")


def polymul_negacyclic(fwd: Plan, inv: Plan, d_a, d_b, d_c, batch: int, stream=None) -> None:
    _check(fwd._lib.nttb200_polymul_negacyclic(fwd._h, inv._h, _addr(d_a), _addr(d_b), _addr(d_c),
                                               batch, _stream(stream)), "polymul_negacyclic")


def ntt(a: np.ndarray, n: int, roots: np.ndarray, p: int, stage: int = -1, device: int = 0,
        flags: int = 0) -> np.ndarray:
    """Drop-in for the golden call ``ntt(a, n, roots_rev, p, stage)`` (src/test.cpp:34):
    host arrays in, transformed copy out, computed on the GPU through
    ``nttb200_gs_host``.  ``a`` may be 1-D (one polynomial) or 2-D (a batch).  Like the
    golden, inputs need not be reduced (``a[i] = i`` with n > p is fine): they are taken
    mod p on first touch."""
    a = np.ascontiguousarray(a, dtype=np.int32)
    if a.shape[-1] != n or n & (n - 1) or n < 2:
        raise ValueError("last dimension must be n, a power of two >= 2")
    out = np.empty_like(a)
    batch = a.size // n
    with Plan(n.bit_length() - 1, p, roots, device=device, flags=flags) as plan:
        plan.gs_host(a, out, batch, stage)
    return out
