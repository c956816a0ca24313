"""CPU tests of the C-ABI library: it loads, exports every symbol include/nttb200.h
declares, its host-side table builders match the oracle, and it fails LOUDLY (no CPU
fallback) when no GPU is present."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import Q29, ROOT


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "nttb200.h")).read()
    return sorted(set(re.findall(r"NTTB200_API[^;(]*?\b(nttb200_\w+)\s*\(", text)))


def test_header_symbols_are_exported(lib):
    names = _declared_symbols()
    assert len(names) >= 19
    cdll = ctypes.CDLL(lib.lib_path())
    for name in names:
        assert hasattr(cdll, name), f"{name} declared in nttb200.h but not exported"
    assert set(lib.EXPORTED_SYMBOLS) == set(names)


def test_only_abi_symbols_are_visible(lib):
    import subprocess
    out = subprocess.run(["nm", "-D", "--defined-only", lib.lib_path()], capture_output=True,
                         text=True, check=True).stdout
    ours = [l.split()[-1] for l in out.splitlines() if " T " in l and "nttb200" in l]
    assert set(ours) == set(_declared_symbols())


def test_host_tables_match_oracle(lib, oracle_mod):
    for n, p, g in ((2048, 3329, 3), (4096, 3329, 3), (4096, Q29, 3), (1 << 16, Q29, 3), (2, 17, 3)):
        assert np.array_equal(lib.make_roots(n, p, g), oracle_mod.make_roots(n, p, g))
    for n in (2, 64, 4096):
        assert np.array_equal(lib.make_bitrev_table(n, Q29, 12345),
                              oracle_mod.make_bitrev_table(n, Q29, 12345))
    assert lib.powmod(3, 1 << 40, Q29) == pow(3, 1 << 40, Q29)
    fwd, inv = lib.negacyclic_tables(1024, Q29, 3)
    assert all((int(a) * int(b)) % Q29 == 1 for a, b in zip(fwd[1:50], inv[1:50]))


def test_argument_validation_without_device(lib):
    L = lib.load_library()
    roots = lib.make_roots(16, 17, 3)
    h = ctypes.c_void_p()
    ptr = roots.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))
    assert L.nttb200_plan_create(ctypes.byref(h), 0, 0, 17, ptr, 0) == 1      # logn out of range
    assert L.nttb200_plan_create(ctypes.byref(h), 0, 4, 1 << 31, ptr, 0) == 2  # modulus >= 2^31
    assert L.nttb200_plan_create(ctypes.byref(h), 0, 4, 1, ptr, 0) == 2
    assert L.nttb200_plan_create(ctypes.byref(h), 0, 4, 5, ptr, 0) == 3       # entry >= q
    assert L.nttb200_plan_create(ctypes.byref(h), 0, 3, 17, ptr, 1) == 1      # AIE order needs N>=16
    assert L.nttb200_gs_batch(None, None, None, 1, -1, None) == 1
    assert L.nttb200_plan_destroy(None) == 0
    assert b"CPU" in L.nttb200_strerror(5)


def test_no_cpu_fallback(lib):
    """Without a CUDA device plan creation must fail, not compute on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present; the failure path is exercised on the CPU box")
    with pytest.raises(lib.NttError) as err:
        lib.Plan(11, 3329, lib.make_roots(2048, 3329, 3))
    assert err.value.status in (4, 5)
    with pytest.raises(lib.NttError):
        lib.ntt(np.arange(2048, dtype=np.int32), 2048, lib.make_roots(2048, 3329, 3), 3329, 10)


def test_product_does_not_reference_oracle():
    """The product tree must not import, link or name the oracle."""
    pkg = os.path.join(ROOT, "ntt-aie_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                if f == "build.py":   # builds the tests/host harness, which links the oracle
                    continue
                assert "oracle" not in text.lower(), f"{f} mentions the oracle"
