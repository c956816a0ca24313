"""CPU tests (gloo, world_size 2 and 4) of the multi-GPU host logic: batch sharding
and the four-step split with its all-to-all transpose.  The local stage work is done
by a CPU engine built on the oracle so the test runs without a GPU; the product's
engine is the CUDA plan (covered by the gpu tests)."""
import importlib.util
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import Q29, ROOT


def _fourstep():
    spec = importlib.util.spec_from_file_location(
        "nttb200_fourstep", os.path.join(ROOT, "ntt-aie_b200", "fourstep.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def gs_stage_range_numpy(a, table, p, sb, se):
    """Stages [sb, se) of the golden network (src/test.cpp:36-59), vectorised."""
    a = a.astype(np.int64).copy()
    n = a.shape[0]
    for s in range(sb, se):
        t, h = 1 << s, n >> (s + 1)
        v = a.reshape(h, 2, t)
        w = table[h:2 * h].astype(np.int64)[:, None]
        x, y = v[:, 0, :].copy(), v[:, 1, :].copy()
        v[:, 0, :] = (x + y) % p
        v[:, 1, :] = ((x + p - y) % p) * w % p
    return a.astype(np.int32)


class OracleEngine:
    def __init__(self, q, t_local, t_cross):
        self.q, self.t_local, self.t_cross = q, t_local, t_cross

    def local_full(self, buf):
        import oracle
        buf.copy_(torch.from_numpy(oracle.ntt_gs(buf.numpy(), self.t_local, self.q)))

    def cross_stages(self, buf, sb, se):
        buf.copy_(torch.from_numpy(gs_stage_range_numpy(buf.numpy(), self.t_cross, self.q, sb, se)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, logn, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import oracle
        fs = _fourstep()
        n = 1 << logn
        rng = np.random.default_rng(42)
        table = rng.integers(0, Q29, n, dtype=np.int32)        # arbitrary table: table-driven
        a = rng.integers(0, Q29, n, dtype=np.int32)
        want = oracle.ntt_gs(a, table, Q29)
        s = n // world
        eng = OracleEngine(Q29, fs.local_table(table, world, rank), fs.cross_table(table, world, s))
        plan = fs.FourStepNTT(logn, Q29, table, rank, world, engine=eng)
        # natural order
        shard = torch.from_numpy(a[rank * s:(rank + 1) * s].copy())
        res = plan.forward(shard, torch.empty_like(shard), natural_order=True).numpy()
        ok_nat = np.array_equal(res, want[rank * s:(rank + 1) * s])
        # transposed "device" order: out_k[r][c] = NTT(a)[r*S + k*S/G + c]
        shard = torch.from_numpy(a[rank * s:(rank + 1) * s].copy())
        res = plan.forward(shard, torch.empty_like(shard), natural_order=False).numpy()
        c = s // world
        exp = np.concatenate([want[r * s + rank * c: r * s + (rank + 1) * c] for r in range(world)])
        ok_dev = np.array_equal(res, exp)
        # batched sharding: no communication on the data path
        batch, nb = 37, 256
        roots = oracle.make_roots(nb, Q29, 3)
        polys = np.random.default_rng(7).integers(0, Q29, (batch, nb), dtype=np.int32)
        b0, b1 = fs.shard_batch(batch, world, rank)
        mine = oracle.ntt_gs(polys[b0:b1], roots, Q29)
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)                   # test-only gather (ragged)
        ok_batch = np.array_equal(np.concatenate(gathered), oracle.ntt_gs(polys, roots, Q29))
        with open(os.path.join(out_dir, f"rank{rank}.txt"), "w") as f:
            f.write(f"{int(ok_nat)} {int(ok_dev)} {int(ok_batch)}")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,logn", [(2, 10), (4, 12), (2, 3)])
def test_fourstep_and_sharding_over_gloo(tmp_path, oracle_mod, world, logn):
    port = _free_port()
    mp.spawn(_worker, args=(world, port, logn, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert open(tmp_path / f"rank{r}.txt").read() == "1 1 1", f"rank {r}"


def test_stage_range_helper_matches_oracle(oracle_mod):
    rng = np.random.default_rng(1)
    n = 256
    table = rng.integers(0, Q29, n, dtype=np.int32)
    a = rng.integers(0, Q29, n, dtype=np.int32)
    for k in range(8):
        assert np.array_equal(gs_stage_range_numpy(a, table, Q29, 0, k + 1),
                              oracle_mod.ntt_gs(a, table, Q29, k))


def test_shard_helpers():
    fs = _fourstep()
    assert [fs.shard_batch(10, 4, r) for r in range(4)] == [(0, 2), (2, 5), (5, 7), (7, 10)]
    assert fs.shard_batch(0, 2, 1) == (0, 0)
    table = np.arange(64, dtype=np.int32)
    t1 = fs.local_table(table, 4, 1)            # S = 16, rank 1: T[h+i] = table[h*(4+1)+i]
    for h in (8, 4, 2, 1):
        assert np.array_equal(t1[h:2 * h], table[5 * h:5 * h + h])
    assert np.array_equal(fs.cross_table(table, 4, 16)[:4], table[:4])
