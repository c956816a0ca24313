"""One large transform split across G GPUs: local stages, ONE all-to-all, cross-GPU stages.

The reference splits a single transform the same way one level down: each AIE tile
runs the stages whose stride fits its contiguous slice (ntt_stage0_to_Nminus5,
src/aie_core.cc:189-361), then the remaining log2(#tiles) stages pair whole slices
between tiles with one broadcast twiddle each (ntt_1stage, src/aie_core.cc:161-187,
scheduled in src/aie2.py:178-295).  Here the "tiles" are GPUs and the neighbour-memory
exchange is an NCCL all-to-all over NVLink (torch.distributed.all_to_all_single) --
the only collective on the path.

    N = G * S, rank r owns the contiguous shard a[r*S, (r+1)*S).

    step 1  stages 0 .. log2(S)-1 are local.  They are an ordinary length-S golden
            transform with the DERIVED table
                T_r[h + i] = table[h*(G + r) + i],   h = S/2, S/4, .., 1,  i < h
            (global block index of local block i at a stage with h local blocks is
            r*h + i, and the stage has G*h blocks in total).
    step 2  all-to-all: rank r sends its k-th S/G slice to rank k.  Rank k then holds
            rows[r][c] = a[r*S + k*S/G + c].
    step 3  stages log2(S) .. log2(N)-1 pair rows r and r + 2^m; the twiddle is
            table[(G >> (m+1)) + (r >> (m+1))] for every column.  On the local buffer
            viewed as one length-S vector these are stages log2(S/G) .. log2(S)-1 of a
            plan whose table starts with table[0..G).
    step 4  (optional) a second all-to-all returns the result to the golden's natural
            order; without it the output stays in the transposed order
            out_k[r][c] = NTT(a)[r*S + k*S/G + c]   (documented like ans_order).

Because the per-stage twiddles are taken from the caller's table by the golden's own
index rule (src/test.cpp:45), the result is bit-exact against the golden ntt() for ANY
table, not only DFT-valid ones.

Fused exchange (`fused=True`, CUDA engine only).  NCCL's all-to-all costs a launch and a
protocol round trip that dominate at 8 GPUs (28 MiB per rank).  With the receive
buffers in symmetric memory (torch.distributed._symmetric_memory: every rank maps every
peer's buffer) the LAST local pass stores each result straight into the peer that owns
it after the transpose (`nttb200_gs_stage_range_scatter`), so the transpose rides on the
pass's own NVLink stores and overlaps its math; the natural-order return trip is fused
into the cross-GPU pass the same way.  Only two tiny signal barriers remain.

Nothing here computes on the host: the local work goes through an *engine* -- the
CUDA plans of this package by default.  Tests inject a CPU engine to exercise the
sharding logic over gloo.
"""
from __future__ import annotations

from typing import Optional

import numpy as np


def shard_batch(batch: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous [begin, end) of `batch` independent polynomials owned by `rank`
    (batched configs shard with no communication)."""
    return batch * rank // world, batch * (rank + 1) // world


def local_table(table: np.ndarray, world: int, rank: int) -> np.ndarray:
    """T_r of step 1 from the global table (length N, golden index rule)."""
    n = table.shape[0]
    s = n // world
    out = np.zeros(s, dtype=np.int32)
    h = s // 2
    while h >= 1:
        lo = h * (world + rank)
        out[h:2 * h] = table[lo:lo + h]
        h //= 2
    return out


def cross_table(table: np.ndarray, world: int, shard_len: int) -> np.ndarray:
    """Table of step 3: a length-S table whose first G entries are table[0..G)."""
    out = np.zeros(shard_len, dtype=np.int32)
    out[:world] = table[:world]
    return out


class CudaEngine:
    """Local stage executor backed by the CUDA plans (the product path)."""

    def __init__(self, logs: int, q: int, t_local: np.ndarray, t_cross: np.ndarray, device: int):
        from . import api
        self.plan_local = api.Plan(logs, q, t_local, device=device)
        self.plan_cross = api.Plan(logs, q, t_cross, device=device)

    @classmethod
    def generated(cls, logn: int, logs: int, q: int, kind: int, base: int, world: int, rank: int,
                  device: int) -> "CudaEngine":
        """Both per-rank tables generated on the device from (kind, base): rank r's local
        table is table[h*(G+r)+i], the cross table table[0..G) -- nothing is built on or
        shipped from the host (nttb200_plan_create_generated)."""
        from . import api
        self = cls.__new__(cls)
        self.plan_local = api.Plan.generated(logs, q, kind, base, gen_logn=logn,
                                             block_mult=world + rank, device=device)
        self.plan_cross = api.Plan.generated(logs, q, kind, base, gen_logn=logn, block_mult=1,
                                             device=device)
        return self

    def local_full(self, buf) -> None:
        self.plan_local.gs(buf, buf, 1)

    def cross_stages(self, buf, stage_begin: int, stage_end: int) -> None:
        self.plan_cross.gs_stage_range(buf, buf, 1, stage_begin, stage_end)

    def close(self) -> None:
        self.plan_local.close()
        self.plan_cross.close()


class SymmetricBuffers:
    """Receive buffers every rank can store into (NVLink peer mappings)."""

    def __init__(self, shard_len: int, device: int, group):
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        grp = group if group is not None else dist.group.WORLD
        dev = torch.device("cuda", device)
        # TWO sets, alternated per call: a peer that has already entered call k+1 scatters
        # into the other set while this rank may still read its result of call k; it cannot
        # be two calls ahead because every call contains barriers all ranks take part in.
        self.scratch, self.final, self.scratch_ptrs, self.final_ptrs = [], [], [], []
        self.handles = []
        for _ in range(2):
            sc = symm_mem.empty(shard_len, dtype=torch.int32, device=dev)
            fi = symm_mem.empty(shard_len, dtype=torch.int32, device=dev)
            hs, hf = symm_mem.rendezvous(sc, grp), symm_mem.rendezvous(fi, grp)
            self.scratch.append(sc)
            self.final.append(fi)
            self.handles += [hs, hf]
            self.scratch_ptrs.append([int(p) for p in hs.buffer_ptrs])
            self.final_ptrs.append([int(p) for p in hf.buffer_ptrs])
        self.calls = 0

    def barrier(self) -> None:
        self.handles[0].barrier(channel=0)


class FourStepNTT:
    """Golden GS network of length N = 2^logn over `world` ranks (power of two)."""

    def __init__(self, logn: int, q: int, table: Optional[np.ndarray], rank: int, world: int,
                 device: Optional[int] = None, engine=None, group=None, fused: bool = False,
                 generated: Optional[tuple] = None):
        """`table`: the caller's length-N table (any residues), or None together with
        `generated=(kind, base)` to have every rank generate its own tables on its GPU
        (GEN_POWERS with base = g^((q-1)/N) is the reference's make_roots table)."""
        if world & (world - 1) or world < 1:
            raise ValueError("world size must be a power of two")
        self.logn, self.q, self.rank, self.world, self.group = logn, q, rank, world, group
        self.n = 1 << logn
        self.shard = self.n // world
        self.logs = logn - (world.bit_length() - 1)
        if self.shard < world:
            raise ValueError("shard must hold at least one element per peer")
        dev = 0 if device is None else device
        if generated is not None:
            if engine is not None or table is not None:
                raise ValueError("generated tables need the CUDA engine and no host table")
            kind, base = generated
            self.t_local = self.t_cross = None
            self.engine = CudaEngine.generated(logn, self.logs, q, kind, base, world, rank, dev)
        else:
            table = np.ascontiguousarray(table, dtype=np.int32)
            if table.shape[0] != self.n:
                raise ValueError("table must hold N words")
            self.t_local = local_table(table, world, rank)
            self.t_cross = cross_table(table, world, self.shard)
            self.engine = engine if engine is not None else CudaEngine(
                self.logs, q, self.t_local, self.t_cross, dev)
        self.symm = None
        if fused and world > 1:
            if engine is not None:
                raise ValueError("the fused exchange needs the CUDA engine")
            if self.logs < 13:
                raise ValueError("fused exchange needs shards of at least 2^13 coefficients")
            self.symm = SymmetricBuffers(self.shard, dev, group)

    def close(self) -> None:
        if hasattr(self.engine, "close"):
            self.engine.close()

    def forward(self, shard, scratch, natural_order: bool = True):
        """`shard`: this rank's S contiguous coefficients (torch int32 tensor on the
        engine's device); `scratch`: same shape.  Returns the tensor holding the result
        (one of the two buffers)."""
        import torch.distributed as dist
        eng, world = self.engine, self.world
        if self.symm is not None:
            return self._forward_fused(shard, natural_order)
        eng.local_full(shard)                                   # step 1
        if world == 1:
            return shard
        dist.all_to_all_single(scratch, shard, group=self.group)  # step 2 (the transpose)
        logc = self.logs - (world.bit_length() - 1)
        eng.cross_stages(scratch, logc, self.logs)               # step 3
        if not natural_order:
            return scratch
        dist.all_to_all_single(shard, scratch, group=self.group)  # step 4
        return shard

    def _forward_fused(self, shard, natural_order: bool):
        """Steps 1-4 with both transposes fused into the passes' stores.  Returns the
        symmetric buffer that holds the result; it stays valid until the call AFTER the next
        one (the receive buffers are double-buffered, see SymmetricBuffers), so a rank may
        still read its result while a faster peer has already started the next transform."""
        eng, sy = self.engine, self.symm
        logc = self.logs - (self.world.bit_length() - 1)
        k = sy.calls & 1
        sy.calls += 1
        scratch, final = sy.scratch[k], sy.final[k]
        # steps 1+2: local stages; the last pass scatters into every rank's scratch
        eng.plan_local.gs_stage_range_scatter(shard, 0, self.logs, sy.scratch_ptrs[k], self.rank)
        sy.barrier()                      # all slices have landed everywhere
        if not natural_order:
            eng.plan_cross.gs_stage_range(scratch, scratch, 1, logc, self.logs)   # step 3
            sy.barrier()                  # keeps the ranks within one call of each other
            return scratch
        # steps 3+4: cross-GPU stages; their results go straight home
        eng.plan_cross.gs_stage_range_scatter(scratch, logc, self.logs, sy.final_ptrs[k], self.rank)
        sy.barrier()
        return final
