"""Multi-GPU host logic of the NIQKI path: one process per GPU, index sharded by genome id.

SURVEY.md §8e: the path shards naturally.  Rank r owns the contiguous gid block
``[r*ceil(N/G), (r+1)*ceil(N/G))`` — it sketches those genomes and builds its own CSR, no
communication.  Queries: every rank sketches its slice of the query entries, ONE all-gather of the
query sketches (torch.distributed: NCCL over NVLink on GPUs, gloo in the CPU tests), every rank
counts all queries against its shard; per-shard hit lists are disjoint in gid, so the merge on
rank 0 is concatenate + sort by (count, gid) descending — exactly the order
``Index::query_sketch`` produces (/root/reference/src/niqki_index.cpp:685).  Thresholding is per
genome, so it commutes with sharding.

On GPUs the exchange is the library's own: ``Comm`` wraps ``nq_comm_init_rank`` /
``nq_allgather_sketches`` / ``nq_bcast_sketches`` (NCCL behind the C ABI, u16 on the wire) and
``merge_hits`` wraps ``nq_hits_merge``; ``torch.distributed`` only carries the 128-byte NCCL id and
the small per-rank hit lists to rank 0.  ``ShardedIndex`` keeps the same logic over a generic
``dist`` object so that it also runs on CPU (gloo) in the tests.

The compute object (``engine``) is anything with the ``niqki_b200.Index`` methods used here
(``compute_sketches`` / ``sketch_many``, ``insert_sketches``, ``query_sketches``); the product
passes a ``niqki_b200.Index`` (CUDA), the CPU tests pass an adapter.
"""
from __future__ import annotations

import numpy as np


def shard_range(n_total: int, world: int, rank: int):
    """Contiguous block of ids owned by ``rank``: gpu = gid // ceil(N/G)."""
    per = -(-n_total // world) if n_total else 0
    lo = min(n_total, rank * per)
    return lo, min(n_total, lo + per)


def owner_of(gid: int, n_total: int, world: int) -> int:
    per = -(-n_total // world)
    return gid // per


def merge_hit_lists(parts, nq: int):
    """``parts`` = per-shard ``(ptr u64[nq+1], counts u32[], gids u32[])``.  Returns the merged
    triple with every query's hits sorted by (count, gid) descending (std::greater<pair>, :685)."""
    ptr = np.zeros(nq + 1, np.uint64)
    keys = []
    for q in range(nq):
        segs = []
        for p, c, g in parts:
            a, b = int(p[q]), int(p[q + 1])
            if b > a:
                segs.append((c[a:b].astype(np.uint64) << np.uint64(32)) | g[a:b].astype(np.uint64))
        k = np.sort(np.concatenate(segs))[::-1] if segs else np.zeros(0, np.uint64)
        keys.append(k)
        ptr[q + 1] = ptr[q] + np.uint64(k.size)
    allk = np.concatenate(keys) if keys else np.zeros(0, np.uint64)
    return ptr, (allk >> np.uint64(32)).astype(np.uint32), (allk & np.uint64(0xFFFFFFFF)).astype(np.uint32)


class ShardedIndex:
    """One rank's view of a gid-sharded index.  ``dist`` is ``torch.distributed`` (initialised by
    the caller) or ``None`` for a single process."""

    def __init__(self, engine, dist=None, device=None):
        self.engine = engine
        self.dist = dist if (dist is not None and dist.is_initialized() and dist.get_world_size() > 1) else None
        self.world = self.dist.get_world_size() if self.dist else 1
        self.rank = self.dist.get_rank() if self.dist else 0
        self.device = device
        self.n_total = 0

    # ---- index phase: no communication
    def index_entries(self, n_total: int, load_entries):
        """``load_entries(lo, hi)`` returns this rank's entries (list of byte arrays) for gids
        [lo, hi).  Builds the shard with global gids."""
        self.n_total = n_total
        lo, hi = shard_range(n_total, self.world, self.rank)
        self.gid_lo, self.gid_hi = lo, hi
        if hi > lo:
            sk, _ = self.engine.sketch_many(load_entries(lo, hi))
            self.engine.insert_sketches(sk, gid_base=lo)
        return lo, hi

    # ---- query phase: one all-gather of sketches, then local counting, then a gather of hits
    def allgather_sketches(self, local):
        """``local`` int32 [n_local][F] (numpy or torch) -> all ranks' sketches in rank order."""
        if not self.dist:
            return local
        import torch

        t = torch.as_tensor(local)
        if self.device is not None:
            t = t.to(self.device)
        F = t.shape[1] if t.ndim == 2 else self.engine.F
        counts = torch.zeros(self.world, dtype=torch.int64, device=t.device)
        mine = torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device)
        self.dist.all_gather_into_tensor(counts, mine)
        cmax = int(counts.max().item())
        padded = torch.full((cmax, F), -1, dtype=torch.int32, device=t.device)
        padded[: t.shape[0]] = t
        out = torch.empty((self.world * cmax, F), dtype=torch.int32, device=t.device)
        self.dist.all_gather_into_tensor(out, padded)
        keep = [out[r * cmax: r * cmax + int(counts[r].item())] for r in range(self.world)]
        return torch.cat(keep, dim=0)

    def query_entries(self, nq_total: int, load_queries, min_score=None):
        """Every rank sketches queries [lo,hi) of its slice, all-gathers, counts against its shard.
        Returns the merged ``(ptr, counts, gids)`` on rank 0 (``None`` elsewhere)."""
        lo, hi = shard_range(nq_total, self.world, self.rank)
        if hi > lo:
            local, _ = self.engine.sketch_many(load_queries(lo, hi))
        else:
            local = np.zeros((0, self.engine.F), np.int32)
        allsk = self.allgather_sketches(local)
        if self.dist and self.device is None:
            allsk = allsk.numpy()
        if self.gid_hi > self.gid_lo:
            part = self.engine.query_sketches(allsk, min_score)
        else:
            part = (np.zeros(nq_total + 1, np.uint64), np.zeros(0, np.uint32), np.zeros(0, np.uint32))
        return self.gather_and_merge(part, nq_total)

    def gather_and_merge(self, part, nq: int):
        if not self.dist:
            return merge_hit_lists([part], nq)
        parts = [None] * self.world if self.rank == 0 else None
        self.dist.gather_object(tuple(np.asarray(x) for x in part), parts, dst=0)
        if self.rank != 0:
            return None
        return merge_hit_lists(parts, nq)


# ---- the library's own exchange step (NCCL behind the C ABI) ----------------------------------
class Comm:
    """One rank's ``nq_comm`` (include/niqki_b200.h): created from a 128-byte NCCL id made on rank 0
    and handed to the other ranks by ``bcast_bytes`` (e.g. over torch.distributed)."""

    def __init__(self, ctx, rank: int, world: int, bcast_bytes):
        import ctypes as C

        from .capi import check, lib

        self.L, self.ctx, self.rank, self.world = lib(), ctx, rank, world
        ident = (C.c_ubyte * 128)()
        if rank == 0:
            check(self.L.nq_comm_unique_id(ident))
        raw = bcast_bytes(bytes(ident))
        ident = (C.c_ubyte * 128).from_buffer_copy(raw)
        h = C.c_void_p()
        check(self.L.nq_comm_init_rank(ctx.h, ident, world, rank, C.byref(h)))
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.L.nq_comm_destroy(self.h)
            self.h = None

    __del__ = close

    def info(self):
        import ctypes as C

        from .capi import check

        r, n, v = C.c_int(), C.c_int(), C.c_int()
        check(self.L.nq_comm_info(self.h, C.byref(r), C.byref(n), C.byref(v)))
        return dict(rank=r.value, nranks=n.value, nccl_version=v.value)

    def allgather_sketches(self, params, local, out):
        """``local`` int32 [n_local][F], ``out`` int32 [world*n_local][F]: device tensors."""
        import ctypes as C

        from .capi import check

        check(self.L.nq_allgather_sketches(self.h, C.byref(params), C.c_void_p(local.data_ptr()), int(local.shape[0]),
                                           C.c_void_p(out.data_ptr())))
        return out

    def bcast_sketches(self, params, sketches, root: int):
        import ctypes as C

        from .capi import check

        check(self.L.nq_bcast_sketches(self.h, C.byref(params), C.c_void_p(sketches.data_ptr()), int(sketches.shape[0]), root))
        return sketches


def torch_bcast_bytes(dist, device):
    """``bcast_bytes`` for ``Comm`` over an initialised torch.distributed group."""

    def f(raw: bytes) -> bytes:
        import torch

        t = torch.tensor(list(raw), dtype=torch.uint8, device=device)
        dist.broadcast(t, src=0)
        return bytes(t.cpu().tolist())

    return f


def merge_hits(parts):
    """``nq_hits_merge`` over per-shard ``(ptr, counts, gids)`` triples of the same query batch."""
    import ctypes as C

    from .capi import check, lib

    L = lib()
    nq = len(parts[0][0]) - 1
    handles = (C.c_void_p * len(parts))()
    keep = []
    for i, (p, c, g) in enumerate(parts):
        p = np.ascontiguousarray(p, np.uint64)
        c = np.ascontiguousarray(c if len(c) else np.zeros(1, np.uint32), np.uint32)
        g = np.ascontiguousarray(g if len(g) else np.zeros(1, np.uint32), np.uint32)
        keep.append((p, c, g))
        h = C.c_void_p()
        check(L.nq_hits_from_arrays(p.ctypes.data, c.ctypes.data, g.ctypes.data, nq, C.byref(h)))
        handles[i] = h
    out = C.c_void_p()
    try:
        check(L.nq_hits_merge(handles, len(parts), C.byref(out)))
        total = int(L.nq_hits_total(out))
        ptr = np.ctypeslib.as_array(L.nq_hits_ptr(out), shape=(nq + 1,)).copy()
        if total:
            counts = np.ctypeslib.as_array(L.nq_hits_counts(out), shape=(total,)).copy()
            gids = np.ctypeslib.as_array(L.nq_hits_gids(out), shape=(total,)).copy()
        else:
            counts, gids = np.zeros(0, np.uint32), np.zeros(0, np.uint32)
    finally:
        for h in handles:
            L.nq_hits_free(h)
        if out:
            L.nq_hits_free(out)
    return ptr, counts, gids
