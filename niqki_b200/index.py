"""Host-side mirror of the reference's ``class Index`` L1 surface over the C ABI.

Method names and argument meaning follow /root/reference/src/niqki_index.h (compute_sketch :104,
insert_sketch :108, query_sketch :142, query_matrix :206, query_range :208, select_best_H :211);
the batch forms are what the reference's OpenMP file loops amount to.  numpy arrays are host
buffers; objects exposing ``data_ptr()`` (torch CUDA tensors) are passed as device pointers.
"""
from __future__ import annotations

import ctypes as C
import weakref

import numpy as np

from . import capi
from .capi import Params, check, lib


def _is_device(x) -> bool:
    return hasattr(x, "data_ptr") and getattr(x, "is_cuda", False)


def _ptr(x):
    if x is None:
        return None
    if hasattr(x, "data_ptr"):
        return C.c_void_p(x.data_ptr())
    return C.c_void_p(x.ctypes.data)


class Context:
    """One CUDA device + stream (``nq_ctx``)."""

    def __init__(self, device: int = 0, stream=None):
        self.L = lib()
        h = C.c_void_p()
        sp = None
        if stream is not None:
            sp = C.c_void_p(int(getattr(stream, "cuda_stream", stream)))
        check(self.L.nq_ctx_create(device, sp, C.byref(h)))
        self.h = h
        self.device = device
        # a context that owns its stream is not ordered against the caller's (torch's) stream: the
        # device-pointer methods of Index then fence on both sides (see Index._fence_in / _fence_out)
        self.own_stream = stream is None
        self._indexes = weakref.WeakSet()  # posting lists living on this context: freed before it (nq_index_free uses the ctx)

    def close(self):
        if getattr(self, "h", None):
            for ix in list(getattr(self, "_indexes", ())):
                ix.close_index()
            self.L.nq_ctx_destroy(self.h)
            self.h = None

    __del__ = close

    def sync(self):
        check(self.L.nq_ctx_sync(self.h))

    @property
    def launches(self) -> int:
        return int(self.L.nq_ctx_launch_count(self.h))

    KINDS = ("scan", "densify", "transpose", "cell_sort", "query", "matrix", "slab")

    def set_host_packing(self, mode: int = -1, threads: int = 0):
        """K1 on the host side of the host-buffer sketch calls: -1 automatic, 0 never, 1 always."""
        check(self.L.nq_ctx_set_host_packing(self.h, int(mode), int(threads)))

    def set_timing(self, on: bool = True):
        check(self.L.nq_ctx_set_timing(self.h, int(on)))

    def timing_reset(self):
        check(self.L.nq_ctx_timing_reset(self.h))

    def timing(self):
        """{kernel family: (device ms, launches)} accumulated since the last reset."""
        out = {}
        for k, name in enumerate(self.KINDS):
            ms, n = C.c_double(), C.c_uint64()
            check(self.L.nq_ctx_timing(self.h, k, C.byref(ms), C.byref(n)))
            out[name] = (ms.value, int(n.value))
        return out

    @property
    def h2d_bytes(self) -> int:
        """Bytes of sequence (packed or not) the host-buffer sketch calls have copied to the device so far."""
        return int(self.L.nq_ctx_h2d_bytes(self.h))

    @property
    def last_query_gathered(self) -> int:
        return int(self.L.nq_ctx_last_query_gathered(self.h))


class Index:
    """Parameters + (optionally) one index shard in HBM.

    ``Index(S, K, W, H, min_fract)`` follows the reference constructor's argument order
    (src/niqki_index.cpp:13) minus the output filename, which belongs to the host writers.
    """

    def __init__(self, S=15, K=31, W=12, H=4, min_fract=0.0, ctx: Context | None = None, device: int = 0,
                 genome_size: float = 0):
        self.L = lib()
        self.p = Params()
        check(self.L.nq_params_init(C.byref(self.p), K, S, W, H, float(min_fract)))
        if genome_size:
            self.select_best_H(genome_size)
        self._own_ctx = ctx is None
        self.ctx = ctx if ctx is not None else Context(device)
        self.ctx._indexes.add(self)
        self.ix = None
        self.gid_base = 0
        self.genome_numbers = 0

    # ---- parameters
    @property
    def F(self):
        return int(self.p.F)

    @property
    def min_score(self):
        return int(self.p.min_score)

    def select_best_H(self, genome_size: float):
        check(self.L.nq_params_select_best_H(C.byref(self.p), float(genome_size)))

    def close(self):
        if getattr(self, "ix", None):
            self.L.nq_index_free(self.ix)
            self.ix = None
        if getattr(self, "_own_ctx", False) and getattr(self, "ctx", None):
            self.ctx.close()
            self.ctx = None

    __del__ = close

    # ---- stream ordering of the device-pointer forms: the library launches on the context's stream and
    # returns without syncing.  When that stream is torch's current stream (Context(device, stream)) the
    # calls are ordered like any other torch op.  When the context owns a stream of its own, inputs
    # produced on torch's stream must be complete before the call and outputs before torch reads them.
    def _fence_in(self, t):
        if self.ctx.own_stream and _is_device(t):
            import torch

            torch.cuda.current_stream(t.device).synchronize()

    def _fence_out(self):
        if self.ctx.own_stream:
            self.ctx.sync()

    # ---- sketching (compute_sketch + sketch_densification)
    def compute_sketches(self, bases, offsets, out=None, flags=None):
        """Batch form of Index::compute_sketch on fresh sketches.  ``bases`` u8 (numpy = host,
        torch CUDA tensor = device), ``offsets`` u64[n+1] on the host.  Returns (sketches, flags)."""
        offsets = np.ascontiguousarray(offsets, np.uint64)
        n = offsets.size - 1
        if _is_device(bases):
            import torch

            if out is None:
                out = torch.empty((n, self.F), dtype=torch.int32, device=bases.device)
            if flags is None:
                flags = torch.empty((max(n, 1),), dtype=torch.int32, device=bases.device)
            self._fence_in(bases)
            check(self.L.nq_sketch_batch_device(self.ctx.h, C.byref(self.p), _ptr(bases), bases.numel(), _ptr(offsets),
                                                n, _ptr(out), _ptr(flags)))
            self._fence_out()
            return out, flags[:n]
        bases = np.ascontiguousarray(bases, np.uint8)
        if out is None:
            out = np.empty((n, self.F), np.int32)
        if flags is None:
            flags = np.zeros(max(n, 1), np.uint32)
        check(self.L.nq_sketch_batch(self.ctx.h, C.byref(self.p), _ptr(bases), _ptr(offsets), n, _ptr(out), _ptr(flags)))
        return out, flags[:n]

    def sketch_records_to_device(self, bases, offsets, out, flags=None, rec_entry=None, n_entries=None):
        """``nq_sketch_records`` with the sketches left in HBM: host characters in (pinned or pageable;
        long entries cross PCIe packed to 2 bits per base), ``out`` = device int32 [n_entries][F]."""
        offsets = np.ascontiguousarray(offsets, np.uint64)
        n_rec = offsets.size - 1
        ne = n_rec if n_entries is None else int(n_entries)
        bases = np.ascontiguousarray(bases, np.uint8)
        if flags is None:
            flags = np.zeros(max(ne, 1), np.uint32)
        re = None if rec_entry is None else np.ascontiguousarray(rec_entry, np.uint32)
        check(self.L.nq_sketch_records(self.ctx.h, C.byref(self.p), _ptr(bases), _ptr(offsets), n_rec, _ptr(re), ne, _ptr(out),
                                       _ptr(flags), 1))
        return out, flags[:ne]

    def close_index(self):
        """Free the posting lists (the parameters and the context stay)."""
        if getattr(self, "ix", None):
            self.L.nq_index_free(self.ix)
            self.ix = None

    def compute_sketch(self, seq):
        """Index::compute_sketch(reference, sketch) on a fresh sketch -> int32[F]."""
        if isinstance(seq, str):
            seq = seq.encode("latin-1")
        b = np.frombuffer(bytes(seq), np.uint8) if isinstance(seq, (bytes, bytearray)) else np.ascontiguousarray(seq, np.uint8)
        bb = b if b.size else np.zeros(1, np.uint8)
        sk, _ = self.compute_sketches(bb, np.array([0, b.size], np.uint64))
        return sk[0]

    def sketch_many(self, seqs):
        arrs = [np.frombuffer(bytes(s), np.uint8) if isinstance(s, (bytes, bytearray)) else np.ascontiguousarray(s, np.uint8)
                for s in seqs]
        offs = np.zeros(len(arrs) + 1, np.uint64)
        if arrs:
            offs[1:] = np.cumsum([a.size for a in arrs])
        bases = np.concatenate(arrs) if arrs and offs[-1] else np.zeros(1, np.uint8)
        return self.compute_sketches(bases, offs)

    def densify(self, sketches):
        """Index::sketch_densification on device sketches in place (torch int32 [n][F])."""
        import torch

        flags = torch.zeros((sketches.shape[0],), dtype=torch.int32, device=sketches.device)
        self._fence_in(sketches)
        check(self.L.nq_densify_device(self.ctx.h, C.byref(self.p), _ptr(sketches), sketches.shape[0], _ptr(flags)))
        self._fence_out()
        return sketches, flags

    # ---- index (insert_sketch over a batch)
    def insert_sketches(self, sketches, gid_base: int = 0):
        """Build this shard's posting lists from ``sketches`` [n][F]; gids gid_base..gid_base+n-1."""
        if self.ix:
            self.L.nq_index_free(self.ix)
            self.ix = None
        n = int(sketches.shape[0])
        h = C.c_void_p()
        if _is_device(sketches):
            self._fence_in(sketches)
            check(self.L.nq_index_build_device(self.ctx.h, C.byref(self.p), _ptr(sketches), n, gid_base, C.byref(h)))
        else:
            sk = np.ascontiguousarray(sketches, np.int32)
            check(self.L.nq_index_build(self.ctx.h, C.byref(self.p), _ptr(sk), n, gid_base, C.byref(h)))
        self.ix = h
        self.gid_base = gid_base
        self.genome_numbers = n

    def info(self):
        npost, ng, gb, db = C.c_uint64(), C.c_uint32(), C.c_uint32(), C.c_uint64()
        check(self.L.nq_index_info(self.ix, C.byref(npost), C.byref(ng), C.byref(gb), C.byref(db)))
        return dict(n_postings=npost.value, n_genomes=ng.value, gid_base=gb.value, device_bytes=db.value)

    def export_postings(self):
        """(list_sizes u32[range*F], gids u32[n_postings]) in list-id order (dump layout, A7)."""
        info = self.info()
        sizes = np.zeros(int(self.p.range) * self.F, np.uint32)
        gids = np.zeros(max(info["n_postings"], 1), np.uint32)
        check(self.L.nq_index_export(self.ix, _ptr(sizes), _ptr(gids), gids.size))
        return sizes, gids[: info["n_postings"]]

    def import_postings(self, sizes, gids, n_genomes, gid_base=0):
        if self.ix:
            self.L.nq_index_free(self.ix)
        h = C.c_void_p()
        sizes = np.ascontiguousarray(sizes, np.uint32)
        gids = np.ascontiguousarray(gids, np.uint32)
        g = gids if gids.size else np.zeros(1, np.uint32)
        check(self.L.nq_index_import(self.ctx.h, C.byref(self.p), _ptr(sizes), _ptr(g), n_genomes, gid_base, C.byref(h)))
        self.ix, self.gid_base, self.genome_numbers = h, gid_base, n_genomes

    # ---- query (query_sketch over a batch)
    def query_sketches(self, sketches, min_score=None, fetch=True):
        """-> (hit_ptr u64[nq+1], counts u32[], gids u32[]) each query sorted (count,gid) descending.
        ``fetch=False`` (device sketches only) leaves the hits in HBM and returns None."""
        ms = self.min_score if min_score is None else int(min_score)
        nq = int(sketches.shape[0])
        h = C.c_void_p()
        if _is_device(sketches):
            self._fence_in(sketches)
            check(self.L.nq_query_batch_device(self.ix, _ptr(sketches), nq, ms, C.byref(h) if fetch else None))
            if not fetch:
                return None
        else:
            sk = np.ascontiguousarray(sketches, np.int32)
            check(self.L.nq_query_batch(self.ix, _ptr(sk if nq else np.zeros(1, np.int32)), nq, ms, C.byref(h)))
        try:
            total = int(self.L.nq_hits_total(h))
            ptr = np.ctypeslib.as_array(self.L.nq_hits_ptr(h), shape=(nq + 1,)).copy()
            if total:
                counts = np.ctypeslib.as_array(self.L.nq_hits_counts(h), shape=(total,)).copy()
                gids = np.ctypeslib.as_array(self.L.nq_hits_gids(h), shape=(total,)).copy()
            else:
                counts, gids = np.zeros(0, np.uint32), np.zeros(0, np.uint32)
        finally:
            self.L.nq_hits_free(h)
        return ptr, counts, gids

    def query_sketch(self, sketch, min_score=None):
        """Index::query_sketch -> (counts, gids) sorted by (count, gid) descending."""
        ptr, c, g = self.query_sketches(np.ascontiguousarray(sketch, np.int32)[None, :], min_score)
        return c, g

    # ---- all-vs-all (query_range / query_matrix)
    def query_range(self, begin: int, end: int, wrap16: bool = True):
        """Integer counts [end-begin][n]; wrap16=True is the reference's uint16 behaviour (B6)."""
        n = self.genome_numbers
        out = np.zeros((max(end - begin, 1), max(n, 1)), np.uint32)
        check(self.L.nq_matrix_rows(self.ix, begin, end, int(wrap16), _ptr(out)))
        return out[: end - begin, :n]

    def matrix_tile(self, row_sketches, wrap16: bool = True):
        """One tile of the genome x genome grid: counts [rows][n_shard] of the given row sketches (device
        int32 [rows][F], from any shard) against this shard's columns (nq_matrix_tile)."""
        rows = int(row_sketches.shape[0])
        n = self.genome_numbers
        out = np.zeros((max(rows, 1), max(n, 1)), np.uint32)
        self._fence_in(row_sketches)
        check(self.L.nq_matrix_tile(self.ix, _ptr(row_sketches), rows, int(wrap16), _ptr(out)))
        return out[:rows, :n]

    def index_sketches(self, begin: int, end: int, out):
        """Sketches of local genomes [begin, end) rebuilt from the posting lists into ``out`` (device)."""
        check(self.L.nq_index_sketches_device(self.ix, begin, end, _ptr(out)))
        return out

    def query_matrix(self, wrap16: bool = True):
        return self.query_range(0, self.genome_numbers, wrap16)
