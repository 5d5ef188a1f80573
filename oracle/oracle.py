"""ctypes front-ends for the TEST-ONLY checkers.

* ``Oracle``  – liboracle.so, the plain-C restatement (oracle/niqki_oracle.c).
* ``Ref``     – oracle/_ref/libniqki_ref.so, the unmodified reference behind oracle/ref_shim.cpp
                (present when ``make -C oracle ref`` ran where /root/reference is mounted).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import
this module.  Nothing under niqki_b200/ does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_ORACLE = os.path.join(HERE, "liboracle.so")
LIB_REF = os.path.join(HERE, "_ref", "libniqki_ref.so")
REF_CLI = os.path.join(HERE, "_ref", "niqki_ref")

_u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_u64p = np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS")
_u16p = np.ctypeslib.ndpointer(np.uint16, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")


def build(ref: bool = True) -> None:
    """Compile liboracle.so (always) and oracle/_ref (only where the reference tree is mounted)."""
    subprocess.run(["make", "-s", "-C", HERE, "all"], check=True)
    if ref and os.path.isdir("/root/reference/src"):
        subprocess.run(["make", "-s", "-C", HERE, "ref"], check=True)


class Params(C.Structure):
    """Mirror of ``nqo_params`` (== scalar fields of class Index, src/niqki_index.h:38-50)."""

    _fields_ = [(n, C.c_uint32) for n in ("K", "S", "W", "H", "M", "F", "mask_M", "maxrem")] + [
        ("range", C.c_int32),
        ("min_score", C.c_uint32),
    ]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


def _as_bytes(seq) -> np.ndarray:
    if isinstance(seq, str):
        seq = seq.encode("latin-1")
    if isinstance(seq, (bytes, bytearray)):
        return np.frombuffer(bytes(seq), dtype=np.uint8)
    return np.ascontiguousarray(seq, dtype=np.uint8)


def concat_entries(seqs):
    """list of sequences -> (bases u8, offsets u64[n+1])"""
    arrs = [_as_bytes(s) for s in seqs]
    offs = np.zeros(len(arrs) + 1, dtype=np.uint64)
    if arrs:
        offs[1:] = np.cumsum([a.size for a in arrs], dtype=np.uint64)
    bases = np.concatenate(arrs) if arrs else np.zeros(0, np.uint8)
    if bases.size == 0:
        bases = np.zeros(1, np.uint8)
    return np.ascontiguousarray(bases), offs


class Oracle:
    """The C restatement.  One instance == one parameter block (+ optionally one index)."""

    _lib = None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            if not os.path.exists(LIB_ORACLE):
                build(ref=False)
            L = C.CDLL(LIB_ORACLE)
            P = C.POINTER(Params)
            L.nqo_params_init.argtypes = [P, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_double]
            L.nqo_select_best_H.argtypes = [P, C.c_double]
            for f in ("nqo_revhash64", "nqo_unrevhash64", "nqo_mix"):
                getattr(L, f).argtypes = [C.c_uint64]
                getattr(L, f).restype = C.c_uint64
            L.nqo_hash_family.argtypes = [C.c_uint64, C.c_uint32]
            L.nqo_hash_family.restype = C.c_uint64
            L.nqo_get_fingerprint.argtypes = [P, C.c_uint64]
            L.nqo_get_fingerprint.restype = C.c_int32
            L.nqo_str2numstrand.argtypes = [C.c_char_p, C.c_size_t]
            L.nqo_str2numstrand.restype = C.c_uint64
            L.nqo_rcb.argtypes = [P, C.c_uint64]
            L.nqo_rcb.restype = C.c_uint64
            L.nqo_compute_sketch.argtypes = [P, _u8p, C.c_size_t, _i32p, C.c_long]
            L.nqo_compute_sketch.restype = C.c_long
            L.nqo_sketch_scan.argtypes = [P, _u8p, C.c_size_t, _i32p]
            L.nqo_sketch_scan.restype = C.c_uint32
            L.nqo_sketch_densification.argtypes = [P, _i32p, C.c_uint32, C.c_long]
            L.nqo_sketch_densification.restype = C.c_long
            L.nqo_sketch_batch.argtypes = [P, _u8p, _u64p, C.c_size_t, _i32p, C.c_int]
            L.nqo_index_new.argtypes = [P]
            L.nqo_index_new.restype = C.c_void_p
            L.nqo_index_free.argtypes = [C.c_void_p]
            L.nqo_index_insert.argtypes = [C.c_void_p, _i32p, C.c_uint32]
            L.nqo_index_finalize.argtypes = [C.c_void_p]
            L.nqo_index_num_genomes.argtypes = [C.c_void_p]
            L.nqo_index_num_genomes.restype = C.c_uint32
            L.nqo_index_num_postings.argtypes = [C.c_void_p]
            L.nqo_index_num_postings.restype = C.c_uint64
            L.nqo_index_row_ptr.argtypes = [C.c_void_p]
            L.nqo_index_row_ptr.restype = C.POINTER(C.c_uint64)
            L.nqo_index_gids.argtypes = [C.c_void_p]
            L.nqo_index_gids.restype = C.POINTER(C.c_uint32)
            L.nqo_query_sketch.argtypes = [C.c_void_p, _i32p, _u32p, _u32p, C.c_size_t]
            L.nqo_query_sketch.restype = C.c_size_t
            L.nqo_query_batch.argtypes = [C.c_void_p, _i32p, C.c_size_t, C.c_void_p, C.c_void_p,
                                          C.c_void_p, C.c_size_t, C.c_int]
            L.nqo_query_batch.restype = C.c_size_t
            L.nqo_matrix_counts.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, _u16p, C.c_int]
            L.nqo_synth_genome.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, _u8p]
            L.nqo_synth_mutant.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_double, C.c_uint64, _u8p]
            L.nqo_synth_read.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32, _u8p]
            cls._lib = L
        return cls._lib

    def __init__(self, K=31, S=15, W=12, H=4, J=0.0, genome_size=0):
        self.L = self.lib()
        self.p = Params()
        self.L.nqo_params_init(C.byref(self.p), K, S, W, H, float(J))
        if genome_size:
            self.L.nqo_select_best_H(C.byref(self.p), float(genome_size))
        self._ix = None

    def __del__(self):
        if getattr(self, "_ix", None):
            self.L.nqo_index_free(self._ix)
            self._ix = None

    # -- scalars
    @property
    def F(self):
        return int(self.p.F)

    def revhash64(self, x):
        return int(self.L.nqo_revhash64(x))

    def unrevhash64(self, x):
        return int(self.L.nqo_unrevhash64(x))

    def hash_family(self, x, f):
        return int(self.L.nqo_hash_family(x, f))

    def get_fingerprint(self, h):
        return int(self.L.nqo_get_fingerprint(C.byref(self.p), h))

    def str2numstrand(self, s: bytes):
        return int(self.L.nqo_str2numstrand(s, len(s)))

    def rcb(self, x):
        return int(self.L.nqo_rcb(C.byref(self.p), x))

    # -- sketches
    def compute_sketch(self, seq, sketch=None, max_passes=0):
        """Index::compute_sketch; returns (sketch int32[F], passes)."""
        b = _as_bytes(seq)
        if sketch is None:
            sketch = np.full(self.F, -1, np.int32)
        bb = b if b.size else np.zeros(1, np.uint8)
        passes = self.L.nqo_compute_sketch(C.byref(self.p), bb, b.size, sketch, max_passes)
        return sketch, int(passes)

    def sketch_scan(self, seq):
        b = _as_bytes(seq)
        sketch = np.full(self.F, -1, np.int32)
        bb = b if b.size else np.zeros(1, np.uint8)
        filled = self.L.nqo_sketch_scan(C.byref(self.p), bb, b.size, sketch)
        return sketch, int(filled)

    def densify(self, sketch, max_passes=0):
        sk = np.ascontiguousarray(sketch, np.int32).copy()
        empty = int((sk == -1).sum())
        passes = self.L.nqo_sketch_densification(C.byref(self.p), sk, empty, max_passes)
        return sk, int(passes)

    def sketch_batch(self, bases, offsets, nthreads=0):
        n = len(offsets) - 1
        out = np.empty((n, self.F), np.int32)
        self.L.nqo_sketch_batch(C.byref(self.p), bases, np.ascontiguousarray(offsets, np.uint64), n,
                                out.reshape(-1) if n else np.zeros(1, np.int32), nthreads)
        return out

    def sketch_many(self, seqs, nthreads=0):
        bases, offs = concat_entries(seqs)
        return self.sketch_batch(bases, offs, nthreads)

    # -- index
    def index_reset(self):
        if self._ix:
            self.L.nqo_index_free(self._ix)
        self._ix = self.L.nqo_index_new(C.byref(self.p))

    def insert_sketch(self, sketch, gid):
        if not self._ix:
            self.index_reset()
        self.L.nqo_index_insert(self._ix, np.ascontiguousarray(sketch, np.int32), gid)

    def insert_sketches(self, sketches, gid_base=0):
        for i, sk in enumerate(sketches):
            self.insert_sketch(sk, gid_base + i)

    @property
    def num_genomes(self):
        return int(self.L.nqo_index_num_genomes(self._ix)) if self._ix else 0

    def csr(self):
        """(row_ptr u64[range*F+1], gids u32[postings]) — copies"""
        nrows = int(self.p.range) * self.F
        npost = int(self.L.nqo_index_num_postings(self._ix))
        rp = np.ctypeslib.as_array(self.L.nqo_index_row_ptr(self._ix), shape=(nrows + 1,)).copy()
        g = (np.ctypeslib.as_array(self.L.nqo_index_gids(self._ix), shape=(max(npost, 1),))[:npost]).copy()
        return rp, g

    def query_sketch(self, sketch):
        """Index::query_sketch -> (counts u32[], gids u32[]) sorted (count,gid) descending."""
        n = max(self.num_genomes, 1)
        c = np.empty(n, np.uint32)
        g = np.empty(n, np.uint32)
        nh = self.L.nqo_query_sketch(self._ix, np.ascontiguousarray(sketch, np.int32), c, g, n)
        return c[:nh].copy(), g[:nh].copy()

    def query_batch(self, sketches, nthreads=0):
        sk = np.ascontiguousarray(sketches, np.int32)
        nq = sk.shape[0]
        hp = np.zeros(nq + 1, np.uint64)
        flat = sk.reshape(-1) if nq else np.zeros(1, np.int32)
        total = self.L.nqo_query_batch(self._ix, flat, nq, hp.ctypes.data, None, None, 0, nthreads)
        c = np.empty(max(total, 1), np.uint32)
        g = np.empty(max(total, 1), np.uint32)
        self.L.nqo_query_batch(self._ix, flat, nq, hp.ctypes.data, c.ctypes.data, g.ctypes.data, total, nthreads)
        return hp, c[:total], g[:total]

    def matrix_counts(self, begin=0, end=None, nthreads=0):
        N = self.num_genomes
        end = N if end is None else end
        out = np.zeros((max(end - begin, 1), max(N, 1)), np.uint16)
        self.L.nqo_matrix_counts(self._ix, begin, end, out.reshape(-1), nthreads)
        return out[: end - begin, :N]

    # -- synthetic inputs
    def mix(self, x):
        return int(self.L.nqo_mix(x))

    def synth_genome(self, g, length, seed=42):
        out = np.empty(length, np.uint8)
        self.L.nqo_synth_genome(seed, g, length, out)
        return out

    def synth_mutant(self, g, q, d, length, seed=42):
        out = np.empty(length, np.uint8)
        self.L.nqo_synth_mutant(seed, g, q, float(d), length, out)
        return out

    def synth_read(self, r, genome_len, read_len=150, seed=42):
        out = np.empty(read_len, np.uint8)
        self.L.nqo_synth_read(seed, r, genome_len, read_len, out)
        return out


def ref_available() -> bool:
    return os.path.exists(LIB_REF)


class Ref:
    """The unmodified reference (class Index) through oracle/ref_shim.cpp."""

    _lib = None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            L = C.CDLL(LIB_REF)
            L.ref_index_new.argtypes = [C.c_uint32] * 4 + [C.c_char_p, C.c_double]
            L.ref_index_new.restype = C.c_void_p
            L.ref_index_free.argtypes = [C.c_void_p]
            L.ref_select_best_H.argtypes = [C.c_void_p, C.c_double]
            L.ref_get_params.argtypes = [C.c_void_p, _u32p]
            for f in ("ref_revhash64", "ref_unrevhash64", "ref_rcb"):
                getattr(L, f).argtypes = [C.c_void_p, C.c_uint64]
                getattr(L, f).restype = C.c_uint64
            L.ref_hash_family.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32]
            L.ref_hash_family.restype = C.c_uint64
            L.ref_get_fingerprint.argtypes = [C.c_void_p, C.c_uint64]
            L.ref_get_fingerprint.restype = C.c_int32
            L.ref_str2numstrand.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t]
            L.ref_str2numstrand.restype = C.c_uint64
            L.ref_compute_sketch.argtypes = [C.c_void_p, _u8p, C.c_size_t, _i32p]
            L.ref_sketch_batch.argtypes = [C.c_void_p, _u8p, _u64p, C.c_size_t, C.c_void_p, C.c_int]
            L.ref_sketch_batch.restype = C.c_double
            L.ref_insert_sketch.argtypes = [C.c_void_p, _i32p, C.c_uint32, C.c_char_p]
            L.ref_num_genomes.argtypes = [C.c_void_p]
            L.ref_num_genomes.restype = C.c_uint32
            L.ref_export_postings.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]
            L.ref_export_postings.restype = C.c_uint64
            L.ref_query_sketch.argtypes = [C.c_void_p, _i32p, _u32p, _u32p, C.c_size_t]
            L.ref_query_sketch.restype = C.c_size_t
            L.ref_query_batch.argtypes = [C.c_void_p, _i32p, C.c_size_t, C.POINTER(C.c_uint64), C.c_int]
            L.ref_query_batch.restype = C.c_double
            L.ref_query_matrix.argtypes = [C.c_void_p]
            L.ref_output_query.argtypes = [C.c_void_p, _u32p, _u32p, C.c_size_t, C.c_char_p, C.c_int]
            L.ref_flush_output.argtypes = [C.c_void_p]
            L.ref_dump.argtypes = [C.c_void_p, C.c_char_p]
            cls._lib = L
        return cls._lib

    def __init__(self, K=31, S=15, W=12, H=4, J=0.0, genome_size=0, out_path=None):
        self.L = self.lib()
        if out_path is None:
            fd, out_path = tempfile.mkstemp(prefix="niqki_ref_", suffix=".gz")
            os.close(fd)
            self._tmp = out_path
        else:
            self._tmp = None
        self.out_path = out_path
        self.h = self.L.ref_index_new(S, K, W, H, out_path.encode(), float(J))
        if genome_size:
            self.L.ref_select_best_H(self.h, float(genome_size))
        prm = np.zeros(10, np.uint32)
        self.L.ref_get_params(self.h, prm)
        self.params = dict(zip(("K", "S", "W", "H", "M", "F", "mask_M", "maxrem", "range", "min_score"),
                               (int(v) for v in prm)))
        self.F = self.params["F"]

    def close(self):
        if self.h:
            self.L.ref_index_free(self.h)  # closes the output stream (src/niqki_index.cpp:106-109)
            self.h = None
        if self._tmp and os.path.exists(self._tmp):
            os.unlink(self._tmp)
            self._tmp = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def revhash64(self, x):
        return int(self.L.ref_revhash64(self.h, x))

    def unrevhash64(self, x):
        return int(self.L.ref_unrevhash64(self.h, x))

    def hash_family(self, x, f):
        return int(self.L.ref_hash_family(self.h, x, f))

    def get_fingerprint(self, x):
        return int(self.L.ref_get_fingerprint(self.h, x))

    def str2numstrand(self, s: bytes):
        return int(self.L.ref_str2numstrand(self.h, s, len(s)))

    def rcb(self, x):
        return int(self.L.ref_rcb(self.h, x))

    def compute_sketch(self, seq, sketch=None):
        b = _as_bytes(seq)
        if sketch is None:
            sketch = np.full(self.F, -1, np.int32)
        self.L.ref_compute_sketch(self.h, b if b.size else np.zeros(1, np.uint8), b.size, sketch)
        return sketch

    def sketch_batch(self, bases, offsets, nthreads=0, want_out=True):
        n = len(offsets) - 1
        out = np.empty((n, self.F), np.int32) if want_out else None
        secs = self.L.ref_sketch_batch(self.h, bases, np.ascontiguousarray(offsets, np.uint64), n,
                                       out.ctypes.data if want_out else None, nthreads)
        return out, float(secs)

    def insert_sketch(self, sketch, gid, name=""):
        self.L.ref_insert_sketch(self.h, np.ascontiguousarray(sketch, np.int32), gid, name.encode())

    @property
    def num_genomes(self):
        return int(self.L.ref_num_genomes(self.h))

    def export_postings(self):
        rows = self.params["range"] * self.F
        sizes = np.zeros(rows, np.uint32)
        n = int(self.L.ref_export_postings(self.h, sizes.ctypes.data, None, 0))
        gids = np.zeros(max(n, 1), np.uint32)
        self.L.ref_export_postings(self.h, sizes.ctypes.data, gids.ctypes.data, n)
        return sizes, gids[:n]

    def query_sketch(self, sketch):
        n = max(self.num_genomes, 1)
        c = np.empty(n, np.uint32)
        g = np.empty(n, np.uint32)
        nh = self.L.ref_query_sketch(self.h, np.ascontiguousarray(sketch, np.int32), c, g, n)
        return c[:nh].copy(), g[:nh].copy()

    def query_batch_timed(self, sketches, nthreads=0):
        sk = np.ascontiguousarray(sketches, np.int32)
        hits = C.c_uint64(0)
        secs = self.L.ref_query_batch(self.h, sk.reshape(-1), sk.shape[0], C.byref(hits), nthreads)
        return float(secs), int(hits.value)

    def query_matrix(self):
        self.L.ref_query_matrix(self.h)
        self.L.ref_flush_output(self.h)

    def output_query(self, counts, gids, name, pretty=True):
        self.L.ref_output_query(self.h, np.ascontiguousarray(counts, np.uint32),
                                np.ascontiguousarray(gids, np.uint32), len(counts), name.encode(), int(pretty))
        self.L.ref_flush_output(self.h)

    def dump(self, path):
        self.L.ref_dump(self.h, path.encode())

    def read_output(self) -> bytes:
        """Decompressed bytes written so far (every flush ends a gzip member)."""
        import zlib

        data = open(self.out_path, "rb").read()
        out = b""
        while data:
            d = zlib.decompressobj(31)
            out += d.decompress(data)
            data = d.unused_data
        return out
