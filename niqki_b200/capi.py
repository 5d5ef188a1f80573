"""ctypes binding of include/niqki_b200.h (libniqki_b200.so).  No compute happens in Python."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
# NIQKI_B200_LIB: another build of the same library (the measurement build of `make tuning`)
LIB_PATH = os.environ.get("NIQKI_B200_LIB") or os.path.join(HERE, "lib", "libniqki_b200.so")
CSRC = os.path.join(HERE, "csrc")

NQ_OK, NQ_ERR_INVALID, NQ_ERR_CUDA, NQ_ERR_UNSUPPORTED, NQ_ERR_OVERFLOW = range(5)
NQ_ENTRY_SKIPPED, NQ_ENTRY_DENSIFY_STALLED = 1, 2


class NiqkiError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libniqki_b200 error {code}: {msg}")
        self.code = code


class Params(C.Structure):
    """``nq_params`` == the scalar fields of class Index (src/niqki_index.h:38-50)."""

    _fields_ = [(n, C.c_uint32) for n in ("K", "S", "W", "H", "M", "F", "mask_M", "maxrem")] + [
        ("range", C.c_int32),
        ("min_score", C.c_uint32),
    ]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


# every symbol include/niqki_b200.h declares: name -> (restype, argtypes)
_P = C.POINTER(Params)
_VP = C.c_void_p
SYMBOLS = {
    "nq_last_error": (C.c_char_p, []),
    "nq_version": (C.c_char_p, []),
    "nq_params_init": (C.c_int, [_P, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_double]),
    "nq_params_select_best_H": (C.c_int, [_P, C.c_double]),
    "nq_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "nq_ctx_create": (C.c_int, [C.c_int, _VP, C.POINTER(_VP)]),
    "nq_ctx_destroy": (C.c_int, [_VP]),
    "nq_ctx_sync": (C.c_int, [_VP]),
    "nq_ctx_launch_count": (C.c_uint64, [_VP]),
    "nq_ctx_set_timing": (C.c_int, [_VP, C.c_int]),
    "nq_ctx_timing": (C.c_int, [_VP, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_uint64)]),
    "nq_ctx_timing_reset": (C.c_int, [_VP]),
    "nq_ctx_last_query_gathered": (C.c_uint64, [_VP]),
    "nq_ctx_h2d_bytes": (C.c_uint64, [_VP]),
    "nq_host_alloc": (_VP, [C.c_size_t]),
    "nq_host_free": (None, [_VP]),
    "nq_sketch_batch": (C.c_int, [_VP, _P, _VP, _VP, C.c_uint64, _VP, _VP]),
    "nq_sketch_records": (C.c_int, [_VP, _P, _VP, _VP, C.c_uint64, _VP, C.c_uint64, _VP, _VP, C.c_int]),
    "nq_device_alloc": (C.c_int, [_VP, C.c_size_t, C.POINTER(_VP)]),
    "nq_device_free": (C.c_int, [_VP, _VP]),
    "nq_device_copy": (C.c_int, [_VP, _VP, _VP, C.c_size_t, C.c_int]),
    "nq_device_copy_peer": (C.c_int, [_VP, _VP, _VP, _VP, C.c_size_t]),
    "nq_device_fill": (C.c_int, [_VP, _VP, C.c_int, C.c_size_t]),
    "nq_ctx_set_host_packing": (C.c_int, [_VP, C.c_int, C.c_uint]),
    "nq_pack_sizes": (C.c_int, [C.c_uint64, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "nq_pack_sequences": (C.c_int, [_VP, _VP, C.c_uint64, C.c_uint32, _VP, _VP, _VP, C.c_uint64, C.POINTER(C.c_uint64), C.c_uint]),
    "nq_sketch_batch_packed_device": (C.c_int, [_VP, _P, _VP, _VP, _VP, _VP, C.c_uint64, _VP, _VP]),
    "nq_sketch_batch_device": (C.c_int, [_VP, _P, _VP, C.c_uint64, _VP, C.c_uint64, _VP, _VP]),
    "nq_densify_device": (C.c_int, [_VP, _P, _VP, C.c_uint64, _VP]),
    "nq_index_sketches_device": (C.c_int, [_VP, C.c_uint32, C.c_uint32, _VP]),
    "nq_matrix_tile": (C.c_int, [_VP, _VP, C.c_uint32, C.c_int, _VP]),
    "nq_comm_unique_id": (C.c_int, [_VP]),
    "nq_comm_init_rank": (C.c_int, [_VP, _VP, C.c_int, C.c_int, C.POINTER(_VP)]),
    "nq_comm_init_all": (C.c_int, [C.POINTER(_VP), C.c_int, C.POINTER(_VP)]),
    "nq_comm_destroy": (C.c_int, [_VP]),
    "nq_comm_info": (C.c_int, [_VP, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "nq_shard_range": (C.c_int, [C.c_uint64, C.c_int, C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "nq_allgather_sketches": (C.c_int, [_VP, _P, _VP, C.c_uint64, _VP]),
    "nq_bcast_sketches": (C.c_int, [_VP, _P, _VP, C.c_uint64, C.c_int]),
    "nq_hits_merge": (C.c_int, [C.POINTER(_VP), C.c_int, C.POINTER(_VP)]),
    "nq_hits_from_arrays": (C.c_int, [_VP, _VP, _VP, C.c_uint64, C.POINTER(_VP)]),
    "nq_index_build": (C.c_int, [_VP, _P, _VP, C.c_uint64, C.c_uint32, C.POINTER(_VP)]),
    "nq_index_build_device": (C.c_int, [_VP, _P, _VP, C.c_uint64, C.c_uint32, C.POINTER(_VP)]),
    "nq_index_free": (C.c_int, [_VP]),
    "nq_index_info": (C.c_int, [_VP, C.POINTER(C.c_uint64), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32),
                                C.POINTER(C.c_uint64)]),
    "nq_index_export": (C.c_int, [_VP, _VP, _VP, C.c_uint64]),
    "nq_index_import": (C.c_int, [_VP, _P, _VP, _VP, C.c_uint32, C.c_uint32, C.POINTER(_VP)]),
    "nq_query_batch": (C.c_int, [_VP, _VP, C.c_uint64, C.c_uint32, C.POINTER(_VP)]),
    "nq_query_batch_device": (C.c_int, [_VP, _VP, C.c_uint64, C.c_uint32, C.POINTER(_VP)]),
    "nq_hits_total": (C.c_uint64, [_VP]),
    "nq_hits_ptr": (C.POINTER(C.c_uint64), [_VP]),
    "nq_hits_counts": (C.POINTER(C.c_uint32), [_VP]),
    "nq_hits_gids": (C.POINTER(C.c_uint32), [_VP]),
    "nq_hits_free": (None, [_VP]),
    "nq_matrix_rows": (C.c_int, [_VP, C.c_uint32, C.c_uint32, C.c_int, _VP]),
    "nq_synth_genomes_device": (C.c_int, [_VP, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, _VP]),
    "nq_synth_mutants_device": (C.c_int, [_VP, C.c_uint64, _VP, _VP, _VP, C.c_uint64, C.c_uint64, _VP]),
    "nq_synth_reads_device": (C.c_int, [_VP, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32, _VP]),
}

_lib = None


def library_available() -> bool:
    return os.path.exists(LIB_PATH)


def build_library(verbose: bool = False) -> str:
    """nvcc -gencode arch=compute_100a,code=sm_100a ... (niqki_b200/csrc/Makefile), in-tree."""
    subprocess.run(["make", "-C", CSRC, "-j8"] + ([] if verbose else ["-s"]), check=True)
    return LIB_PATH


def lib():
    """Load libniqki_b200.so; fails loudly when it has not been built (no fallback path exists)."""
    global _lib
    if _lib is None:
        if not library_available():
            raise NiqkiError(-1, f"{LIB_PATH} is missing: run `make -C niqki_b200/csrc` (or __graft_entry__.build()); "
                                 "there is no CPU fallback")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(status: int):
    if status != NQ_OK:
        raise NiqkiError(status, lib().nq_last_error().decode(errors="replace"))
