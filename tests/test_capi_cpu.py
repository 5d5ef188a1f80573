"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/niqki_b200.h declares, does its host-side scalar work, and fails loudly without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from oracle.oracle import Oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def capi():
    from niqki_b200 import capi as m

    if not m.library_available():
        m.build_library()
    return m


def test_header_symbols_all_exported(capi):
    hdr = open(os.path.join(ROOT, "include", "niqki_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(nq_[a-z0-9_A-Z]+)\s*\(", hdr))
    assert len(declared) >= 30
    L = C.CDLL(capi.LIB_PATH)
    missing = [s for s in sorted(declared) if not hasattr(L, s)]
    assert not missing, missing
    assert declared == set(capi.SYMBOLS), declared ^ set(capi.SYMBOLS)


def test_params_match_reference_semantics(capi):
    L = capi.lib()
    for K, S, W, H, J, G in [(31, 15, 12, 4, 0.1, 0), (31, 18, 12, 4, 0.1, 0), (21, 8, 12, 4, 0.3, 1000),
                             (31, 10, 12, 4, 0.0, 5e6), (11, 7, 8, 2, 0.999, 0), (31, 9, 16, 6, 0.05, 123456)]:
        p = capi.Params()
        capi.check(L.nq_params_init(C.byref(p), K, S, W, H, J))
        if G:
            capi.check(L.nq_params_select_best_H(C.byref(p), float(G)))
        assert p.as_dict() == Oracle(K=K, S=S, W=W, H=H, J=J, genome_size=G).p.as_dict()


def test_parameter_limits(capi):
    L = capi.lib()
    p = capi.Params()
    assert L.nq_params_init(C.byref(p), 32, 15, 12, 4, 0.0) == capi.NQ_ERR_INVALID  # K <= 31 (B14)
    assert L.nq_params_init(C.byref(p), 31, 20, 12, 4, 0.0) == capi.NQ_ERR_INVALID  # W+S < 32
    assert b"W+S" in L.nq_last_error()


def test_no_silent_cpu_fallback(capi):
    """Without a CUDA device the product path must refuse to run (the oracle is never a fallback)."""
    L = capi.lib()
    n = C.c_int(-1)
    st = L.nq_device_count(C.byref(n))
    if st == capi.NQ_OK and n.value > 0:
        pytest.skip("a GPU is present")
    h = C.c_void_p()
    assert L.nq_ctx_create(0, None, C.byref(h)) == capi.NQ_ERR_CUDA
    import niqki_b200

    with pytest.raises(niqki_b200.NiqkiError):
        niqki_b200.Index(S=8, K=31, W=8, H=4)


def test_product_never_imports_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py may touch oracle/."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "niqki_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")) or f == "Makefile":
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle" not in txt.replace("SURVEY", "").lower() or f in ("synth.cu", "__init__.py"), f
