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


def test_shard_range_and_hits_merge_host_side(capi):
    """nq_shard_range / nq_hits_from_arrays / nq_hits_merge are host code: contiguous gid blocks and
    the merge of per-shard hit lists == one sort by (count, gid) descending (niqki_index.cpp:685)."""
    L = capi.lib()
    for n, R in [(100_000, 8), (10, 3), (7, 8), (0, 2)]:
        covered = []
        for r in range(R):
            b, e = C.c_uint64(), C.c_uint64()
            capi.check(L.nq_shard_range(n, R, r, C.byref(b), C.byref(e)))
            assert b.value <= e.value <= n
            covered += list(range(b.value, e.value)) if n <= 100 else [(b.value, e.value)]
        if n <= 100:
            assert covered == list(range(n))
        else:
            assert covered[0][0] == 0 and covered[-1][1] == n and all(covered[i][1] == covered[i + 1][0] for i in range(R - 1))
    rng = np.random.default_rng(3)
    nq, shards = 17, 3
    parts, allhits = [], [[] for _ in range(nq)]
    handles = (C.c_void_p * shards)()
    for s in range(shards):
        ptr, cnt, gid = [0], [], []
        for q in range(nq):
            k = int(rng.integers(0, 6))
            gs = rng.choice(np.arange(s * 1000, (s + 1) * 1000), size=k, replace=False)
            cs = rng.integers(1, 5, size=k)
            order = np.lexsort((gs, cs))[::-1]
            for i in order:
                cnt.append(int(cs[i])); gid.append(int(gs[i])); allhits[q].append((int(cs[i]), int(gs[i])))
            ptr.append(len(cnt))
        p = np.array(ptr, np.uint64); c = np.array(cnt + [0], np.uint32); g = np.array(gid + [0], np.uint32)
        h = C.c_void_p()
        capi.check(L.nq_hits_from_arrays(p.ctypes.data, c.ctypes.data, g.ctypes.data, nq, C.byref(h)))
        handles[s] = h
    out = C.c_void_p()
    capi.check(L.nq_hits_merge(handles, shards, C.byref(out)))
    total = int(L.nq_hits_total(out))
    mp = np.ctypeslib.as_array(L.nq_hits_ptr(out), shape=(nq + 1,))
    mc = np.ctypeslib.as_array(L.nq_hits_counts(out), shape=(max(total, 1),))
    mg = np.ctypeslib.as_array(L.nq_hits_gids(out), shape=(max(total, 1),))
    for q in range(nq):
        exp = sorted(allhits[q], reverse=True)
        got = list(zip(mc[mp[q]:mp[q + 1]].tolist(), mg[mp[q]:mp[q + 1]].tolist()))
        assert got == exp
    for s in range(shards):
        L.nq_hits_free(handles[s])
    L.nq_hits_free(out)
