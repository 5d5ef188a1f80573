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


def test_packer_wire_format_matches_the_character_rules(capi):
    """K1 host half (pack.cpp): codes / `other` masks / seed rewriting against a plain numpy statement
    of nuc2int, nuc2intrc and str2numstrand (niqki_index.cpp:114-123, 211-221, 255-273) — random
    records with lower case, N runs and foreign bytes inside and outside the seeds, record
    boundaries at every alignment, AVX2 and scalar tails."""
    L = capi.lib()
    rng = np.random.default_rng(11)
    alphabet = np.frombuffer(b"ACGTACGTACGTACGTacgtNnRY*-\x00\xff", np.uint8)
    for K in (31, 21, 5):
        lens = [0, 1, K - 2, K - 1, K, K + 1, 47, 512, 513, 1000, 4097, 65536 + 13, 300_001]
        recs = []
        for i, n in enumerate(lens):
            r = rng.choice(alphabet[:4], size=n)
            dirty = rng.random(n) < (0.0, 0.02, 0.3)[i % 3]
            r[dirty] = rng.choice(alphabet, size=int(dirty.sum()))
            recs.append(r.astype(np.uint8))
        bases = np.concatenate(recs)
        offs = np.zeros(len(recs) + 1, np.uint64)
        offs[1:] = np.cumsum([len(r) for r in recs])
        words, blocks = C.c_uint64(), C.c_uint64()
        capi.check(L.nq_pack_sizes(int(offs[-1]), C.byref(words), C.byref(blocks)))
        codes = np.full(words.value, 0xDEADBEEF, np.uint32)
        blk = np.full(blocks.value, 7, np.uint32)
        pool = np.full(blocks.value * 32, 0xABCD, np.uint16)
        used = C.c_uint64()
        for threads in (1, 5):
            capi.check(L.nq_pack_sequences(bases.ctypes.data, offs.ctypes.data, len(recs), K, codes.ctypes.data, blk.ctypes.data,
                                           pool.ctypes.data, blocks.value, C.byref(used), threads))
            # expected per-base forward code / other flag with the seed rule applied
            fw = np.zeros(256, np.uint8); fw[ord("C")] = 1; fw[ord("G")] = 2; fw[ord("T")] = 3
            ok = np.zeros(256, bool); ok[[ord(c) for c in "ACGT"]] = True
            seedc = np.full(256, 4, np.uint8)
            for j, c in enumerate("ACGT"):
                seedc[ord(c)] = j; seedc[ord(c.lower())] = j
            exp_code = fw[bases].copy(); exp_other = ~ok[bases]
            for r in range(len(recs)):
                e0, n = int(offs[r]), len(recs[r])
                ns = min(n, K - 1)
                sc = seedc[bases[e0:e0 + ns]]
                exp_code[e0:e0 + ns] = sc if (sc < 4).all() else 0
                exp_other[e0:e0 + ns] = False
            lead = 512
            total = words.value * 16
            stream_code = np.zeros(total, np.uint8); stream_other = np.zeros(total, bool)
            stream_code[lead:lead + len(bases)] = exp_code; stream_other[lead:lead + len(bases)] = exp_other
            got_code = ((codes[:, None] >> (2 * np.arange(16, dtype=np.uint32))[None, :]) & 3).astype(np.uint8).ravel()
            assert np.array_equal(got_code, stream_code), (K, threads)
            exp_words = (stream_other.reshape(-1, 16) * (1 << np.arange(16))[None, :]).sum(1).astype(np.uint16)
            for b in range(blocks.value):
                w = exp_words[b * 32:(b + 1) * 32]
                if blk[b] == 0xFFFFFFFF:
                    assert not w.any(), (K, b)
                else:
                    assert blk[b] < used.value
                    slot = pool[blk[b] * 32:blk[b] * 32 + 32]
                    assert np.array_equal(slot[:len(w)], w), (K, b)
