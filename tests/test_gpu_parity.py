"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle and the committed golden
fixtures.  Bit-exact everywhere — sketches, posting lists, hit lists, matrix counts."""
import json
import zlib

import numpy as np
import pytest

from tests.util import c1_genomes, load_json, load_npz, random_dna

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def nb():
    import niqki_b200

    return niqki_b200


@pytest.fixture(scope="module")
def ctx(nb):
    c = nb.Context(0)
    yield c
    c.close()


def oracle(**kw):
    from oracle.oracle import Oracle

    return Oracle(**kw)


def gpu_index(nb, ctx, K=31, S=15, W=12, H=4, J=0.0, genome_size=0):
    return nb.Index(S=S, K=K, W=W, H=H, min_fract=J, ctx=ctx, genome_size=genome_size)


def test_params_match_oracle(nb, ctx):
    for ps in [dict(K=31, S=15, W=12, H=4), dict(K=21, S=8, W=12, H=4, genome_size=1000),
               dict(K=31, S=10, W=12, H=4, genome_size=5e6), dict(K=31, S=15, W=12, H=4, J=0.1)]:
        g = gpu_index(nb, ctx, **ps)
        assert g.p.as_dict() == oracle(**ps).p.as_dict()


def test_golden_sketches(nb, ctx):
    """Adversarial sequences x parameter sets, expected values from the reference."""
    z = load_npz("sketches.npz")
    psets = json.loads(str(z["param_sets"]))
    checked = 0
    for pi, ps in enumerate(psets):
        g = gpu_index(nb, ctx, **ps)
        names = [str(n) for n in z["names"] if f"sk_{pi}_{n}" in z.files]
        sks, flags = g.sketch_many([z["seq_" + n] for n in names])
        for i, n in enumerate(names):
            assert np.array_equal(sks[i], z[f"sk_{pi}_{n}"]), (ps, n)
            assert flags[i] == 0
            checked += 1
    assert checked > 100


def test_sketch_random_vs_oracle(nb, ctx):
    rng = np.random.default_rng(7)
    alpha = b"ACGTNacgtRYK-"
    pr = np.array([.22, .22, .22, .22, .03, .02, .02, .02, .01, .005, .005, .005, .005])
    pr /= pr.sum()
    for ps in [dict(K=31, S=8, W=12, H=4), dict(K=21, S=10, W=12, H=4), dict(K=13, S=6, W=8, H=3),
               dict(K=31, S=12, W=12, H=4, genome_size=1000), dict(K=16, S=7, W=10, H=4), dict(K=17, S=7, W=10, H=4)]:
        o = oracle(**ps)
        g = gpu_index(nb, ctx, **ps)
        seqs = []
        while len(seqs) < 60:
            s = rng.choice(np.frombuffer(alpha, np.uint8), size=int(rng.integers(ps["K"] + 1, 3000)), p=pr)
            if o.compute_sketch(s, max_passes=100000)[1] >= 0:  # the reference must terminate on it
                seqs.append(s)
        seqs += [b"ACGT" * 3, b"", b"A" * ps["K"]]  # len <= K: skipped entries
        sks, flags = g.sketch_many(seqs)
        exp = o.sketch_many(seqs)
        assert np.array_equal(sks, exp), ps
        assert list(flags[-3:]) == [1, 1, 1] and not flags[:-3].any()


def test_sketch_long_sliced_and_batched(nb, ctx):
    """Long entries are sliced into spans and merged with atomicMin; must equal the sequential scan."""
    rng = np.random.default_rng(11)
    o = oracle(K=31, S=12, W=12, H=4)
    g = gpu_index(nb, ctx, K=31, S=12, W=12, H=4)
    seqs = [random_dna(rng, 3_000_000), random_dna(rng, 70_001), random_dna(rng, 40), random_dna(rng, 1_000_003, b"ACGTN")]
    sks, flags = g.sketch_many(seqs)
    assert np.array_equal(sks, o.sketch_many(seqs))
    assert not flags.any()


def test_sketch_large_S_global_path(nb, ctx):
    """S > 15: the sketch no longer fits in shared memory."""
    rng = np.random.default_rng(12)
    for S, W, H, K in [(16, 12, 4, 31), (18, 12, 4, 31), (17, 9, 3, 21), (16, 6, 2, 31)]:  # 8-bit / 4-bit coarse filter, generic parameters
        o = oracle(K=K, S=S, W=W, H=H)
        g = gpu_index(nb, ctx, K=K, S=S, W=W, H=H)
        seqs = [random_dna(rng, 2_000_000), random_dna(rng, 300_000)]
        sks, _ = g.sketch_many(seqs)
        assert np.array_equal(sks, o.sketch_many(seqs))


def test_densify_many_passes(nb, ctx):
    """Short read at large S: hundreds of densification passes."""
    rng = np.random.default_rng(13)
    for S in (10, 12, 13):
        o = oracle(K=31, S=S, W=12, H=4)
        g = gpu_index(nb, ctx, K=31, S=S, W=12, H=4)
        reads = [random_dna(rng, 150) for _ in range(6)]
        sks, flags = g.sketch_many(reads)
        assert np.array_equal(sks, o.sketch_many(reads))
        assert not flags.any()


def test_densify_stall_is_flagged(nb, ctx):
    """Entries on which the reference spins forever are flagged instead of hanging the GPU."""
    o = oracle(K=31, S=6, W=8, H=4)
    g = gpu_index(nb, ctx, K=31, S=6, W=8, H=4)
    rng = np.random.default_rng(1)
    stalled = None
    for _ in range(400):
        s = random_dna(rng, 32)
        if o.compute_sketch(s, max_passes=5000)[1] < 0:
            stalled = s
            break
    assert stalled is not None
    sk, flags = g.sketch_many([stalled])
    assert flags[0] & 2
    scan, _ = o.sketch_scan(stalled)
    assert np.array_equal(sk[0][scan != -1], scan[scan != -1])


def test_c1_end_to_end(nb, ctx):
    """Config 1 (nine bundled E. coli genomes, defaults): sketches, postings, hits, matrix."""
    seqs, z = c1_genomes()
    g = gpu_index(nb, ctx)
    sks, flags = g.sketch_many(seqs)
    assert not flags.any()
    assert [zlib.crc32(s.astype("<i4").tobytes()) for s in sks] == list(z["sketch_crc32"])
    g.insert_sketches(sks)
    info = g.info()
    assert info["n_postings"] == 294912 and info["n_genomes"] == 9
    sizes, gids = g.export_postings()
    assert int((sizes > 0).sum()) == 40522 and int(sizes.max()) == 9
    assert zlib.crc32(sizes.astype("<u4").tobytes()) == int(z["sizes_crc32"])
    assert zlib.crc32(gids.astype("<u4").tobytes()) == int(z["postings_crc32"])
    ptr, c, gid = g.query_sketches(sks)
    hm = np.zeros((9, 9), np.uint32)
    for q in range(9):
        seg = slice(int(ptr[q]), int(ptr[q + 1]))
        hm[q, gid[seg]] = c[seg]
        packed = (c[seg].astype(np.uint64) << np.uint64(32)) | gid[seg]
        assert np.all(packed[:-1] > packed[1:])  # (count,gid) strictly descending
    assert np.array_equal(hm, z["hit_matrix"])
    assert np.array_equal(g.query_matrix(), z["hit_matrix"])
    # thresholded query: min_score = (uint32)(0.9*F)
    ptr, c, gid = g.query_sketches(sks, min_score=int(0.9 * 32768))
    for q in range(9):
        exp = np.sort(z["hit_matrix"][q][z["hit_matrix"][q] >= int(0.9 * 32768)])[::-1]
        assert np.array_equal(c[int(ptr[q]):int(ptr[q + 1])], exp)


def test_golden_small_index(nb, ctx):
    z = load_npz("small_index.npz")
    ps = json.loads(str(z["params"]))
    J = ps.pop("J")
    g = gpu_index(nb, ctx, J=J, **ps)
    n = z["sketches"].shape[0]
    sks, _ = g.sketch_many([z[f"entry_{i}"] for i in range(n)])
    assert np.array_equal(sks, z["sketches"])
    g.insert_sketches(sks)
    sizes, gids = g.export_postings()
    assert np.array_equal(sizes, z["sizes"]) and np.array_equal(gids, z["gids"])
    qsk, _ = g.sketch_many([z[f"query_{i}"] for i in range(4)])
    assert np.array_equal(qsk, z["qsketches"])
    ptr, c, gid = g.query_sketches(qsk)
    for q in range(4):
        seg = slice(int(ptr[q]), int(ptr[q + 1]))
        assert np.array_equal(c[seg], z[f"hit_counts_{q}"]) and np.array_equal(gid[seg], z[f"hit_gids_{q}"])
    m = g.query_matrix()
    lines = bytes(z["matrix_text"]).decode().split("\n")
    for q in range(n):
        exp = ["%g" % (int(v) / g.F) if v >= g.min_score else "0" for v in m[q]]
        assert lines[1 + q].split("\t")[1:-1] == exp


def _family_entries(rng, n, length, nfam=5, nmut=30, alphabet=b"ACGT"):
    fam = [random_dna(rng, length, alphabet) for _ in range(nfam)]
    ents = []
    for i in range(n):
        s = fam[i % nfam].copy()
        pos = rng.integers(0, length, nmut)
        s[pos] = random_dna(rng, nmut)
        ents.append(s)
    return ents


@pytest.mark.parametrize("ps,n,J,gid_base", [
    (dict(K=31, S=8, W=8, H=4), 300, 0.1, 0),
    (dict(K=31, S=10, W=12, H=4), 77, 0.0, 1000),
    (dict(K=21, S=6, W=5, H=2), 500, 0.3, 7),
    (dict(K=31, S=9, W=12, H=4, genome_size=1000), 64, 0.05, 0),   # B7: out-of-range fps are not posted
    (dict(K=31, S=16, W=10, H=4), 40, 0.2, 0),                      # u32 counters
])
def test_index_query_matrix_vs_oracle(nb, ctx, ps, n, J, gid_base):
    rng = np.random.default_rng(n)
    o = oracle(J=J, **ps)
    g = gpu_index(nb, ctx, J=J, **ps)
    ents = _family_entries(rng, n, 4000 if ps["S"] < 16 else 200000)
    sks, _ = g.sketch_many(ents)
    assert np.array_equal(sks, o.sketch_many(ents))
    g.insert_sketches(sks, gid_base=gid_base)
    for i, sk in enumerate(sks):
        o.insert_sketch(sk, gid_base + i)
    rp, ogids = o.csr()
    sizes, gids = g.export_postings()
    assert np.array_equal(sizes, np.diff(rp).astype(np.uint32))
    assert np.array_equal(gids, ogids)
    queries = ents[:10] + [random_dna(rng, 4000)]
    qs, _ = g.sketch_many(queries)
    ptr, c, gid = g.query_sketches(qs)
    ohp, oc, og = o.query_batch(qs)
    if gid_base:
        # the oracle's index has gid_base empty genomes in front; with min_score 0 they are reported
        keep = og >= gid_base
        counts_per_q = [int(keep[int(ohp[q]):int(ohp[q + 1])].sum()) for q in range(len(queries))]
        oc, og = oc[keep], og[keep]
        ohp = np.concatenate([[0], np.cumsum(counts_per_q)]).astype(np.uint64)
    assert np.array_equal(ptr, ohp) and np.array_equal(c, oc) and np.array_equal(gid, og)
    if gid_base == 0:
        m = g.query_matrix()
        assert np.array_equal(m, o.matrix_counts().astype(np.uint32))
        assert np.array_equal(g.query_range(3, 9), m[3:9])


def test_matrix_wrap16(nb, ctx):
    """uint16 counters wrap for S >= 16; unwrapped counts only on request."""
    rng = np.random.default_rng(5)
    ps = dict(K=31, S=17, W=4, H=2)
    o = oracle(**ps)
    g = gpu_index(nb, ctx, **ps)
    s = random_dna(rng, 400000)
    sks, _ = g.sketch_many([s, s, random_dna(rng, 400000)])
    g.insert_sketches(sks)
    o.insert_sketches(sks)
    assert np.array_equal(g.query_matrix(wrap16=True), o.matrix_counts().astype(np.uint32))
    raw = g.query_matrix(wrap16=False)
    assert raw[0, 1] == 1 << 17 and g.query_matrix(wrap16=True)[0, 1] == 0


def test_index_import_export_roundtrip(nb, ctx):
    rng = np.random.default_rng(21)
    ps = dict(K=31, S=7, W=8, H=4)
    g = gpu_index(nb, ctx, **ps)
    sks, _ = g.sketch_many(_family_entries(rng, 50, 3000))
    g.insert_sketches(sks)
    sizes, gids = g.export_postings()
    h = gpu_index(nb, ctx, **ps)
    h.import_postings(sizes, gids, 50)
    s2, g2 = h.export_postings()
    assert np.array_equal(sizes, s2) and np.array_equal(gids, g2)
    a = g.query_sketches(sks[:5])
    b = h.query_sketches(sks[:5])
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
    # shard import: keep gids 20..39 only
    h.import_postings(sizes, gids, 20, gid_base=20)
    ptr, c, gid = h.query_sketches(sks[:5])
    assert gid.size and gid.min() >= 20 and gid.max() < 40


def test_sharded_query_equals_single(nb, ctx):
    """§8e: shard by genome id, query every shard, merge = concatenate + sort (count,gid) desc."""
    rng = np.random.default_rng(31)
    ps = dict(K=31, S=9, W=12, H=4)
    J = 0.05
    ents = _family_entries(rng, 120, 5000)
    g = gpu_index(nb, ctx, J=J, **ps)
    sks, _ = g.sketch_many(ents)
    g.insert_sketches(sks)
    ptr, c, gid = g.query_sketches(sks[:16])
    shards = []
    for lo, hi in [(0, 50), (50, 95), (95, 120)]:
        s = gpu_index(nb, ctx, J=J, **ps)
        s.insert_sketches(sks[lo:hi], gid_base=lo)
        shards.append(s.query_sketches(sks[:16]))
    for q in range(16):
        parts = [(cc[int(pp[q]):int(pp[q + 1])].astype(np.uint64) << np.uint64(32)) | gg[int(pp[q]):int(pp[q + 1])]
                 for pp, cc, gg in shards]
        merged = np.sort(np.concatenate(parts))[::-1]
        seg = slice(int(ptr[q]), int(ptr[q + 1]))
        assert np.array_equal(merged, (c[seg].astype(np.uint64) << np.uint64(32)) | gid[seg])


@pytest.mark.parametrize("n", [5_000, 20_000, 30_000, 65_400, 70_000])
def test_index_build_synthetic_sketches(nb, ctx, n):
    """Index build on raw sketches: skewed fingerprints, lists far longer than a warp (clusters of
    identical genomes), unposted cells (-1 and fp >= 2^W), u16 -> u32 id switch at 65400."""
    rng = np.random.default_rng(n)
    ps = dict(K=31, S=5, W=12, H=4)
    F = 32
    o = oracle(J=0.2, **ps)
    g = gpu_index(nb, ctx, J=0.2, **ps)
    sks = (rng.integers(0, 4096, size=(n, F)) & rng.integers(0, 4096, size=(n, F))).astype(np.int32)  # skewed to small values
    clones = rng.random(n) < 0.3
    sks[clones] = sks[0]                          # 30% identical genomes: one list of ~0.3 n per cell
    sks[rng.random((n, F)) < 0.01] = -1
    sks[rng.random((n, F)) < 0.01] = 5000         # out of range: not posted
    g.insert_sketches(sks)
    o.insert_sketches(sks)
    rp, ogids = o.csr()
    sizes, gids = g.export_postings()
    assert np.array_equal(sizes, np.diff(rp).astype(np.uint32))
    assert np.array_equal(gids, ogids)
    q = np.concatenate([sks[:3], sks[n // 2:n // 2 + 3]])
    ptr, c, gid = g.query_sketches(q)
    ohp, oc, og = o.query_batch(q)
    assert np.array_equal(ptr, ohp) and np.array_equal(c, oc) and np.array_equal(gid, og)


@pytest.mark.parametrize("n", [150_000, 2_200_000])
def test_large_n_global_counter_path(nb, ctx, n):
    """More genomes than fit in shared-memory counters: counters move to HBM (2.2M: the entry
    count of a reads-as-entries index, beyond the old 2M-per-build limit)."""
    rng = np.random.default_rng(41)
    ps = dict(K=31, S=4, W=6, H=3)
    o = oracle(J=0.5, **ps)
    g = gpu_index(nb, ctx, J=0.5, **ps)
    sks = rng.integers(0, 64, size=(n, 16)).astype(np.int32)
    sks[rng.random((n, 16)) < 0.01] = -1
    g.insert_sketches(sks)
    o.insert_sketches(sks)
    q = sks[:5]
    ptr, c, gid = g.query_sketches(q)
    ohp, oc, og = o.query_batch(q)
    assert np.array_equal(ptr, ohp) and np.array_equal(c, oc) and np.array_equal(gid, og)


def test_matrix_rows_equal_pairwise_cell_matches(nb, ctx):
    """--matrix rows are computed as dense queries of the genomes' own (index-rebuilt) sketches:
    check them against the definition, counts[a][b] = #cells where a and b post the same fingerprint
    (src/niqki_index.cpp:570-598), on sketches with unposted cells, out-of-range fingerprints and clones."""
    rng = np.random.default_rng(7)
    n, ps = 3000, dict(K=31, S=10, W=12, H=4)
    F = 1024
    g = gpu_index(nb, ctx, J=0.0, **ps)
    sks = (rng.integers(0, 4096, size=(n, F)) & rng.integers(0, 4096, size=(n, F))).astype(np.int32)
    sks[rng.random(n) < 0.05] = sks[1]
    sks[rng.random((n, F)) < 0.02] = -1
    sks[rng.random((n, F)) < 0.01] = 4096 + 17
    g.insert_sketches(sks, gid_base=0)
    valid = (sks >= 0) & (sks < 4096)
    for a, b in [(0, 40), (1490, 1530), (n - 7, n)]:
        m = g.query_range(a, b, wrap16=False)
        exp = np.stack([((sks == sks[r]) & valid & valid[r]).sum(axis=1) for r in range(a, b)]).astype(np.uint32)
        assert np.array_equal(m, exp)
    assert g.query_range(5, 5).shape[0] == 0


def test_posting_index_beyond_32_bits(nb, ctx):
    """S=18 with > 16384 genomes: F * gid_stride exceeds 2^32, the kernels switch to 64-bit posting
    indexes.  Sketches are generated on the device; expected counts come from torch on the same tensors."""
    import torch

    n, S = 16_500, 18
    F = 1 << S
    g = gpu_index(nb, ctx, K=31, S=S, W=12, H=4, J=0.0)
    gen = torch.Generator(device="cuda").manual_seed(5)
    sk = torch.randint(0, 48, (n, F), dtype=torch.int32, device="cuda", generator=gen)
    sk[:, ::97] = -1
    sk[3] = sk[n - 1]
    g.insert_sketches(sk, gid_base=0)
    assert g.info()["n_postings"] == int((sk >= 0).sum())
    rows = [0, 3, n - 1]
    q = sk[rows].contiguous()
    ptr, c, gid = g.query_sketches(q, min_score=0)
    for i, r in enumerate(rows):
        exp = ((sk == q[i]) & (q[i] >= 0)).sum(dim=1).cpu().numpy().astype(np.uint32)
        seg = slice(int(ptr[i]), int(ptr[i + 1]))
        got = np.zeros(n, np.uint32)
        got[gid[seg]] = c[seg]
        assert np.array_equal(got, exp)
        m = g.query_range(r, r + 1, wrap16=True)[0]
        assert np.array_equal(m, exp & 0xFFFF)
    assert c[int(ptr[1])] == int((q[1] >= 0).sum()) and set(gid[int(ptr[1]):int(ptr[1]) + 2]) == {3, n - 1}
    del sk


def test_device_pointer_api_and_synth(nb, ctx):
    """Device-resident path (torch tensors as plain device pointers) + the synthetic generators."""
    import ctypes as C

    import torch

    from niqki_b200.capi import check, lib

    o = oracle(K=31, S=12, W=12, H=4, J=0.1)
    g = gpu_index(nb, ctx, K=31, S=12, W=12, H=4, J=0.1)
    L = lib()
    n, length = 6, 200_000
    d = torch.empty(n * length + 64, dtype=torch.uint8, device="cuda")
    check(L.nq_synth_genomes_device(ctx.h, 42, 3, n, length, C.c_void_p(d.data_ptr())))
    ctx.sync()
    host = d.cpu().numpy()
    for i in range(n):
        assert np.array_equal(host[i * length:(i + 1) * length], o.synth_genome(3 + i, length))
    gs = np.arange(3, 3 + n, dtype=np.uint64)
    qs = np.arange(100, 100 + n, dtype=np.uint64)
    rates = [0.001, 0.01, 0.05, 0.0, 0.5, 0.01]
    thr = np.array([int(np.ldexp(np.longdouble(r), 64)) if r > 0 else 0 for r in rates], dtype=np.uint64)
    dm = torch.empty(n * length + 64, dtype=torch.uint8, device="cuda")
    check(L.nq_synth_mutants_device(ctx.h, 42, gs.ctypes.data, qs.ctypes.data, thr.ctypes.data, n, length,
                                    C.c_void_p(dm.data_ptr())))
    ctx.sync()
    hm = dm.cpu().numpy()
    for i in range(n):
        assert np.array_equal(hm[i * length:(i + 1) * length], o.synth_mutant(3 + i, 100 + i, rates[i], length))
    dr = torch.empty(1000 * 150 + 64, dtype=torch.uint8, device="cuda")
    check(L.nq_synth_reads_device(ctx.h, 42, 5000, 1000, 5_000_000, 150, C.c_void_p(dr.data_ptr())))
    ctx.sync()
    hr = dr.cpu().numpy()
    for i in (0, 1, 999):
        assert np.array_equal(hr[i * 150:(i + 1) * 150], o.synth_read(5000 + i, 5_000_000))
    # sketch / index / query entirely on device pointers
    offs = np.arange(n + 1, dtype=np.uint64) * length
    sk, flags = g.compute_sketches(d, offs)
    ctx.sync()
    exp = o.sketch_batch(host[: n * length].copy(), offs)
    assert np.array_equal(sk.cpu().numpy(), exp)
    g.insert_sketches(sk)
    qsk, _ = g.compute_sketches(dm, offs)
    ptr, c, gid = g.query_sketches(qsk)
    o.insert_sketches(exp)
    ohp, oc, og = o.query_batch(qsk.cpu().numpy())
    assert np.array_equal(ptr, ohp) and np.array_equal(c, oc) and np.array_equal(gid, og)
    assert g.query_sketches(qsk, fetch=False) is None


def test_reads_lines_mode_small(nb, ctx):
    """Config-4 shape in miniature: many 150 bp entries, S=8 (densification on every entry)."""
    o = oracle(K=31, S=8, W=12, H=4, J=0.2)
    g = gpu_index(nb, ctx, K=31, S=8, W=12, H=4, J=0.2)
    reads = [o.synth_read(r, 100_000) for r in range(3000)]
    sks, flags = g.sketch_many(reads)
    assert np.array_equal(sks, o.sketch_many(reads)) and not flags.any()
    g.insert_sketches(sks)
    o.insert_sketches(sks)
    ptr, c, gid = g.query_sketches(sks[:20])
    ohp, oc, og = o.query_batch(sks[:20])
    assert np.array_equal(ptr, ohp) and np.array_equal(c, oc) and np.array_equal(gid, og)


@pytest.mark.parametrize("ps", [dict(K=31, S=8, W=12, H=4), dict(K=21, S=4, W=10, H=3), dict(K=27, S=10, W=12, H=4),
                                dict(K=15, S=11, W=8, H=2)])
def test_reads_kernel_edge_cases(nb, ctx, ps):
    """Warp-per-entry fused sketch + densification (short entries, S <= 11): ragged lengths around K,
    empty entries, lower case / N / foreign bytes inside and outside the K-1 seed characters (B3, B4),
    entries whose densification never ends (flagged, scan part still exact)."""
    rng = np.random.default_rng(ps["K"] * 100 + ps["S"])
    o = oracle(**ps)
    g = gpu_index(nb, ctx, **ps)
    K = ps["K"]
    dna = lambda n: random_dna(rng, n).tobytes()  # noqa: E731
    seqs = [b"", b"A", dna(K - 1), dna(K), dna(K + 1), dna(K + 2)]
    for L in (40, 64, 65, 100, 150, 151, 250, 1000, 4000):
        for variant in range(4):
            s = bytearray(dna(L))
            if variant == 1:   # lower case in the seed and beyond
                for pos in rng.integers(0, L, size=max(1, L // 10)):
                    s[pos] = ord(chr(s[pos]).lower())
            elif variant == 2:  # a foreign byte inside the seed
                s[int(rng.integers(0, K - 1))] = ord("N")
            elif variant == 3:  # foreign bytes after the seed
                for pos in rng.integers(K - 1, L, size=max(1, L // 20)):
                    s[pos] = ord("N") if pos % 2 else ord("x")
            seqs.append(bytes(s))
    sks, flags = g.sketch_many(seqs)
    flags = np.asarray(flags)
    for i, sq in enumerate(seqs):
        exp, passes = o.compute_sketch(sq, max_passes=20000)
        scan, filled = o.sketch_scan(sq)
        if len(sq) <= K:
            assert flags[i] & 1 and (sks[i] == -1).all(), i
        elif passes < 0:  # the reference would spin forever
            assert flags[i] & 2, i
            assert np.array_equal(sks[i][scan != -1], scan[scan != -1]), i
        else:
            assert flags[i] == 0 and np.array_equal(sks[i], exp), (i, len(sq))


def test_full_size_properties(nb, ctx):
    """BASELINE-sized entries (5 Mbp, defaults), checked through size-independent properties:
    sketching is deterministic, a genome's best hit is itself with count F, mutated copies rank
    their parent first, and sharding commutes with the merge."""
    import ctypes as C

    import torch

    from niqki_b200.capi import check, lib

    L = lib()
    g = gpu_index(nb, ctx, J=0.1)
    n, length = 24, 5_000_000
    d = torch.empty(n * length + 64, dtype=torch.uint8, device="cuda")
    check(L.nq_synth_genomes_device(ctx.h, 42, 0, n, length, C.c_void_p(d.data_ptr())))
    offs = np.arange(n + 1, dtype=np.uint64) * length
    sk, flags = g.compute_sketches(d, offs)
    sk2, _ = g.compute_sketches(d, offs)
    assert torch.equal(sk, sk2) and not flags.cpu().numpy().any()
    assert int((sk < 0).sum()) == 0
    # genome 0 against the CPU oracle (one full-size entry keeps the CPU time in seconds)
    o = oracle()
    assert np.array_equal(sk[0].cpu().numpy(), o.compute_sketch(o.synth_genome(0, length))[0])
    g.insert_sketches(sk)
    nq = 6
    gs = np.arange(nq, dtype=np.uint64)
    qs = np.arange(nq, dtype=np.uint64)
    rates = [0.001, 0.01, 0.05] * 2
    thr = np.array([int(np.ldexp(np.longdouble(r), 64)) for r in rates], dtype=np.uint64)
    dm = torch.empty(nq * length + 64, dtype=torch.uint8, device="cuda")
    check(L.nq_synth_mutants_device(ctx.h, 42, gs.ctypes.data, qs.ctypes.data, thr.ctypes.data, nq, length,
                                    C.c_void_p(dm.data_ptr())))
    qsk, _ = g.compute_sketches(dm, offs[: nq + 1])
    ptr, c, gid = g.query_sketches(qsk)
    for q in range(nq):
        seg = slice(int(ptr[q]), int(ptr[q + 1]))
        assert gid[seg][0] == q and c[seg][0] >= 3276  # parent first, above (uint32)(0.1*32768)
        assert len(gid[seg]) == 1  # unrelated genomes share ~22 cells, far below the threshold
    ptr, c, gid = g.query_sketches(sk[:4])
    assert list(c) == [32768] * 4 and list(gid) == [0, 1, 2, 3]
    m = g.query_matrix()
    assert np.array_equal(np.diag(m), np.full(n, 32768, np.uint32)) and np.array_equal(m, m.T)
    assert 5 <= np.median(m[~np.eye(n, dtype=bool)]) <= 45  # background collisions ~22


# ---- kernel forms of the headline configurations at realistic list statistics --------------------
def _fp_cdf(lam=152.6, W=12, H=4):
    """Distribution of a sketch cell of a random genome: min over ~Poisson(lam) k-mers of the W-bit
    fingerprint (HLL part on top of W-H low hash bits) — what the posting lists of configs[1]/[2]
    look like without sketching 50 Gbases in a test."""
    M = W - H
    top = (1 << H) - 1
    pm = np.zeros(1 << W)
    for h in range(top + 1):
        ph = 2.0 ** -(top + 1 - h) if h > 0 else 2.0 ** -top
        pm[h << M:(h + 1) << M] = ph / (1 << M)
    cdf = np.cumsum(pm)
    surv = np.exp(-lam * np.concatenate([[0.0], cdf]))
    p = surv[:-1] - surv[1:]
    return np.cumsum(p / p.sum())


def _realistic_sketches(rng, n, F, cdf):
    table = np.minimum(np.searchsorted(cdf, (np.arange(65536) + 0.5) / 65536.0, side="right"), len(cdf) - 1).astype(np.int32)
    out = np.empty((n, F), np.int32)
    for r0 in range(0, n, 4096):
        r1 = min(n, r0 + 4096)
        out[r0:r1] = table[rng.integers(0, 65536, size=(r1 - r0, F), dtype=np.uint16)]
    return out


def _queries_like(rng, sks, nq, cdf):
    """Mutated copies (10 %, 40 %, 90 % of the cells redrawn) of random genomes + unrelated sketches."""
    n, F = sks.shape
    q = _realistic_sketches(rng, nq, F, cdf)
    for i in range(nq - 4):
        keep = rng.random(F) >= (0.1, 0.4, 0.9)[i % 3]
        parent = sks[rng.integers(0, n)]
        q[i, keep] = parent[keep]
    q[-1, ::7] = -1       # empty cells and out-of-range fingerprints are not probed (:655)
    q[-2, ::5] = 1 << 20
    return q


@pytest.mark.parametrize("n,S,J,what", [
    (10_000, 13, 0.1, "slab G=8, 128-thread CTAs (configs[1] shard)"),
    (12_500, 12, 0.1, "slab G=8, 256-thread CTAs (configs[2] shard)"),
    (25_000, 11, 0.1, "slab G=16 (100k index on 4 GPUs)"),
    (50_000, 10, 0.1, "slab G=32/64 (100k index on 2 GPUs)"),
    (65_400, 10, 0.0, "largest u16 shard, every genome reported"),
    (72_000, 10, 0.1, "split16 segment-table form, 1024 threads (100k index on 1 GPU)"),
    (140_000, 9, 0.1, "u32 ids beyond split16"),
])
def test_query_forms_at_realistic_list_lengths(nb, ctx, n, S, J, what):
    """Postings and hit lists of >= 32 queries against the oracle, at the shard sizes of the headline
    configurations and with their list-length statistics (many groups of 32 cells per warp)."""
    rng = np.random.default_rng(n + S)
    ps = dict(K=31, S=S, W=12, H=4)
    F = 1 << S
    cdf = _fp_cdf()
    o = oracle(J=J, **ps)
    g = gpu_index(nb, ctx, J=J, **ps)
    sks = _realistic_sketches(rng, n, F, cdf)
    sks[1::997] = sks[0]                      # a cluster of identical genomes: lists longer than three granules
    sks[rng.random((n, F)) < 0.002] = -1
    g.insert_sketches(sks)
    o.insert_sketches(sks)
    rp, ogids = o.csr()
    sizes, gids = g.export_postings()
    assert np.array_equal(sizes, np.diff(rp).astype(np.uint32)), what
    assert np.array_equal(gids, ogids), what
    q = _queries_like(rng, sks, 40, cdf)
    ptr, c, gid = g.query_sketches(q)
    ohp, oc, og = o.query_batch(q)
    assert np.array_equal(ptr, ohp) and np.array_equal(c, oc) and np.array_equal(gid, og), what
    assert ptr[-1] >= 24, "the mutated copies must report hits"
    # device-sketch entry point, many queries (more than one resident wave for the small forms)
    import torch
    reps = 64 if n <= 25_000 else 8
    dq = torch.from_numpy(np.tile(q, (reps, 1))).cuda()
    ptr2, c2, gid2 = g.query_sketches(dq)
    ctx.sync()
    for r in (0, reps // 2, reps - 1):
        lo, hi = ptr2[r * 40], ptr2[(r + 1) * 40]
        assert np.array_equal(ptr2[r * 40:(r + 1) * 40 + 1] - lo, ohp), what
        assert np.array_equal(c2[lo:hi], oc) and np.array_equal(gid2[lo:hi], og), what


# ---- K1: the 2-bit packed wire format of the host-buffer sketch calls ----------------------------
@pytest.fixture()
def packed_ctx(nb):
    c = nb.Context(0)
    c.set_host_packing(1, 3)   # every entry travels packed, three packer threads
    yield c
    c.close()


def test_packed_path_golden_sketches(nb, packed_ctx):
    """Every adversarial case of sketches.npz (lower case, N runs, foreign bytes in and out of the
    seed, len around K, -G masks, K in {11, 21, 31}) through pack.cpp + sketch_scan_packed_kernel."""
    z = load_npz("sketches.npz")
    psets = json.loads(str(z["param_sets"]))
    checked = 0
    for pi, ps in enumerate(psets):
        g = gpu_index(nb, packed_ctx, **ps)
        names = [str(n) for n in z["names"] if f"sk_{pi}_{n}" in z.files]
        sks, flags = g.sketch_many([z["seq_" + n] for n in names])
        for i, n in enumerate(names):
            assert np.array_equal(sks[i], z[f"sk_{pi}_{n}"]), (ps, n)
            checked += 1
    assert checked > 100


def test_packed_path_random_and_long_vs_oracle(nb, packed_ctx):
    rng = np.random.default_rng(17)
    alpha = b"ACGTNacgtRYK-"
    pr = np.array([.22, .22, .22, .22, .03, .02, .02, .02, .01, .005, .005, .005, .005])
    pr /= pr.sum()
    for ps in [dict(K=31, S=8, W=12, H=4), dict(K=21, S=10, W=12, H=4), dict(K=13, S=6, W=8, H=3),
               dict(K=31, S=12, W=12, H=4, genome_size=1000), dict(K=16, S=7, W=10, H=4), dict(K=17, S=7, W=10, H=4),
               dict(K=11, S=4, W=6, H=2), dict(K=31, S=11, W=12, H=4)]:
        o = oracle(**ps)
        g = gpu_index(nb, packed_ctx, **ps)
        seqs = []
        while len(seqs) < 40:
            s = rng.choice(np.frombuffer(alpha, np.uint8), size=int(rng.integers(ps["K"] + 1, 3000)), p=pr)
            if o.compute_sketch(s, max_passes=100000)[1] >= 0:
                seqs.append(s)
        seqs += [b"ACGT" * 2, b"", b"A" * ps["K"]]  # len <= K (all K here are >= 11): skipped entries
        sks, flags = g.sketch_many(seqs)
        assert np.array_equal(sks, o.sketch_many(seqs)), ps
        assert list(flags[-3:]) == [1, 1, 1] and not flags[:-3].any()
    # long entries: spans cut at 1024-base boundaries, 1024-thread CTAs, word-aligned thread runs
    o = oracle(K=31, S=12, W=12, H=4)
    g = gpu_index(nb, packed_ctx, K=31, S=12, W=12, H=4)
    seqs = [random_dna(rng, 3_000_000), random_dna(rng, 70_001), random_dna(rng, 40), random_dna(rng, 1_000_003, b"ACGTN"),
            random_dna(rng, 500_000, b"ACGTacgtN")]
    sks, flags = g.sketch_many(seqs)
    assert np.array_equal(sks, o.sketch_many(seqs))
    # S > 15: global sketch behind the coarse filter, generic K
    for S, W, H, K in [(16, 12, 4, 31), (17, 9, 3, 21)]:
        o = oracle(K=K, S=S, W=W, H=H)
        g = gpu_index(nb, packed_ctx, K=K, S=S, W=W, H=H)
        seqs = [random_dna(rng, 1_500_000), random_dna(rng, 300_000, b"ACGTN")]
        sks, _ = g.sketch_many(seqs)
        assert np.array_equal(sks, o.sketch_many(seqs))


def test_packed_and_character_paths_agree_on_a_multi_batch_call(nb, ctx, packed_ctx):
    """More than one 512 MB batch through both host paths (double-buffered staging), equal sketches."""
    rng = np.random.default_rng(23)
    g0 = gpu_index(nb, ctx, K=31, S=10, W=12, H=4)
    g1 = gpu_index(nb, packed_ctx, K=31, S=10, W=12, H=4)
    one = random_dna(rng, 4_000_000, b"ACGTN")
    seqs = [np.roll(np.frombuffer(one, np.uint8), 1000 * i) for i in range(300)]   # 1.2 GB
    ctx.set_host_packing(0, 0)
    a, _ = g0.sketch_many(seqs)
    ctx.set_host_packing(-1, 0)
    b, _ = g1.sketch_many(seqs)
    assert np.array_equal(a, b)
    o = oracle(K=31, S=10, W=12, H=4)
    assert np.array_equal(a[:2], o.sketch_many(seqs[:2]))


@pytest.mark.parametrize("W,H,n", [(13, 4, 3000), (15, 5, 3000), (14, 4, 70_000), (15, 4, 200)])
def test_index_and_query_with_wide_fingerprints(nb, ctx, W, H, n):
    """W = 13..15: 2^W counters of a cell no longer fit eight warps' worth of shared memory; the build
    kernels size themselves.  Postings and hits against the oracle."""
    rng = np.random.default_rng(W * 1000 + n)
    ps = dict(K=31, S=6, W=W, H=H)
    F = 64
    o = oracle(J=0.1, **ps)
    g = gpu_index(nb, ctx, J=0.1, **ps)
    sks = rng.integers(0, 1 << W, size=(n, F)).astype(np.int32)
    dup = rng.random((n, F)) < 0.3
    sks[dup] = np.broadcast_to(sks[0], (n, F))[dup]     # shared fingerprints: lists of ~0.3 n
    sks[rng.random((n, F)) < 0.01] = -1
    g.insert_sketches(sks)
    o.insert_sketches(sks)
    rp, ogids = o.csr()
    sizes, gids = g.export_postings()
    assert np.array_equal(sizes, np.diff(rp).astype(np.uint32))
    assert np.array_equal(gids, ogids)
    q = np.concatenate([sks[:4], rng.integers(0, 1 << W, size=(4, F)).astype(np.int32)])
    ptr, c, gid = g.query_sketches(q)
    ohp, oc, og = o.query_batch(q)
    assert np.array_equal(ptr, ohp) and np.array_equal(c, oc) and np.array_equal(gid, og)


def test_import_of_unsorted_lists(nb, ctx):
    """The reference's own dumps hold a list in the order its OpenMP threads appended.
    Import must not depend on it: shuffled lists, more than 65600 genomes (split16 directory)."""
    rng = np.random.default_rng(99)
    n, F = 66_000, 32
    ps = dict(K=31, S=5, W=12, H=4)
    o = oracle(J=0.2, **ps)
    g = gpu_index(nb, ctx, J=0.2, **ps)
    sks = (rng.integers(0, 4096, size=(n, F)) & rng.integers(0, 4096, size=(n, F))).astype(np.int32)
    sks[rng.random(n) < 0.2] = sks[1]
    o.insert_sketches(sks)
    rp, ogids = o.csr()
    sizes = np.diff(rp).astype(np.uint32)
    shuffled = ogids.copy()
    for l in np.flatnonzero(sizes > 1)[:200000]:
        seg = shuffled[int(rp[l]):int(rp[l + 1])]
        rng.shuffle(seg)
    g.import_postings(sizes, shuffled, n)
    q = np.concatenate([sks[:3], sks[n - 3:]])
    ptr, c, gid = g.query_sketches(q)
    ohp, oc, og = o.query_batch(q)
    assert np.array_equal(ptr, ohp) and np.array_equal(c, oc) and np.array_equal(gid, og)
    s2, g2 = g.export_postings()
    assert np.array_equal(s2, sizes) and np.array_equal(g2, ogids)   # lists come back gid-ascending


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_query_over_processes_equals_single_gpu(tmp_path, world):
    """One process per GPU under torch.distributed.run: index sharded by genome id, query sketches
    all-gathered by nq_allgather_sketches (NCCL behind the C ABI), hit lists merged by nq_hits_merge —
    the merged (ptr, counts, gids) must equal the single-GPU answer.  Skipped with fewer GPUs."""
    import os
    import socket
    import subprocess
    import sys

    import torch

    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} CUDA devices")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    out = tmp_path / "result.txt"
    cfg = dict(n=3001, S=10, nq=37, J=0.05, seed=world, out=str(out))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr",
                        "127.0.0.1", "--master-port", str(port), os.path.join(root, "tests", "_gpu_shard_worker.py"), json.dumps(cfg)],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert out.read_text() == "OK"
