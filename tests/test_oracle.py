"""Pins the CPU oracle (oracle/niqki_oracle.c) to the reference: SURVEY.md App. C known answers,
the committed fixtures generated from the compiled reference, and oracle/_ref itself when present."""
import json
import zlib

import numpy as np
import pytest

from oracle.oracle import Oracle, Ref, ref_available
from tests.util import c1_genomes, load_json, load_npz, random_dna

needs_ref = pytest.mark.skipif(not ref_available(), reason="oracle/_ref not built (reference tree not mounted)")


def test_appendix_c1_scalars():
    o = Oracle(K=31, S=6, W=8, H=4)
    assert o.revhash64(1) == 0x4179B061E0C0E0D0
    assert o.unrevhash64(1) == 0xB471E5C8635F305A
    assert o.revhash64(0x0123456789ABCDEF) == 0x3EFDEA49C590F4EC
    assert o.unrevhash64(0x0123456789ABCDEF) == 0xC25314746B222512
    assert o.unrevhash64(o.revhash64(42)) == 42
    assert o.hash_family(7, 3) == 0xB4EB977EAFB069A6
    assert o.revhash64(0) == 0 and o.unrevhash64(0) == 0
    for h, fp in [(0x8000000000000000, 240), (0x4000000000000001, 225), (0x00010000000000FF, 15), (0xAB, 11),
                  (1, 1), (0, 0)]:
        assert o.get_fingerprint(h) == fp


def test_appendix_c4_104mer():
    o = Oracle(K=31, S=6, W=8, H=4)
    z = load_npz("sketches.npz")
    sk, _ = o.compute_sketch(z["seq_c4_104mer"])
    exp = ("229 245 223 196 244 254 198 233 238 223 199 254 243 245 208 211 199 245 252 184 243 253 240 251 228 "
           "196 227 238 240 238 245 245 199 247 237 233 236 214 181 229 214 253 247 225 248 248 208 240 229 213 "
           "247 211 246 194 242 253 243 236 189 228 248 244 240 214")
    assert list(sk) == [int(x) for x in exp.split()]


def test_golden_scalars():
    g = load_json("scalars.json")
    o = Oracle()
    for c in g["hash"]:
        assert o.revhash64(c["x"]) == c["rev"]
        assert o.unrevhash64(c["x"]) == c["unrev"]
        assert o.hash_family(c["x"], 3) == c["fam3"]
        assert o.hash_family(c["x"], 1000) == c["fam1000"]
    for c in g["seed"]:
        assert o.str2numstrand(c["s"].encode()) == c["f0"]
        assert o.rcb(c["f0"]) == c["r0"]
    for blk in g["fingerprint"]:
        o = Oracle(**blk["params"])
        assert o.p.as_dict() == blk["derived"]  # includes the stale -G masks (B7)
        for h, fp in blk["cases"]:
            assert o.get_fingerprint(h) == fp
    for c in g["params"]:
        assert Oracle(K=31, S=c["S"], W=4, H=2, J=c["J"]).p.min_score == c["min_score"]


def test_golden_sketches():
    z = load_npz("sketches.npz")
    psets = json.loads(str(z["param_sets"]))
    n = 0
    for pi, ps in enumerate(psets):
        o = Oracle(**ps)
        for name in z["names"]:
            key = f"sk_{pi}_{name}"
            if key not in z.files:
                continue
            sk, passes = o.compute_sketch(z["seq_" + str(name)], max_passes=10**6)
            assert passes >= 0
            assert np.array_equal(sk, z[key]), (ps, name)
            n += 1
    assert n > 100


def test_golden_small_index():
    z = load_npz("small_index.npz")
    ps = json.loads(str(z["params"]))
    o = Oracle(**ps)
    n = z["sketches"].shape[0]
    sks = np.stack([o.compute_sketch(z[f"entry_{i}"])[0] for i in range(n)])
    assert np.array_equal(sks, z["sketches"])
    o.insert_sketches(sks)
    rp, gids = o.csr()
    assert np.array_equal(np.diff(rp).astype(np.uint32), z["sizes"])
    assert np.array_equal(gids, z["gids"])
    for qi in range(z["qsketches"].shape[0]):
        qs, _ = o.compute_sketch(z[f"query_{qi}"])
        assert np.array_equal(qs, z["qsketches"][qi])
        c, g = o.query_sketch(qs)
        assert np.array_equal(c, z[f"hit_counts_{qi}"]) and np.array_equal(g, z[f"hit_gids_{qi}"])
    hp, c, g = o.query_batch(z["qsketches"])
    assert np.array_equal(c, np.concatenate([z[f"hit_counts_{q}"] for q in range(4)]))
    assert np.array_equal(g, np.concatenate([z[f"hit_gids_{q}"] for q in range(4)]))
    # matrix text: rows of count/F with default ostream formatting == '%g'
    m = o.matrix_counts()
    lines = bytes(z["matrix_text"]).decode().split("\n")
    assert lines[0].startswith("##Names\t")
    F = o.F
    for q in range(n):
        cells = lines[1 + q].split("\t")
        assert cells[0] == f"entry{q}"
        exp = ["%g" % (int(v) / F) if v >= o.p.min_score else "0" for v in m[q]]
        assert cells[1:-1] == exp


def test_c1_sketches_and_hits():
    """Config 1 at defaults: sketch CRCs (App. C2), index stats, integer hit matrix (App. C3)."""
    seqs, z = c1_genomes()
    o = Oracle(K=31, S=15, W=12, H=4)
    sks = o.sketch_many(seqs)
    assert [zlib.crc32(s.astype("<i4").tobytes()) for s in sks] == list(z["sketch_crc32"])
    assert "%08x" % z["sketch_crc32"][0] == "efd8d480" and "%08x" % z["sketch_crc32"][8] == "baa5136c"
    assert list(sks[0][:4]) == [1895, 2012, 2088, 1521] and int(sks[0].sum()) == 67034039
    o.insert_sketches(sks)
    rp, gids = o.csr()
    sizes = np.diff(rp)
    assert gids.size == 294912 and int((sizes > 0).sum()) == 40522 and int(sizes.max()) == 9
    assert zlib.crc32(gids.astype("<u4").tobytes()) == int(z["postings_crc32"])
    assert zlib.crc32(sizes.astype("<u4").tobytes()) == int(z["sizes_crc32"])
    hm = np.zeros((9, 9), np.uint32)
    for q in range(9):
        c, g = o.query_sketch(sks[q])
        hm[q, g] = c
    assert np.array_equal(hm, z["hit_matrix"])
    assert list(hm[0]) == [32768, 31712, 30737, 29845, 28993, 28220, 27415, 26677, 25930]
    assert np.array_equal(o.matrix_counts().astype(np.uint32), hm)


def test_synth_known_answers():
    g = load_json("synth.json")
    o = Oracle()
    for x, y in g["mix"]:
        assert o.mix(x) == y
    assert bytes(o.synth_genome(0, 64)).decode() == g["genome0_head"]
    assert bytes(o.synth_genome(7, 64)).decode() == g["genome7_head"]
    assert zlib.crc32(bytes(o.synth_genome(3, 100000))) == g["genome3_crc_100k"]
    mut = o.synth_mutant(5, 9, 0.01, 100000)
    assert zlib.crc32(bytes(mut)) == g["mutant_g5_q9_d01_crc_100k"]
    assert int((mut != o.synth_genome(5, 100000)).sum()) == g["mutant_g5_q9_d01_nsub"]
    assert 800 < g["mutant_g5_q9_d01_nsub"] < 1200
    assert bytes(o.synth_read(12345, 5000000)).decode() == g["read12345"]


def test_matrix_wraps_mod_65536():
    """--matrix counters are uint16 for every S; S=17 identical sketches give 2^17 mod 2^16 = 0."""
    o = Oracle(K=31, S=17, W=4, H=2)
    rng = np.random.default_rng(5)
    sk, _ = o.compute_sketch(random_dna(rng, 400000))
    o.insert_sketch(sk, 0)
    o.insert_sketch(sk, 1)
    m = o.matrix_counts()
    assert int(m[0, 1]) == (1 << 17) % 65536 == 0
    c, g = o.query_sketch(sk)
    assert list(c) == [1 << 17, 1 << 17] and list(g) == [1, 0]  # (count,gid) descending


@needs_ref
@pytest.mark.ref
def test_differential_vs_reference():
    rng = np.random.default_rng(99)
    alpha = np.frombuffer(b"ACGTNacgtRYK-", np.uint8)
    pr = np.array([.22, .22, .22, .22, .03, .02, .02, .02, .01, .005, .005, .005, .005])
    pr /= pr.sum()
    for ps in [dict(K=31, S=6, W=8, H=4), dict(K=21, S=9, W=12, H=4), dict(K=31, S=8, W=12, H=4, genome_size=1000)]:
        o, r = Oracle(**ps), Ref(**ps)
        assert o.p.as_dict() == r.params
        checked = 0
        for _ in range(150):
            seq = rng.choice(alpha, size=int(rng.integers(ps["K"] + 1, 600)), p=pr)
            a, passes = o.compute_sketch(seq, max_passes=50000)
            if passes < 0:
                continue  # the reference would spin forever on this entry (1-2 k-mers)
            assert np.array_equal(a, r.compute_sketch(seq))
            checked += 1
        assert checked > 100
        r.close()


@needs_ref
@pytest.mark.ref
def test_index_query_matrix_vs_reference():
    rng = np.random.default_rng(3)
    ps = dict(K=31, S=7, W=8, H=4)
    o, r = Oracle(J=0.1, **ps), Ref(J=0.1, **ps)
    base = random_dna(rng, 2000)
    ents = []
    for i in range(30):
        s = base.copy() if i % 2 else random_dna(rng, 2000)
        pos = rng.integers(0, 2000, 10)
        s[pos] = random_dna(rng, 10)
        ents.append(s)
    sks = o.sketch_many(ents)
    for g, sk in enumerate(sks):
        o.insert_sketch(sk, g)
        r.insert_sketch(sk, g, f"e{g}")
    sizes, gids = r.export_postings()
    rp, og = o.csr()
    assert np.array_equal(np.diff(rp).astype(np.uint32), sizes) and np.array_equal(og, gids)
    for q in range(0, 30, 7):
        c1, g1 = o.query_sketch(sks[q])
        c2, g2 = r.query_sketch(sks[q])
        assert np.array_equal(c1, c2) and np.array_equal(g1, g2)
    r.query_matrix()
    lines = r.read_output().decode().split("\n")
    m = o.matrix_counts()
    for q in range(30):
        exp = ["%g" % (int(v) / o.F) if v >= o.p.min_score else "0" for v in m[q]]
        assert lines[1 + q].split("\t")[1:-1] == exp
    r.close()
