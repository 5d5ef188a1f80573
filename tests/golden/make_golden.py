#!/usr/bin/env python
"""Generate the committed golden fixtures by RUNNING THE REFERENCE ITSELF (oracle/_ref).

Run in the build container, where /root/reference is mounted:

    make -C oracle all ref && python tests/golden/make_golden.py

Outputs (all small, committed):
  tests/golden/scalars.json     hashes / fingerprints / seeds / parameter blocks (incl. -G, B7)
  tests/golden/sketches.npz     adversarial sequences and the reference's sketches for them
  tests/golden/small_index.npz  a 48-entry index: postings, query hit lists, matrix text, dump bytes
  tests/golden/c1_genomes.npz   config-1 genomes (ecoli01 2-bit packed + substitution diffs for
                                02..09), their sketch CRCs, the integer hit matrix, CLI output md5s
The GPU box has no /root/reference: tests only ever read these files.
"""
import gzip
import hashlib
import json
import os
import subprocess
import sys
import tempfile
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.oracle import REF_CLI, Oracle, Ref  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
REFRES = "/root/reference/resources"


def adversarial_sequences(rng):
    """Sequences hitting SURVEY App. B: non-ACGT, lowercase, seed poisoning, short, runs."""
    acgt = np.frombuffer(b"ACGT", np.uint8)

    def rnd(n):
        return bytes(rng.choice(acgt, size=n))

    seqs = {}
    seqs["c4_104mer"] = (b"ACGTACGTTAGCTAGCTAGGATCGATCGATTTAGCGCGATATCGCGGCTAGCTAGCATCGATCAGCTACGACTAGCATCAGC"
                         b"ATCGACTAGCATCAGCATCGAC")
    seqs["rand_150"] = rnd(150)
    seqs["rand_1k"] = rnd(1000)
    seqs["rand_20k"] = rnd(20000)
    s = bytearray(rnd(600)); s[5] = ord("N")
    seqs["seed_poison_N"] = bytes(s)                      # B4: N inside the first K-1 bases
    s = bytearray(rnd(600)); s[3:9] = b"acgtac"
    seqs["seed_lowercase"] = bytes(s)                     # lowercase accepted in the seed only
    s = bytearray(rnd(600)); s[100:140] = b"n" * 40; s[300] = ord("a"); s[301] = ord("R")
    seqs["body_lower_other"] = bytes(s)                   # B3: body non-ACGT -> 0 on both strands
    seqs["polyA_200"] = b"A" * 200                        # canon = 0 -> h = 0 (B8)
    seqs["polyT_200"] = b"T" * 200
    seqs["polyN_200"] = b"N" * 200
    seqs["N_run_mid"] = rnd(200) + b"N" * 64 + rnd(200)
    seqs["len_K_plus_3"] = rnd(34)
    seqs["len_K_plus_9"] = rnd(40)
    seqs["mixed_case_all"] = bytes(rng.choice(np.frombuffer(b"ACGTacgtNn", np.uint8), size=800))
    seqs["boundary_30_N"] = rnd(29) + b"N" + rnd(300)     # last seed char (index K-2) foreign
    seqs["boundary_31_N"] = rnd(30) + b"N" + rnd(300)     # first rolling char (index K-1) foreign
    seqs["boundary_30_lc"] = rnd(29) + b"g" + rnd(300)
    seqs["boundary_31_lc"] = rnd(30) + b"g" + rnd(300)
    seqs["newline_inside"] = rnd(100) + b"\n" + rnd(100)
    return seqs


PARAM_SETS = [
    dict(K=31, S=6, W=8, H=4),
    dict(K=31, S=8, W=12, H=4),
    dict(K=21, S=10, W=12, H=4),
    dict(K=11, S=7, W=8, H=2),
    dict(K=31, S=12, W=12, H=4),
    dict(K=31, S=8, W=12, H=4, genome_size=1000),      # -G: stale masks (B7)
    dict(K=31, S=10, W=12, H=4, genome_size=5000000),
    dict(K=31, S=9, W=16, H=6),
    dict(K=15, S=4, W=10, H=5),
]


def terminates(o, seq):
    """The reference spins forever on some 1-2 k-mer entries; probe with the (pinned) oracle."""
    return o.compute_sketch(seq, max_passes=200000)[1] >= 0


def gen_scalars(rng):
    out = {"hash": [], "fingerprint": [], "seed": [], "params": []}
    r = Ref(K=31, S=6, W=8, H=4)
    xs = [0, 1, 42, 7, 0x0123456789ABCDEF, 2**62 - 1, 2**64 - 1] + [int(x) for x in rng.integers(0, 2**63, 40)]
    for x in xs:
        out["hash"].append({"x": x, "rev": r.revhash64(x), "unrev": r.unrevhash64(x),
                            "fam3": r.hash_family(x, 3), "fam1000": r.hash_family(x, 1000)})
    for s in [b"ACGT" * 7 + b"AC", b"acgtACGT" * 3 + b"ttttgg", b"ACGTNCGT" * 3 + b"AAAAAA", b"A" * 30, b"T" * 30,
              b"GATTACAGATTACAGATTACAGATTACAGA"[:30]]:
        v = r.str2numstrand(s)
        out["seed"].append({"s": s.decode(), "f0": v, "r0": r.rcb(v)})
    r.close()
    for ps in PARAM_SETS:
        r = Ref(**ps)
        hs = [0, 1, 0xAB, 0x8000000000000000, 0x4000000000000001, 0x00010000000000FF, 2**64 - 1]
        hs += [int(x) >> int(s) for x, s in zip(rng.integers(0, 2**63, 60), rng.integers(0, 63, 60))]
        out["fingerprint"].append({"params": ps, "derived": r.params,
                                   "cases": [[h, r.get_fingerprint(h)] for h in hs]})
        r.close()
    for J, S in [(0.1, 15), (0.1, 18), (0.3, 8), (0.999, 6), (0.05, 12), (0, 15), (1.0, 10)]:
        r = Ref(K=31, S=S, W=4, H=2, J=J)
        out["params"].append({"J": J, "S": S, "min_score": r.params["min_score"]})
        r.close()
    json.dump(out, open(os.path.join(OUT, "scalars.json"), "w"), indent=0)


def gen_sketches(rng):
    seqs = adversarial_sequences(rng)
    names = sorted(seqs)
    arrays = {"names": np.array(names), "param_sets": np.array(json.dumps(PARAM_SETS))}
    for n in names:
        arrays["seq_" + n] = np.frombuffer(seqs[n], np.uint8)
    for pi, ps in enumerate(PARAM_SETS):
        r = Ref(**ps)
        o = Oracle(**ps)
        for n in names:
            if len(seqs[n]) <= ps["K"] or not terminates(o, seqs[n]):
                continue  # gated by callers / reference hangs: no defined answer
            arrays[f"sk_{pi}_{n}"] = r.compute_sketch(seqs[n])
        r.close()
    np.savez_compressed(os.path.join(OUT, "sketches.npz"), **arrays)


def gen_small_index(rng):
    """48 related entries (a few families of mutated copies), S=8 W=8, J=0.05."""
    ps = dict(K=31, S=8, W=8, H=4)
    J = 0.05
    acgt = np.frombuffer(b"ACGT", np.uint8)
    fam = [rng.choice(acgt, size=3000) for _ in range(6)]
    entries = []
    for i in range(48):
        s = fam[i % 6].copy()
        nmut = int(rng.integers(0, 40))
        pos = rng.integers(0, s.size, nmut)
        s[pos] = rng.choice(acgt, size=nmut)
        entries.append(s)
    queries = [entries[3].copy(), fam[2].copy(), rng.choice(acgt, size=3000), entries[47][:1500].copy()]
    with tempfile.TemporaryDirectory() as td:
        r = Ref(J=J, out_path=os.path.join(td, "out.gz"), **ps)
        sketches = np.stack([r.compute_sketch(e) for e in entries])
        for g, sk in enumerate(sketches):
            r.insert_sketch(sk, g, f"entry{g}")
        sizes, gids = r.export_postings()
        qsk = np.stack([r.compute_sketch(q) for q in queries])
        hits = [r.query_sketch(sk) for sk in qsk]
        for qi, (c, g) in enumerate(hits):
            r.output_query(c, g, f"query{qi}", pretty=True)
        pretty = r.read_output()
        r.query_matrix()
        both = r.read_output()
        matrix_text = both[len(pretty):]
        dump_path = os.path.join(td, "dump.gz")
        r.dump(dump_path)
        dump_bytes = gzip.open(dump_path, "rb").read()
        r.close()
        # binary records (B9: unreachable from the CLI, but the format is part of the contract)
        r2 = Ref(J=J, out_path=os.path.join(td, "out2.gz"), **ps)
        for g, sk in enumerate(sketches):
            r2.insert_sketch(sk, g, f"entry{g}")
        for qi, (c, g) in enumerate(hits):
            r2.output_query(c, g, f"query{qi}", pretty=False)
        binary = r2.read_output()
        r2.close()
    arr = dict(params=np.array(json.dumps(dict(ps, J=J))), sketches=sketches, sizes=sizes, gids=gids,
               qsketches=qsk, pretty=np.frombuffer(pretty, np.uint8), binary=np.frombuffer(binary, np.uint8),
               matrix_text=np.frombuffer(matrix_text, np.uint8),
               dump_md5=np.array(hashlib.md5(dump_bytes).hexdigest()), dump_len=np.array(len(dump_bytes)),
               dump_head=np.frombuffer(dump_bytes[:24], np.uint8))
    for i, e in enumerate(entries):
        arr[f"entry_{i}"] = e
    for i, q in enumerate(queries):
        arr[f"query_{i}"] = q
    for qi, (c, g) in enumerate(hits):
        arr[f"hit_counts_{qi}"] = c
        arr[f"hit_gids_{qi}"] = g
    np.savez_compressed(os.path.join(OUT, "small_index.npz"), **arr)


def read_fasta_gz(path):
    lines = gzip.open(path, "rb").read().split(b"\n")
    return lines[0], np.frombuffer(b"".join(l for l in lines[1:] if not l.startswith(b">")), np.uint8)


def gen_c1():
    """Config 1: the nine bundled genomes.  They are a substitution chain (01 -> 02 -> ... -> 09,
    ~2.3k substitutions per step, same length, no non-ACGT), so they are stored as ecoli01 packed
    at 2 bits/base plus eight diff lists — 1.2 MB instead of 13 MB, and no reference file copied."""
    names = [l.strip() for l in open(os.path.join(REFRES, "file_of_file.txt")) if l.strip()]
    hdrs, seqs = zip(*[read_fasta_gz(os.path.join(REFRES, n)) for n in names])
    L = seqs[0].size
    assert all(s.size == L for s in seqs)
    code = np.zeros(256, np.uint8)
    for i, c in enumerate(b"ACGT"):
        code[c] = i
    c0 = code[seqs[0]]
    pad = (-L) % 4
    c0p = np.concatenate([c0, np.zeros(pad, np.uint8)]).reshape(-1, 4)
    packed = (c0p[:, 0] | (c0p[:, 1] << 2) | (c0p[:, 2] << 4) | (c0p[:, 3] << 6)).astype(np.uint8)
    arr = dict(names=np.array(names), headers=np.array([h.decode() for h in hdrs]), length=np.array(L),
               packed01=packed)
    for i in range(1, 9):
        pos = np.nonzero(seqs[i] != seqs[i - 1])[0].astype(np.uint32)
        arr[f"diff_pos_{i}"] = pos
        arr[f"diff_base_{i}"] = seqs[i][pos]
    r = Ref(K=31, S=15, W=12, H=4)
    sks = np.stack([r.compute_sketch(s) for s in seqs])
    for g, sk in enumerate(sks):
        r.insert_sketch(sk, g, names[g])
    arr["sketch_crc32"] = np.array([zlib.crc32(sk.astype("<i4").tobytes()) for sk in sks], np.uint32)
    arr["sketch_sum"] = sks.sum(axis=1).astype(np.int64)
    arr["sketch_head"] = sks[:, :16].copy()
    sizes, gids = r.export_postings()
    arr["postings_total"] = np.array(gids.size)
    arr["nonempty_lists"] = np.array(int((sizes > 0).sum()))
    arr["max_list"] = np.array(int(sizes.max()))
    arr["postings_crc32"] = np.array(zlib.crc32(gids.astype("<u4").tobytes()), np.uint32)
    arr["sizes_crc32"] = np.array(zlib.crc32(sizes.astype("<u4").tobytes()), np.uint32)
    hm = np.zeros((9, 9), np.uint32)
    for q in range(9):
        c, g = r.query_sketch(sks[q])
        hm[q, g] = c
    arr["hit_matrix"] = hm
    r.close()
    # CLI goldens (1 thread => canonical order)
    with tempfile.TemporaryDirectory() as td:
        for n in names:
            os.symlink(os.path.join(REFRES, n), os.path.join(td, n))
        os.symlink(os.path.join(REFRES, "file_of_file.txt"), os.path.join(td, "file_of_file.txt"))
        env = dict(os.environ, OMP_NUM_THREADS="1")
        subprocess.run([REF_CLI, "-M", "file_of_file.txt", "-O", "m.gz"], cwd=td, env=env, check=True,
                       stdout=subprocess.DEVNULL)
        mtxt = gzip.open(os.path.join(td, "m.gz"), "rb").read()
        subprocess.run([REF_CLI, "-I", "file_of_file.txt", "-Q", "file_of_file.txt", "-P", "-O", "q.gz"], cwd=td,
                       env=env, check=True, stdout=subprocess.DEVNULL)
        qtxt = gzip.open(os.path.join(td, "q.gz"), "rb").read()
    arr["cli_matrix_md5"] = np.array(hashlib.md5(mtxt).hexdigest())
    arr["cli_query_md5"] = np.array(hashlib.md5(qtxt).hexdigest())
    arr["cli_matrix_text"] = np.frombuffer(mtxt, np.uint8)
    arr["cli_query_text"] = np.frombuffer(qtxt, np.uint8)
    np.savez_compressed(os.path.join(OUT, "c1_genomes.npz"), **arr)
    print("c1 matrix md5", arr["cli_matrix_md5"], "query md5", arr["cli_query_md5"])


def gen_synth():
    """Known answers for the §8d generator (spec-defined, not reference-defined): pins the C and
    CUDA generators to each other and to this file."""
    o = Oracle()
    out = {"mix": [[x, o.mix(x)] for x in [0, 1, 42, 2**63, 2**64 - 1]],
           "genome0_head": bytes(o.synth_genome(0, 64)).decode(),
           "genome7_head": bytes(o.synth_genome(7, 64)).decode(),
           "genome3_crc_100k": zlib.crc32(bytes(o.synth_genome(3, 100000))),
           "mutant_g5_q9_d01_crc_100k": zlib.crc32(bytes(o.synth_mutant(5, 9, 0.01, 100000))),
           "mutant_g5_q9_d01_nsub": int((o.synth_mutant(5, 9, 0.01, 100000) != o.synth_genome(5, 100000)).sum()),
           "read12345": bytes(o.synth_read(12345, 5000000)).decode()}
    json.dump(out, open(os.path.join(OUT, "synth.json"), "w"), indent=0)


def gen_cli_lines():
    """--indexlines / --querylines / --dump / --load / -G through the reference CLI (1 thread) on a
    small FASTA and FASTQ with multi-line records, lowercase, N runs and records of length < K,
    == K and just above K.  Inputs and decompressed outputs are both stored."""
    rng = np.random.default_rng(77)
    acgt = np.frombuffer(b"ACGT", np.uint8)

    def rnd(n):
        return bytes(rng.choice(acgt, size=n))

    base = [rnd(int(rng.integers(150, 600))) for _ in range(8)]
    recs = []
    for i in range(40):
        s = bytearray(base[i % 8])
        for _ in range(int(rng.integers(0, 12))):
            s[int(rng.integers(0, len(s)))] = int(rng.choice(acgt))
        if i % 7 == 3:
            s[50:58] = b"NNNNNNNN"
        if i % 9 == 4:
            s[5:12] = bytes(s[5:12]).lower()
        if i % 11 == 5:
            s[100:104] = b"acgt"
        recs.append(bytes(s))
    recs[13] = rnd(20)   # < K: dropped by Biogetline
    recs[14] = rnd(31)   # == K: dropped by the > K gate
    recs[15] = rnd(64)
    fa = b""
    for i, s in enumerate(recs):
        fa += b">read%d some description\n" % i
        w = 70 if i % 2 else 10_000  # multi-line and single-line records
        fa += b"".join(s[j:j + w] + b"\n" for j in range(0, len(s), w))
    fq = b"".join(b"@read%d\n%s\n+\n%s\n" % (i, s, b"I" * len(s)) for i, s in enumerate(recs))
    qrecs = [recs[0], recs[9][:120], rnd(200), recs[22]]
    qfa = b"".join(b">q%d\n%s\n" % (i, s) for i, s in enumerate(qrecs))
    env = dict(os.environ, OMP_NUM_THREADS="1")
    arr = {"reads_fa": np.frombuffer(fa, np.uint8), "reads_fq": np.frombuffer(fq, np.uint8),
           "queries_fa": np.frombuffer(qfa, np.uint8)}
    runs = {
        "fa": ["-i", "reads.fa", "-l", "queries.fa", "-S", "8", "-J", "0.2"],
        "fq": ["-i", "reads.fq", "-l", "queries.fa", "-S", "8", "-J", "0.2"],
        "fa_self_J0": ["-i", "reads.fa", "-l", "reads.fa", "-S", "6", "-W", "8", "-K", "21"],
        "fa_G": ["-i", "reads.fa", "-l", "queries.fa", "-S", "8", "-J", "0.1", "-G", "300"],
        "fa_dump": ["-i", "reads.fa", "-S", "8", "-J", "0.2", "-D", "dump.gz"],
        "fa_load": ["-L", "dump.gz", "-l", "queries.fa"],
    }
    with tempfile.TemporaryDirectory() as td:
        open(os.path.join(td, "reads.fa"), "wb").write(fa)
        open(os.path.join(td, "reads.fq"), "wb").write(fq)
        open(os.path.join(td, "queries.fa"), "wb").write(qfa)
        for name, args in runs.items():
            subprocess.run([REF_CLI] + args + ["-O", f"{name}.gz"], cwd=td, env=env, check=True, timeout=600,
                           stdout=subprocess.DEVNULL)
            arr[f"out_{name}"] = np.frombuffer(gzip.open(os.path.join(td, f"{name}.gz"), "rb").read(), np.uint8)
        dump = gzip.open(os.path.join(td, "dump.gz"), "rb").read()
        arr["dump_md5"] = np.array(hashlib.md5(dump).hexdigest())
        arr["dump_len"] = np.array(len(dump))
    arr["runs"] = np.array(json.dumps(runs))
    np.savez_compressed(os.path.join(OUT, "cli_lines.npz"), **arr)
    print("cli_lines:", {k: int(v.size) for k, v in arr.items() if k.startswith("out_")})


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "cli_lines":
        gen_cli_lines()
        sys.exit(0)
    rng = np.random.default_rng(20261017)
    gen_scalars(rng)
    gen_sketches(rng)
    gen_small_index(rng)
    gen_synth()
    gen_c1()
    gen_cli_lines()
    print("golden fixtures written to", OUT)
