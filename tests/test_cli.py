"""The C++ host `niqki_b200/bin/niqki_b200` (niqki CLI contract) against outputs of the reference
binary recorded in tests/golden/ (decompressed bytes, reference run with OMP_NUM_THREADS=1)."""
import gzip
import hashlib
import json
import os
import subprocess

import numpy as np
import pytest

from tests.util import c1_genomes, load_npz

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "niqki_b200", "bin", "niqki_b200")


def run_cli(args, cwd, check=True):
    r = subprocess.run([CLI] + args, cwd=cwd, capture_output=True, text=True, timeout=900)
    if check:
        assert r.returncode == 0, r.stdout + r.stderr
    return r


def gunzip(path):
    return gzip.open(path, "rb").read()


def test_cli_builds_and_prints_usage():
    assert os.path.exists(CLI), "run __graft_entry__.build() first"
    r = subprocess.run([CLI, "--help"], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0 and "--indexlines" in r.stderr and "--querylines, -l" in r.stderr
    r = subprocess.run([CLI, "--bogus"], capture_output=True, text=True, timeout=60)
    assert r.returncode != 0 and "Bad usage!!!" in r.stdout


def test_cli_has_no_cpu_fallback(tmp_path):
    """Without a CUDA device the host must fail loudly, not compute on the CPU."""
    import ctypes as C

    import niqki_b200

    n = C.c_int(0)
    if niqki_b200.lib().nq_device_count(C.byref(n)) == 0 and n.value > 0:
        pytest.skip("a CUDA device is present")
    (tmp_path / "r.fa").write_bytes(b">a\n" + b"ACGT" * 40 + b"\n")
    r = run_cli(["-i", "r.fa", "-S", "6", "-O", "o.gz"], tmp_path, check=False)
    assert r.returncode != 0 and "CUDA" in (r.stdout + r.stderr)


@pytest.fixture(scope="module")
def c1_dir(tmp_path_factory):
    """The nine config-1 genomes as single-record, single-line .fa.gz files + the file of files."""
    d = tmp_path_factory.mktemp("c1")
    seqs, z = c1_genomes()
    for name, hdr, s in zip(z["names"], z["headers"], seqs):
        with gzip.open(d / str(name), "wb", compresslevel=1) as f:
            f.write(str(hdr).encode() + b"\n" + s.tobytes() + b"\n")
    (d / "file_of_file.txt").write_text("".join(f"{n}\n" for n in z["names"]))
    return d, z


@pytest.mark.gpu
def test_cli_c1_matrix(c1_dir):
    """BASELINE configs[0]: `niqki -M file_of_file.txt` at defaults; decompressed output bytes equal
    the reference's (md5 6f390d5a...)."""
    d, z = c1_dir
    r = run_cli(["-M", "file_of_file.txt", "-O", "m.gz"], d)
    out = gunzip(d / "m.gz")
    assert out == bytes(z["cli_matrix_text"])
    assert hashlib.md5(out).hexdigest() == str(z["cli_matrix_md5"])
    assert "Number of indexed genomes" in r.stdout and "| " in r.stdout


@pytest.mark.gpu
def test_cli_c1_index_query(c1_dir):
    d, z = c1_dir
    run_cli(["-I", "file_of_file.txt", "-Q", "file_of_file.txt", "-P", "-O", "q.gz"], d)
    out = gunzip(d / "q.gz")
    assert out == bytes(z["cli_query_text"])
    assert hashlib.md5(out).hexdigest() == str(z["cli_query_md5"])
    # -I chdirs to the list's directory, -Q resolves against the cwd: run from the parent
    sub = os.path.basename(str(d))
    (d.parent / "qfof.txt").write_text("".join(f"{sub}/{n}\n" for n in z["names"][:2]))
    run_cli(["-I", f"{sub}/file_of_file.txt", "-Q", "qfof.txt", "-J", "0.9", "-O", "q2.gz"], d.parent)
    lines = gunzip(d.parent / "q2.gz").decode().splitlines()
    assert lines[0].startswith(f"{sub}/ecoli01p.fa.gz ecoli01p.fa.gz:1 ecoli02p.fa.gz:0.967773 ")
    assert len(lines) == 2


@pytest.fixture(scope="module")
def small_dir(tmp_path_factory):
    d = tmp_path_factory.mktemp("small")
    z = load_npz("small_index.npz")
    n = z["sketches"].shape[0]
    for i in range(n):
        (d / f"entry{i}").write_bytes(b">e\n" + z[f"entry_{i}"].tobytes() + b"\n")
    (d / "fof.txt").write_text("".join(f"entry{i}\n" for i in range(n)))
    for i in range(4):
        # query 1 is written as a multi-line record: line breaks must not matter
        s = z[f"query_{i}"].tobytes()
        body = b"\n".join(s[j:j + 61] for j in range(0, len(s), 61)) if i == 1 else s
        (d / f"query{i}").write_bytes(b">q\n" + body + b"\n")
    (d / "qfof.txt").write_text("".join(f"query{i}\n" for i in range(4)))
    return d, z


@pytest.mark.gpu
def test_cli_small_index_all_outputs(small_dir):
    """48-entry index (S=8 W=8 J=0.05): pretty text, binary records, matrix text and the dump file
    all byte-equal to what the reference wrote."""
    d, z = small_dir
    ps = json.loads(str(z["params"]))
    flags = ["-K", str(ps["K"]), "-S", str(ps["S"]), "-W", str(ps["W"]), "-H", str(ps["H"]), "-J", str(ps["J"])]
    run_cli(flags + ["-I", "fof.txt", "-Q", "qfof.txt", "-O", "pretty.gz", "-D", "dump.gz"], d)
    assert gunzip(d / "pretty.gz") == bytes(z["pretty"])
    dump = gunzip(d / "dump.gz")
    assert len(dump) == int(z["dump_len"]) and dump[:24] == bytes(z["dump_head"])
    assert hashlib.md5(dump).hexdigest() == str(z["dump_md5"])
    run_cli(flags + ["-I", "fof.txt", "-Q", "qfof.txt", "-O", "binary.gz", "--binary"], d)
    assert gunzip(d / "binary.gz") == bytes(z["binary"])
    run_cli(flags + ["-M", "fof.txt", "-O", "matrix.gz"], d)
    assert gunzip(d / "matrix.gz") == bytes(z["matrix_text"])
    # --load: parameters and min_score come from the file; same hits
    run_cli(["-L", "dump.gz", "-Q", "qfof.txt", "-O", "loaded.gz"], d)
    assert gunzip(d / "loaded.gz") == bytes(z["pretty"])
    # --load followed by more entries, then --dump: equals one build over everything
    (d / "fof_a.txt").write_text("".join(f"entry{i}\n" for i in range(30)))
    (d / "fof_b.txt").write_text("".join(f"entry{i}\n" for i in range(30, 48)))
    run_cli(flags + ["-I", "fof_a.txt", "-D", "dump_a.gz", "-O", "x.gz"], d)
    run_cli(["-L", "dump_a.gz", "-I", "fof_b.txt", "-D", "dump_ab.gz", "-Q", "qfof.txt", "-O", "loaded2.gz"], d)
    assert gunzip(d / "dump_ab.gz") == dump
    assert gunzip(d / "loaded2.gz") == bytes(z["pretty"])


@pytest.mark.gpu
def test_cli_lines_mode(tmp_path):
    """--indexlines / --querylines on FASTA and FASTQ (multi-line records, lowercase, N runs, records
    of length < K / == K), -G, --dump / --load: decompressed outputs equal the reference CLI's."""
    z = load_npz("cli_lines.npz")
    (tmp_path / "reads.fa").write_bytes(z["reads_fa"].tobytes())
    (tmp_path / "reads.fq").write_bytes(z["reads_fq"].tobytes())
    (tmp_path / "queries.fa").write_bytes(z["queries_fa"].tobytes())
    with gzip.open(tmp_path / "reads.fa.gz", "wb") as f:
        f.write(z["reads_fa"].tobytes())
    runs = json.loads(str(z["runs"]))
    for name, args in runs.items():
        run_cli(args + ["-O", f"{name}.gz"], tmp_path)
        assert gunzip(tmp_path / f"{name}.gz") == bytes(z[f"out_{name}"]), name
    dump = gunzip(tmp_path / "dump.gz")
    assert len(dump) == int(z["dump_len"]) and hashlib.md5(dump).hexdigest() == str(z["dump_md5"])
    # gzip-compressed input is read transparently (zstr does the same)
    run_cli(["-i", "reads.fa.gz", "-l", "queries.fa", "-S", "8", "-J", "0.2", "-O", "gz.gz"], tmp_path)
    assert gunzip(tmp_path / "gz.gz") == bytes(z["out_fa"])


@pytest.mark.gpu
def test_cli_multi_record_files_are_merged(tmp_path):
    """the reference never terminates on a multi-record file in whole-file mode.  Defined
    behaviour here: records are scanned separately (own seed, no k-mer across the boundary),
    min-merged, densified once — checked against the oracle doing exactly that."""
    from oracle.oracle import Oracle

    rng = np.random.default_rng(3)
    acgt = np.frombuffer(b"ACGT", np.uint8)
    recs = [rng.choice(acgt, size=n).tobytes() for n in (5000, 40, 7000, 31, 12)]
    (tmp_path / "multi.fa").write_bytes(b"".join(b">r%d\n%s\n" % (i, s) for i, s in enumerate(recs)))
    (tmp_path / "single.fa").write_bytes(b">r\n" + recs[0] + b"\n")
    (tmp_path / "fof.txt").write_text("multi.fa\nsingle.fa\n")
    run_cli(["-M", "fof.txt", "-S", "8", "-O", "m.gz"], tmp_path)
    o = Oracle(K=31, S=8, W=12, H=4)
    scans = [o.sketch_scan(np.frombuffer(s, np.uint8))[0] for s in recs if len(s) > 31]
    merged = np.where(scans[0] < 0, np.iinfo(np.int32).max, scans[0])
    for s in scans[1:]:
        merged = np.minimum(merged, np.where(s < 0, np.iinfo(np.int32).max, s))
    merged = np.where(merged == np.iinfo(np.int32).max, -1, merged).astype(np.int32)
    multi = o.densify(merged)[0]
    single = o.compute_sketch(np.frombuffer(recs[0], np.uint8))[0]
    shared = int((multi == single).sum())
    rows = gunzip(tmp_path / "m.gz").decode().splitlines()
    assert rows[1].split("\t")[1:3] == ["1", "%g" % (shared / 256)]
    assert rows[2].split("\t")[1:3] == ["%g" % (shared / 256), "1"]


def _device_count():
    import ctypes as C

    import niqki_b200

    n = C.c_int(0)
    return n.value if niqki_b200.lib().nq_device_count(C.byref(n)) == 0 else 0


@pytest.mark.gpu
@pytest.mark.parametrize("gpus", [2, 3, 4, 8])
def test_cli_sharded_over_gpus_is_byte_equal(c1_dir, small_dir, tmp_path, gpus):
    """--gpus N (index sharded by genome id, queries all-gathered over NCCL, per-shard hits merged on
    the host, --matrix tiled over the devices, dump concatenated in shard order): every output byte
    equals the single-GPU / reference output.  Needs N visible devices (gpurun --gpus N)."""
    if _device_count() < gpus:
        pytest.skip(f"needs {gpus} CUDA devices")
    g = ["--gpus", str(gpus)]
    d, z = c1_dir
    run_cli(g + ["-M", "file_of_file.txt", "-O", f"m{gpus}.gz"], d)
    assert gunzip(d / f"m{gpus}.gz") == bytes(z["cli_matrix_text"])
    run_cli(g + ["-I", "file_of_file.txt", "-Q", "file_of_file.txt", "-P", "-O", f"q{gpus}.gz"], d)
    assert gunzip(d / f"q{gpus}.gz") == bytes(z["cli_query_text"])
    d, z = small_dir
    ps = json.loads(str(z["params"]))
    flags = g + ["-K", str(ps["K"]), "-S", str(ps["S"]), "-W", str(ps["W"]), "-H", str(ps["H"]), "-J", str(ps["J"])]
    run_cli(flags + ["-I", "fof.txt", "-Q", "qfof.txt", "-O", f"pretty{gpus}.gz", "-D", f"dump{gpus}.gz"], d)
    assert gunzip(d / f"pretty{gpus}.gz") == bytes(z["pretty"])
    dump = gunzip(d / f"dump{gpus}.gz")
    assert hashlib.md5(dump).hexdigest() == str(z["dump_md5"])
    run_cli(flags + ["-M", "fof.txt", "-O", f"matrix{gpus}.gz"], d)
    assert gunzip(d / f"matrix{gpus}.gz") == bytes(z["matrix_text"])
    run_cli(g + ["-L", f"dump{gpus}.gz", "-Q", "qfof.txt", "-O", f"loaded{gpus}.gz"], d)
    assert gunzip(d / f"loaded{gpus}.gz") == bytes(z["pretty"])
    (d / "fof_a.txt").write_text("".join(f"entry{i}\n" for i in range(30)))
    (d / "fof_b.txt").write_text("".join(f"entry{i}\n" for i in range(30, 48)))
    run_cli(flags + ["-I", "fof_a.txt", "-D", f"dump_a{gpus}.gz", "-O", "x.gz"], d)
    run_cli(g + ["-L", f"dump_a{gpus}.gz", "-I", "fof_b.txt", "-D", f"dump_ab{gpus}.gz", "-Q", "qfof.txt", "-O", f"loaded2{gpus}.gz"], d)
    assert gunzip(d / f"dump_ab{gpus}.gz") == dump
    assert gunzip(d / f"loaded2{gpus}.gz") == bytes(z["pretty"])
    # lines mode: reads as entries, far more entries than devices
    z = load_npz("cli_lines.npz")
    (tmp_path / "reads.fa").write_bytes(z["reads_fa"].tobytes())
    (tmp_path / "reads.fq").write_bytes(z["reads_fq"].tobytes())
    (tmp_path / "queries.fa").write_bytes(z["queries_fa"].tobytes())
    for name, args in json.loads(str(z["runs"])).items():
        run_cli(g + args + ["-O", f"{name}.gz"], tmp_path)
        assert gunzip(tmp_path / f"{name}.gz") == bytes(z[f"out_{name}"]), name
