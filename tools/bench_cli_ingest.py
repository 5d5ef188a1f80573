#!/usr/bin/env python
"""CLI ingest throughput: niqki_b200 -I over plain and gzip FASTA files, -i over a
reads file.  Writes synthetic inputs to a scratch directory, times the binary, prints one JSON line."""
import gzip
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
CLI = os.path.join(ROOT, "niqki_b200", "bin", "niqki_b200")


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    L = 5_000_000
    rng = np.random.default_rng(1)
    acgt = np.frombuffer(b"ACGT", np.uint8)
    out = {}
    with tempfile.TemporaryDirectory() as d:
        base = acgt[rng.integers(0, 4, size=L, dtype=np.uint8)]
        names_p, names_z = [], []
        for g in range(n):
            s = np.roll(base, 7919 * g).copy()
            s[rng.integers(0, L, 5000)] = acgt[rng.integers(0, 4, 5000)]
            body = b">g%d\n" % g + s.tobytes() + b"\n"
            open(os.path.join(d, f"g{g}.fa"), "wb").write(body)
            with gzip.open(os.path.join(d, f"g{g}.fa.gz"), "wb", compresslevel=1) as f:
                f.write(body)
            names_p.append(f"g{g}.fa"); names_z.append(f"g{g}.fa.gz")
        open(os.path.join(d, "plain.txt"), "w").write("".join(x + "\n" for x in names_p))
        open(os.path.join(d, "gz.txt"), "w").write("".join(x + "\n" for x in names_z))
        nreads = 2_000_000
        pos = rng.integers(0, L - 150, nreads)
        with open(os.path.join(d, "reads.fa"), "wb") as f:
            for i in range(0, nreads, 100000):
                f.write(b"".join(b">r%d\n" % (i + j) + base[p:p + 150].tobytes() + b"\n" for j, p in enumerate(pos[i:i + 100000])))
        for name, args, units in [("fof_plain", ["-I", "plain.txt"], n * L), ("fof_gz", ["-I", "gz.txt"], n * L),
                                  ("lines_plain", ["-i", "reads.fa", "-S", "8"], nreads * 150)]:
            best = None
            for _ in range(2):
                t0 = time.perf_counter()
                r = subprocess.run([CLI] + args + ["-O", "o.gz"], cwd=d, capture_output=True, text=True)
                dt = time.perf_counter() - t0
                assert r.returncode == 0, r.stdout + r.stderr
                lasted = [ln for ln in r.stdout.splitlines() if "Indexing lasted" in ln]
                secs = float(lasted[0].split("|")[2]) if lasted else dt
                best = secs if best is None else min(best, secs)
            out[name] = {"seconds_indexing": best, "mbases_per_s": units / best / 1e6}
    out["host_cores"] = os.cpu_count()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
