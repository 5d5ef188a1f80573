"""Shared helpers for the test-suite (fixture loading, input generators)."""
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_json(name):
    return json.load(open(os.path.join(GOLDEN, name)))


def load_npz(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


_c1_cache = {}


def c1_genomes():
    """The nine config-1 genomes rebuilt from tests/golden/c1_genomes.npz (see make_golden.gen_c1)."""
    if "seqs" not in _c1_cache:
        z = load_npz("c1_genomes.npz")
        L = int(z["length"])
        p = z["packed01"]
        codes = np.stack([p & 3, (p >> 2) & 3, (p >> 4) & 3, (p >> 6) & 3], axis=1).reshape(-1)[:L]
        s = np.frombuffer(b"ACGT", np.uint8)[codes]
        seqs = [s]
        for i in range(1, 9):
            s = s.copy()
            s[z[f"diff_pos_{i}"]] = z[f"diff_base_{i}"]
            seqs.append(s)
        _c1_cache["seqs"] = seqs
        _c1_cache["z"] = z
    return _c1_cache["seqs"], _c1_cache["z"]


def random_dna(rng, n, alphabet=b"ACGT"):
    return rng.choice(np.frombuffer(alphabet, np.uint8), size=n)
