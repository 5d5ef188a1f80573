"""Worker of tests/test_shard_gloo.py: one rank of a world-size-N gloo job on CPU.  The compute
object is an adapter over the CPU oracle (test infrastructure) with the niqki_b200.Index method
names, so the host-side shard / all-gather / merge logic of niqki_b200/shard.py runs unchanged."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


class OracleEngine:
    def __init__(self, **ps):
        from oracle.oracle import Oracle

        self.o = Oracle(**ps)
        self.F = self.o.F
        self.gid_base = 0

    def sketch_many(self, seqs):
        return self.o.sketch_many(seqs), None

    def insert_sketches(self, sketches, gid_base=0):
        self.gid_base = gid_base
        self.n = len(sketches)
        self.o.index_reset()
        for i, sk in enumerate(sketches):
            self.o.insert_sketch(sk, gid_base + i)

    def query_sketches(self, sketches, min_score=None):
        ptr, c, g = self.o.query_batch(np.ascontiguousarray(sketches, np.int32))
        # the oracle's counters start at gid 0: keep this shard's genomes only
        keep = (g >= self.gid_base) & (g < self.gid_base + self.n)
        per_q = [int(keep[int(ptr[q]):int(ptr[q + 1])].sum()) for q in range(len(sketches))]
        return np.concatenate([[0], np.cumsum(per_q)]).astype(np.uint64), c[keep], g[keep]


def entries(n, length, seed):
    rng = np.random.default_rng(seed)
    acgt = np.frombuffer(b"ACGT", np.uint8)
    fam = [rng.choice(acgt, size=length) for _ in range(5)]
    out = []
    for i in range(n):
        s = fam[i % 5].copy()
        pos = rng.integers(0, length, 25)
        s[pos] = rng.choice(acgt, size=25)
        out.append(s)
    return out


def main():
    import torch.distributed as dist

    from niqki_b200.shard import ShardedIndex

    cfg = json.loads(sys.argv[1])
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        dist.init_process_group("gloo")
    ps = dict(K=31, S=8, W=10, H=4, J=cfg["J"])
    ents = entries(cfg["n"], 3000, 1)
    qs = entries(cfg["nq"], 3000, 1)[::-1] + [np.random.default_rng(9).choice(np.frombuffer(b"ACGT", np.uint8), size=3000)]
    sh = ShardedIndex(OracleEngine(**ps), dist if world > 1 else None)
    lo, hi = sh.index_entries(len(ents), lambda a, b: ents[a:b])
    res = sh.query_entries(len(qs), lambda a, b: qs[a:b])
    if sh.rank == 0:
        ptr, c, g = res
        np.savez(cfg["out"], ptr=ptr, counts=c, gids=g, lo=lo, hi=hi)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
