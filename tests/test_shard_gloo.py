"""Host-side multi-GPU logic (niqki_b200/shard.py) on CPU: world-size-2 and -3 gloo jobs must give
the same merged hit lists as a single process (shard by gid, all-gather the query
sketches, merge = concatenate + sort (count, gid) descending)."""
import json
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from niqki_b200.shard import merge_hit_lists, owner_of, shard_range

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORKER = os.path.join(ROOT, "tests", "_shard_worker.py")


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def run_world(world, cfg, tmp_path):
    cfg = dict(cfg, out=str(tmp_path / f"w{world}.npz"))
    env = dict(os.environ, OMP_NUM_THREADS="1")
    if world == 1:
        cmd = [sys.executable, WORKER, json.dumps(cfg)]
    else:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
               "--master-addr", "127.0.0.1", "--master-port", str(free_port()), WORKER, json.dumps(cfg)]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    return np.load(cfg["out"])


def test_shard_range_partition():
    for n in (0, 1, 7, 8, 100000, 12501):
        for world in (1, 2, 3, 8):
            blocks = [shard_range(n, world, r) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
            for r, (lo, hi) in enumerate(blocks):
                if hi > lo:
                    assert owner_of(lo, n, world) == r and owner_of(hi - 1, n, world) == r
    assert shard_range(100000, 8, 3) == (37500, 50000)  # configs[2]: gid g -> GPU g // 12500


def test_merge_hit_lists_order():
    a = (np.array([0, 2, 2], np.uint64), np.array([5, 3], np.uint32), np.array([1, 0], np.uint32))
    b = (np.array([0, 1, 3], np.uint64), np.array([5], np.uint32).repeat(3), np.array([9, 8, 7], np.uint32))
    ptr, c, g = merge_hit_lists([a, b], 2)
    assert list(ptr) == [0, 3, 5]
    assert list(zip(c[:3], g[:3])) == [(5, 9), (5, 1), (3, 0)]  # count desc, then gid desc
    assert list(zip(c[3:], g[3:])) == [(5, 8), (5, 7)]


@pytest.mark.parametrize("J", [0.0, 0.1])
def test_world2_and_world3_equal_single(tmp_path, J):
    cfg = dict(n=23, nq=5, J=J)
    one = run_world(1, cfg, tmp_path)
    for world in (2, 3):
        w = run_world(world, cfg, tmp_path)
        for k in ("ptr", "counts", "gids"):
            assert np.array_equal(one[k], w[k]), (world, k)
    assert one["gids"].size > 0


def test_merge_hits_through_the_c_abi_equals_the_numpy_merge():
    """niqki_b200.shard.merge_hits (nq_hits_from_arrays + nq_hits_merge, host code of the product library)
    against the numpy statement of the same merge."""
    from niqki_b200.shard import merge_hits

    rng = np.random.default_rng(5)
    nq = 23
    parts = []
    for s in range(4):
        ptr, c, g = [0], [], []
        for q in range(nq):
            k = int(rng.integers(0, 5))
            gs = rng.choice(np.arange(s * 500, (s + 1) * 500), size=k, replace=False)
            cs = rng.integers(1, 7, size=k)
            for i in np.lexsort((gs, cs))[::-1]:
                c.append(int(cs[i])); g.append(int(gs[i]))
            ptr.append(len(c))
        parts.append((np.array(ptr, np.uint64), np.array(c, np.uint32), np.array(g, np.uint32)))
    a = merge_hits(parts)
    b = merge_hit_lists(parts, nq)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
