"""Worker of tests/test_gpu_parity.py::test_sharded_query_over_processes_equals_single_gpu: one rank of a
torch.distributed.run job, one GPU per rank.  Index sharded by genome id (nq_shard_range), query sketches
all-gathered through the library's own communicator (nq_comm_init_rank / nq_allgather_sketches, NCCL),
per-shard hit lists gathered to rank 0 and merged with nq_hits_merge; rank 0 then builds the whole index
on its own GPU and compares.  Writes "OK" or the mismatch to the output file."""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist

    import niqki_b200
    from niqki_b200.capi import check, lib
    from niqki_b200.shard import Comm, merge_hits, torch_bcast_bytes
    from tests.test_gpu_parity import _fp_cdf, _queries_like, _realistic_sketches

    cfg = json.loads(sys.argv[1])
    world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    gloo = dist.new_group(backend="gloo")
    n, S, nq, J = cfg["n"], cfg["S"], cfg["nq"], cfg["J"]
    F = 1 << S
    rng = np.random.default_rng(cfg["seed"])            # the same data on every rank
    cdf = _fp_cdf()
    sks = _realistic_sketches(rng, n, F, cdf)
    sks[1::97] = sks[0]
    q = _queries_like(rng, sks, nq, cdf)
    ctx = niqki_b200.Context(local)
    comm = Comm(ctx, rank, world, torch_bcast_bytes(dist, dev))
    assert comm.info()["nranks"] == world
    ix = niqki_b200.Index(S=S, K=31, W=12, H=4, min_fract=J, ctx=ctx)
    b, e = C.c_uint64(), C.c_uint64()
    check(lib().nq_shard_range(n, world, rank, C.byref(b), C.byref(e)))
    lo, hi = b.value, e.value
    if hi > lo:
        ix.insert_sketches(torch.from_numpy(sks[lo:hi]).to(dev), gid_base=lo)
    m = -(-nq // world)                                  # rows per rank, missing rows = -1
    mine = np.full((m, F), -1, np.int32)
    part_q = q[rank * m:(rank + 1) * m]
    mine[:len(part_q)] = part_q
    d_mine = torch.from_numpy(mine).to(dev)
    d_all = torch.empty((m * world, F), dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    comm.allgather_sketches(ix.p, d_mine, d_all)
    ctx.sync()
    if hi > lo:
        part = ix.query_sketches(d_all)
    else:
        part = (np.zeros(m * world + 1, np.uint64), np.zeros(0, np.uint32), np.zeros(0, np.uint32))
    parts = [None] * world if rank == 0 else None
    dist.gather_object(part, parts, dst=0, group=gloo)
    if rank == 0:
        ptr, cnt, gid = merge_hits(parts)
        full = niqki_b200.Index(S=S, K=31, W=12, H=4, min_fract=J, ctx=ctx)
        full.insert_sketches(sks)
        pad = np.full((m * world, F), -1, np.int32)
        pad[:nq] = q
        eptr, ecnt, egid = full.query_sketches(pad)
        ok = np.array_equal(ptr, eptr) and np.array_equal(cnt, ecnt) and np.array_equal(gid, egid) and eptr[-1] > nq // 2
        open(cfg["out"], "w").write("OK" if ok else f"MISMATCH hits {ptr[-1]} vs {eptr[-1]}")
        full.close()
    dist.barrier()
    comm.close()
    ix.close()
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
