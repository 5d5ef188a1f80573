#!/usr/bin/env python
"""Host packer (K1, pack.cpp) throughput: nq_pack_sequences on random genomes, 1..N threads."""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from niqki_b200 import capi  # noqa: E402

L = capi.lib()
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1 << 30
rng = np.random.default_rng(1)
bases = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, size=n, dtype=np.uint8)]
offs = np.array([0, n // 2, n], np.uint64)
words, blocks = C.c_uint64(), C.c_uint64()
capi.check(L.nq_pack_sizes(n, C.byref(words), C.byref(blocks)))
codes = np.zeros(words.value, np.uint32)
blk = np.zeros(blocks.value, np.uint32)
pool = np.zeros(blocks.value * 32, np.uint16)
used = C.c_uint64()
for threads in [1, 2, 4, 8, 16, 32]:
    if threads > (os.cpu_count() or 1) * 2:
        break
    best = 1e9
    for _ in range(3):
        t0 = time.perf_counter()
        capi.check(L.nq_pack_sequences(bases.ctypes.data, offs.ctypes.data, 2, 31, codes.ctypes.data, blk.ctypes.data, pool.ctypes.data,
                                       blocks.value, C.byref(used), threads))
        best = min(best, time.perf_counter() - t0)
    print(f"pack {n / 1e9:.2f} Gbases, {threads:2d} threads: {n / best / 1e9:7.2f} Gbases/s", flush=True)

# ---- device half: the packed scan kernel against the character kernel on the same genomes
try:
    import torch

    if torch.cuda.is_available():
        import niqki_b200

        G, Lg = 400, 5_000_000
        ctx = niqki_b200.Context(0)
        ix = niqki_b200.Index(S=15, K=31, W=12, H=4, ctx=ctx)
        d_bases = torch.empty(G * Lg + 64, dtype=torch.uint8, device="cuda")
        capi.check(L.nq_synth_genomes_device(ctx.h, 42, 0, G, Lg, C.c_void_p(d_bases.data_ptr())))
        torch.cuda.synchronize()
        h = d_bases[: G * Lg].cpu().numpy()
        goffs = np.arange(G + 1, dtype=np.uint64) * Lg
        capi.check(L.nq_pack_sizes(G * Lg, C.byref(words), C.byref(blocks)))
        codes = np.zeros(words.value, np.uint32); blk = np.zeros(blocks.value, np.uint32); pool = np.zeros(64, np.uint16)
        capi.check(L.nq_pack_sequences(h.ctypes.data, goffs.ctypes.data, G, 31, codes.ctypes.data, blk.ctypes.data, pool.ctypes.data, 2,
                                       C.byref(used), 0))
        d_codes = torch.from_numpy(codes.view(np.int32)).cuda(); d_blk = torch.from_numpy(blk.view(np.int32)).cuda()
        d_pool = torch.from_numpy(pool.view(np.int16)).cuda()
        sk_a = torch.empty((G, ix.F), dtype=torch.int32, device="cuda"); sk_p = torch.empty_like(sk_a)
        fl = torch.empty(G, dtype=torch.int32, device="cuda")
        ctx.set_timing(True)
        for name, fn in [("character kernel", lambda: L.nq_sketch_batch_device(ctx.h, C.byref(ix.p), C.c_void_p(d_bases.data_ptr()), d_bases.numel(),
                                                                                goffs.ctypes.data, G, C.c_void_p(sk_a.data_ptr()), C.c_void_p(fl.data_ptr()))),
                         ("packed kernel", lambda: L.nq_sketch_batch_packed_device(ctx.h, C.byref(ix.p), C.c_void_p(d_codes.data_ptr()), C.c_void_p(d_blk.data_ptr()),
                                                                                    C.c_void_p(d_pool.data_ptr()), goffs.ctypes.data, G, C.c_void_p(sk_p.data_ptr()),
                                                                                    C.c_void_p(fl.data_ptr())))]:
            for _ in range(2):
                capi.check(fn())
            ctx.sync(); ctx.timing_reset()
            for _ in range(5):
                capi.check(fn())
            ctx.sync()
            ms, nl = ctx.timing()["scan"]
            print(f"{name}: {G * Lg * 5 / ms / 1e6:.1f} Gbases/s ({ms / 5:.3f} ms per {G} genomes)", flush=True)
        print("sketches equal:", bool(torch.equal(sk_a, sk_p)))
except ImportError:
    pass
