#!/usr/bin/env python
"""SASS excerpts of the hot loops for profiles/ (instruction counts per base / posting must be checkable).

usage: python tools/sass_excerpt.py > profiles/r02_sass_hot_loops.txt
For every kernel below: the innermost loop that holds its shared-memory atomics (from the target of the
loop's backward branch to the branch), with an opcode histogram.  Needs only cuobjdump (no GPU)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "niqki_b200", "lib", "obj")
KERNELS = [
    ("sketch.o", "sketch_scan_kernelILb1ELi1024ELb1ELb1ELb1E", "K1+K2 character scan, default K/W/H, 1024 threads: 16 bases per loop iteration", 16),
    ("sketch_packed.o", "sketch_scan_packed_kernelILb1ELi1024ELb1ELb1E", "K1+K2 packed scan (k-mers as windows of the 2-bit stream), default K/W/H: 16 bases per iteration", 16),
    ("slab.o", "query_slab_kernelILi3ELi256ELi16ELi0E", "K4a granule form, 16-id granules, two queries per CTA (dual-word counters): one batch of 4 rounds = 512 id slots", 0),
    ("query.o", "query_count_seg_kernelItLi0ELi128EjLi8ELi64ELi4ELi3ELb0E", "K4a CSR segment-table form, SEG=8, 128 threads (shards of up to ~15k genomes)", 0),
    ("query.o", "query_count_seg_kernelItLi0ELi1024EjLi32ELi96ELi8ELi2ELb1E", "K4a split16 segment-table form, 1024 threads (65.6k-131k genomes)", 0),
]


def sass(obj, pattern):
    out = subprocess.run(["cuobjdump", "-sass", os.path.join(OBJ, obj)], capture_output=True, text=True).stdout
    blocks = out.split("Function : ")
    for b in blocks[1:]:
        name = b.split("\n", 1)[0].strip()
        if pattern in name:
            lines = []
            for ln in b.splitlines():
                m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
                if m:
                    lines.append((int(m.group(1), 16), m.group(2).strip()))
            return name, lines
    return None, []


def main():
    for obj, pat, what, per in KERNELS:
        name, lines = sass(obj, pat)
        print("=" * 110)
        print(f"{what}\n{name}  ({obj}, {len(lines)} SASS instructions in the kernel)")
        if not lines:
            print("  (not found)")
            continue
        addr = {a: i for i, (a, _) in enumerate(lines)}
        best = None
        for i, (a, ins) in enumerate(lines):
            m = re.search(r"BRA(?:\.\w+)*\s+(?:\w+,\s*)?0x([0-9a-f]+)", ins)
            if not m:
                continue
            t = int(m.group(1), 16)
            if t in addr and addr[t] < i:  # backward branch: loop [addr[t], i]
                body = lines[addr[t]:i + 1]
                atoms = sum(1 for _, x in body if "ATOMS" in x)
                if atoms and (best is None or len(body) < len(best[2]) and atoms >= best[1] // 2 or atoms > best[1] * 2):
                    best = (addr[t], atoms, body)
        if best is None:
            print("  (no loop with ATOMS found)")
            continue
        body = best[2]
        hist = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", x).split()[0].split(".")[0] for _, x in body)
        print(f"loop of {len(body)} instructions, {best[1]} ATOMS" + (f" -> {len(body) / per:.1f} SASS instructions per base" if per else ""))
        print("opcodes: " + ", ".join(f"{k} {v}" for k, v in hist.most_common()))
        for a, x in body[:140]:
            print(f"  /*{a:04x}*/ {x}")
        if len(body) > 140:
            print(f"  ... ({len(body) - 140} more)")


if __name__ == "__main__":
    main()
