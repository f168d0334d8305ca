#!/usr/bin/env python3
"""Per-superblock phase timing of two loop-filter workers (columns 4 and 5 of the 1440p clip).
Needs a library built with -DVP9B_LF_TIMING:  VP9B_LIB=<that .so> python tools/lf_phases.py"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import cuda_vp9_b200 as vb  # noqa: E402

clip = sys.argv[1] if len(sys.argv) > 1 else "bench_1440p"
_, _, chunks = vb.ivf_read(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "clips", clip + ".ivf"))
parser = vb.HostParser()
packed = []
for ch in chunks:
    for fr in vb.parse_superframe(ch):
        p, _ = parser.parse(fr)
        if p is not None:
            packed.append(p)
gpu = vb.GpuRecon(0)
up = [gpu.upload(p) for p in packed]
for rep in range(2):
    for d, h in up[:6]:
        gpu.submit_device(d, h)
gpu.sync()
buf = (C.c_ulonglong * 1024)()
gpu.L.vp9gpu_debug_read(buf)
names = ["dsets", "A", "sync", "H7", "E", "store0", "publish", "waitL", "D+sync", "H", "store1", "carry"]
t00 = min(v for v in buf if v)
for w in range(2):
    print("worker %d:  row start(ns) | %s | total" % (w + 4, " ".join("%7s" % n for n in names)))
    for r in range(32):
        t = [buf[(w * 32 + r) * 16 + k] for k in range(13)]
        if t[0] == 0:
            continue
        # stamp 8 is only written when the worker waited for its left neighbour
        if t[8] == 0:
            t[8] = t[7]
        d = [t[k + 1] - t[k] for k in range(12)]
        print("  %2d %9d | %s | %6d" % (r, t[0] - t00, " ".join("%7d" % x for x in d), t[12] - t[0]))
