#!/usr/bin/env python3
"""Per-superblock phase timing of the loop-filter wavefront (rows 2 and 3 of the 1440p clip).
Needs a library built with -DVP9B_LF_TIMING:  VP9B_LIB=<that .so> python tools/lf_phases.py"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import cuda_vp9_b200 as vb  # noqa: E402

_, _, chunks = vb.ivf_read(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "clips",
                                        "bench_1440p.ivf"))
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
buf = (C.c_ulonglong * 512)()
gpu.L.vp9gpu_debug_read(buf)
t00 = buf[0]
# stamps written by thread 0 of rows 2 and 3 (LFT(k) in k_loopfilter): 0 step start, 3 vertical pass done,
# 4 past B2 (halo rows in place), 5 horizontal pass done, 6 past B3, 7 past B5 (left edge of c+1 + carry done)
for row in range(2):
    prev = None
    print("row %d:  sb  start(ns) |  V  B2wait  H  B3wait  left+carry+B5 | step  gap" % (row + 2))
    for sb in range(32):
        t = [buf[(row * 32 + sb) * 8 + k] for k in range(8)]
        if t[0] == 0:
            continue
        print("  %2d %8d | %5d %5d %5d %5d %5d | %6d %s" % (
            sb, t[0] - t00, t[3] - t[0], t[4] - t[3], t[5] - t[4], t[6] - t[5], t[7] - t[6], t[7] - t[0],
            (t[0] - prev) if prev else "-"))
        prev = t[7]
