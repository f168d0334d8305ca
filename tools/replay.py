#!/usr/bin/env python3
"""Replay one pass of a bench clip on one reconstruction context (profiling target for ncu)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import cuda_vp9_b200 as vb  # noqa: E402


def main():
    clip = sys.argv[1] if len(sys.argv) > 1 else "bench_1440p"
    passes = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "clips", clip + ".ivf")
    _, _, chunks = vb.ivf_read(path)
    parser = vb.HostParser()
    packed = []
    for ch in chunks:
        for fr in vb.parse_superframe(ch):
            p, _ = parser.parse(fr)
            if p is not None:
                packed.append(p)
    gpu = vb.GpuRecon(0)
    up = [gpu.upload(p) for p in packed]
    if "--stages" in sys.argv:
        gpu.set_profiling(True)
        for d, h in up:  # warm
            gpu.submit_device(d, h)
        gpu.sync()
        gpu.stage_ms(reset=True)
        print("frame  intra_only  expand  inter  resid  intra_resid  intra  lf   (ms)  txb_intra  txb_inter  coefs")
        for i, (d, h) in enumerate(up):
            gpu.submit_device(d, h)
            gpu.sync()
            ms = gpu.stage_ms(reset=True)
            print("%3d %d  %s   %d %d %d" % (i, h.intra_only, " ".join("%7.3f" % m for m in ms[:6]),
                                             h.sec[3].count + h.sec[4].count + h.sec[5].count, h.sec[2].count, h.sec[7].count))
        return
    for _ in range(passes):
        for d, h in up:
            gpu.submit_device(d, h)
    gpu.sync()
    print("replayed %d frames x %d, %d launches" % (len(up), passes, gpu.launch_count()))


if __name__ == "__main__":
    main()
