#!/usr/bin/env python3
"""Target for compute-sanitizer (memcheck / racecheck / synccheck): N reconstruction contexts replay
the first frames of small clips concurrently, one host thread per context, so that the two spin-wait
wavefront kernels (k_intra, k_lf_sb) of different streams share the GPU the way they do in the bench.
usage: sanitize_streams.py [streams] [frames] [clip ...]"""
import os
import sys
import threading

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import cuda_vp9_b200 as vb  # noqa: E402


def main():
    n_streams = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    n_frames = int(sys.argv[2]) if len(sys.argv) > 2 else 6
    clips = sys.argv[3:] or ["odd_rt", "comp_lf", "p444", "hbd10"]
    root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "clips")
    for clip in clips:
        _, _, chunks = vb.ivf_read(os.path.join(root, clip + ".ivf"))
        parser = vb.HostParser()
        packed = []
        for ch in chunks:
            for fr in vb.parse_superframe(ch):
                p, _ = parser.parse(fr)
                if p is not None and len(packed) < n_frames:
                    packed.append(p)
        ctxs = [vb.GpuRecon(0) for _ in range(n_streams)]
        up = [[c.upload(p) for p in packed] for c in ctxs]

        def work(i):
            for d, h in up[i]:
                ctxs[i].submit_device(d, h)
            ctxs[i].sync()
        th = [threading.Thread(target=work, args=(i,)) for i in range(n_streams)]
        for t in th:
            t.start()
        for t in th:
            t.join()
        # the same frames once more, all streams through one launch per stage
        for f in range(len(packed)):
            vb.GpuRecon.submit_multi(vb.GpuRecon.make_multi(ctxs, [up[i][f] for i in range(n_streams)]))
        for c in ctxs:
            c.sync()
        print("%s: %d streams x %d frames, %d launches" % (clip, n_streams, len(packed), sum(c.launch_count() for c in ctxs)))
        for c in ctxs:
            c.close()


if __name__ == "__main__":
    main()
