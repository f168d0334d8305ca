#!/usr/bin/env python3
"""Where does the e2e time go?  Per-call wall time inside N decoder threads."""
import os, sys, threading, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import cuda_vp9_b200 as vb

def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    passes = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "clips", "bench_1440p.ivf")
    w, h, chunks = vb.ivf_read(path)
    stats = [None] * n
    bar = threading.Barrier(n + 1)
    def worker(i):
        d = vb.Decoder(0)
        ring = [vb.PinnedFrame(w, h) for _ in range(3)]
        t_dec = t_get = t_sync = 0.0
        def run(p, timed):
            nonlocal t_dec, t_get, t_sync
            k = 0
            for _ in range(p):
                for ch in chunks:
                    t0 = time.perf_counter(); d.decode(ch); t1 = time.perf_counter()
                    ok = d.get_frame_into(ring[k % 3]); t2 = time.perf_counter()
                    if ok:
                        k += 1
                        if k % 3 == 0: d.sync()
                    t3 = time.perf_counter()
                    if timed: t_dec += t1 - t0; t_get += t2 - t1; t_sync += t3 - t2
            d.sync()
            return k
        run(1, False)
        bar.wait()
        t0 = time.perf_counter()
        k = run(passes, True)
        stats[i] = (k, time.perf_counter() - t0, t_dec, t_get, t_sync)
        bar.wait()
    th = [threading.Thread(target=worker, args=(i,)) for i in range(n)]
    for t in th: t.start()
    bar.wait(); t0 = time.perf_counter(); bar.wait(); dt = time.perf_counter() - t0
    for t in th: t.join()
    tot = sum(s[0] for s in stats)
    print("threads %d: %.0f fps total; per thread: frames %d wall %.2fs decode %.2fs get %.2fs sync %.2fs" %
          (n, tot / dt, stats[0][0], stats[0][1], stats[0][2], stats[0][3], stats[0][4]))

if __name__ == "__main__":
    main()
