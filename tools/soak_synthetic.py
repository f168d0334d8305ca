"""Soak: many seeds of the synthetic GPU-vs-oracle cases (tests/test_synthetic.py) on one long-lived
context (FRESH=0, default) or a fresh context per case (FRESH=1).  Usage: python tools/soak_synthetic.py [first_seed n_seeds]"""
import os, sys, threading, time
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import test_synthetic as ts
import cuda_vp9_b200 as vb
from oracle import ref_decoder
bad = 0
t0 = time.time()
prev = None
FRESH = int(os.environ.get("FRESH", "0"))
FIRST = int(sys.argv[1]) if len(sys.argv) > 1 else 100
COUNT = int(sys.argv[2]) if len(sys.argv) > 2 else 40
gpu = vb.GpuRecon(0)
for seed in range(FIRST, FIRST + COUNT):
    for name, make in ts.CASES:
        f, refs = make(seed)
        want = ts.run(ref_decoder.RefRecon(), f, refs)
        if FRESH:
            gpu.close(); gpu = vb.GpuRecon(0)
        got = ts.run(gpu, f, refs)
        if not all(np.array_equal(a, b) for a, b in zip(got, want)):
            bad += 1
            d = [int((a != b).sum()) for a, b in zip(got, want)]
            loc = [tuple(np.argwhere(a != b)[0]) if (a != b).any() else None for a, b in zip(got, want)]
            got2 = ts.run(gpu, f, refs)
            again = all(np.array_equal(a, b) for a, b in zip(got2, want))
            print("MISMATCH", name, seed, d, loc, "prev", prev, "rerun_ok", again, flush=True)
        prev = name
gpu.close()
print("synthetic: %d cases, %d mismatches, %.0fs" % (COUNT * len(ts.CASES), bad, time.time() - t0), flush=True)
sys.exit(1 if bad else 0)
