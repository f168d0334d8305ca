"""Corrupt / truncated streams must produce an error code, never a crash (the reference relies on
vpx_internal_error + setjmp, vp9_decodeframe.c:1431, test/invalid_file_test.cc).  The mutated
streams are parsed in a child process so that a memory fault fails the test instead of pytest."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT, clip_path

CHILD = r"""
import random, sys
sys.path.insert(0, %(root)r)
import cuda_vp9_b200 as vb
path, seed, rounds = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
_, _, chunks = vb.ivf_read(path)
rng = random.Random(seed)
errors = ok = 0
for it in range(rounds):
    p = vb.HostParser()
    victim = rng.randrange(len(chunks))
    for i, ch in enumerate(chunks):
        data = bytearray(ch)
        if i == victim:
            kind = rng.randrange(4)
            if kind == 0 and len(data) > 4:                      # truncate
                data = data[:rng.randrange(1, len(data))]
            elif kind == 1:                                      # flip a few bytes anywhere
                for _ in range(rng.randrange(1, 8)):
                    data[rng.randrange(len(data))] ^= 1 << rng.randrange(8)
            elif kind == 2:                                      # trash the header bytes
                for k in range(min(len(data), rng.randrange(1, 24))):
                    data[k] = rng.randrange(256)
            else:                                                # random garbage of random length
                data = bytearray(rng.randrange(256) for _ in range(rng.randrange(1, 200)))
        try:
            frames = vb.parse_superframe(bytes(data))
            for fr in frames:
                p.parse(fr)
            ok += 1
        except vb.VP9BError:
            errors += 1
            break            # libvpx callers stop (or resync on a key frame) after an error
    p.close()
print("done", ok, errors)
"""


@pytest.mark.parametrize("clip,rounds", [("odd_rt", 60), ("tiles_rt", 12), ("sub8", 25), ("hbd10", 25), ("aq", 25),
                                         ("scaled_rt", 25), ("intra_only", 20), ("p444", 25)])
def test_mutated_streams_do_not_crash(clip, rounds):
    r = subprocess.run([sys.executable, "-c", CHILD % {"root": ROOT}, clip_path(clip), "1234", str(rounds)],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, "parser crashed: rc=%d\n%s" % (r.returncode, r.stderr[-2000:])
    assert r.stdout.startswith("done")
