"""Randomised packed frames: CUDA path vs the C oracle on inputs no encoder would choose
(tests/synth_packed.py).  The CPU part checks that the generator produces buffers the oracle accepts
and that it is deterministic; the GPU part is the parity test."""
import numpy as np
import pytest

import cuda_vp9_b200 as vb
from oracle import ref_decoder

import synth_packed as sp

DST = 3


def build_inter(seed, w, h, bd, ss, scaled, lf):
    rng = np.random.default_rng(seed)
    f = sp.Frame(w, h, bd=bd, ss=ss, dst_slot=DST)
    refs = []
    for slot in range(3):
        if scaled and slot:
            rw = int(w * rng.uniform(0.6, 1.9)) & ~1
            rh = int(h * rng.uniform(0.6, 1.9)) & ~1
        else:
            rw, rh = w, h
        f.refs.append((slot, rw, rh))
        refs.append((slot, sp.random_planes(rng, rw, rh, bd, ss, smooth=True)))
    sp.add_inter_blocks(rng, f)
    sp.add_inter_residuals(rng, f)
    if lf:
        sp.add_loop_filter(rng, f)
    return f, refs


def build_mixed(seed, w, h, bd, ss):
    """Inter frame in which a random third of the superblocks is intra coded."""
    rng = np.random.default_rng(seed)
    f = sp.Frame(w, h, bd=bd, ss=ss, dst_slot=DST)
    refs = []
    for slot in range(3):
        f.refs.append((slot, w, h))
        refs.append((slot, sp.random_planes(rng, w, h, bd, ss, smooth=True)))
    intra_sbs = {(r, c) for r in range(f.sb_rows) for c in range(f.sb_cols) if rng.random() < 0.35}
    sp.add_inter_blocks(rng, f, intra_sbs=intra_sbs)
    sp.add_inter_residuals(rng, f, intra_sbs=intra_sbs)
    sp.add_intra_blocks(rng, f, only_sbs=intra_sbs)
    sp.add_loop_filter(rng, f)
    return f, refs


def build_intra(seed, w, h, bd, ss, lf, lossless=False, tx=None):
    rng = np.random.default_rng(seed)
    f = sp.Frame(w, h, bd=bd, ss=ss, dst_slot=DST, intra_only=True, lossless=lossless)
    sp.add_intra_blocks(rng, f, tx_size_of_sb=tx)
    if lf:
        sp.add_loop_filter(rng, f)
    return f, []


def run(dev, f, refs):
    for slot, planes in refs:
        dev.set_frame(slot, planes, f.bd, f.ss, f.ss)
    blank = [np.zeros(((f.h + (f.ss if p else 0)) >> (f.ss if p else 0), (f.w + (f.ss if p else 0)) >> (f.ss if p else 0)),
                      dtype=np.uint16 if f.bd > 8 else np.uint8) for p in range(3)]
    dev.set_frame(DST, blank, f.bd, f.ss, f.ss)
    dev.submit(f.pack())
    return dev.get_frame(DST, f.w, f.h, f.bd, f.ss, f.ss)


CASES = [
    ("inter8", lambda s: build_inter(s, 200, 136, 8, 1, False, True)),
    ("inter8_scaled", lambda s: build_inter(s, 176, 144, 8, 1, True, False)),
    ("inter10", lambda s: build_inter(s, 136, 200, 10, 1, False, True)),
    ("inter12_scaled", lambda s: build_inter(s, 144, 112, 12, 1, True, True)),
    ("inter8_444", lambda s: build_inter(s, 136, 136, 8, 0, False, True)),
    ("inter10_444", lambda s: build_inter(s, 128, 72, 10, 0, True, True)),
    ("mixed8", lambda s: build_mixed(s, 328, 200, 8, 1)),
    ("mixed10_444", lambda s: build_mixed(s, 200, 136, 10, 0)),
    ("intra8", lambda s: build_intra(s, 200, 136, 8, 1, True)),
    ("intra8_tx32", lambda s: build_intra(s, 192, 128, 8, 1, False, tx=3)),
    ("intra12", lambda s: build_intra(s, 136, 200, 12, 1, True)),
    ("intra8_444", lambda s: build_intra(s, 136, 136, 8, 0, True)),
    ("intra8_lossless", lambda s: build_intra(s, 72, 136, 8, 1, False, lossless=True)),
    ("intra10_lossless", lambda s: build_intra(s, 136, 72, 10, 1, False, lossless=True)),
]


@pytest.mark.parametrize("name,make", CASES[::3], ids=[c[0] for c in CASES[::3]])
def test_oracle_accepts_synthetic_frames(name, make):
    f, refs = make(7)
    a = run(ref_decoder.RefRecon(), f, refs)
    f2, refs2 = make(7)
    assert f.pack() == f2.pack()                       # deterministic generator
    b = run(ref_decoder.RefRecon(), f2, refs2)
    for pa, pb in zip(a, b):
        assert np.array_equal(pa, pb)
    assert any(int(p.max()) > int(p.min()) for p in a)  # something was reconstructed


_shared_gpu = []


@pytest.mark.gpu
@pytest.mark.parametrize("seed", [1, 2, 3, 4])
@pytest.mark.parametrize("name,make", CASES, ids=[c[0] for c in CASES])
def test_gpu_matches_oracle_on_synthetic_frames(name, make, seed):
    f, refs = make(seed)
    want = run(ref_decoder.RefRecon(), f, refs)
    # odd seeds run on one long-lived context, so frame size, bit depth and chroma format change
    # under it from case to case (pool reuse); even seeds get a fresh context
    if seed & 1:
        if not _shared_gpu:
            _shared_gpu.append(vb.GpuRecon(0))
        got = run(_shared_gpu[0], f, refs)
    else:
        gpu = vb.GpuRecon(0)
        try:
            got = run(gpu, f, refs)
        finally:
            gpu.close()
    for p, (g, wv) in enumerate(zip(got, want)):
        if not np.array_equal(g, wv):
            bad = np.argwhere(g != wv)
            pytest.fail("%s seed %d plane %d: %d samples differ, first at (y,x)=%s gpu=%d oracle=%d" % (
                name, seed, p, len(bad), tuple(bad[0]), g[tuple(bad[0])], wv[tuple(bad[0])]))
