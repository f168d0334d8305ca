"""The warp-per-column loop filter (vp9_lf_warp.cuh, 8-bit 4:2:0) checked on the CPU:
* its packed (u8x4 / 16x2) filter arithmetic against the oracle's scalar filter on random and on
  adversarial tap patterns, for every filter size;
* the whole decomposition (tests/emul/lfw_emul.cc runs the kernel's per-lane phase functions worker
  by worker in adversarial but legal schedules) against the oracle's raster-order loop filter
  (vp9_loopfilter.c:1440-1467), on real streams and on random masks."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import ref_decoder

import synth_packed as sp
import test_lf_emul as base
from conftest import ROOT

_LIB = None


def emu():
    global _LIB
    if _LIB is None:
        src = os.path.join(ROOT, "tests", "emul", "lfw_emul.cc")
        hdrs = [os.path.join(ROOT, "cuda-vp9_b200", "csrc", f) for f in ("vp9_lf_warp.cuh", "vp9_lf_sb.cuh")]
        out = os.path.join(ROOT, "tests", "emul", "_build_lfwemu.so")
        if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(f) for f in [src] + hdrs):
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wno-unknown-pragmas",
                                   "-I" + os.path.join(ROOT, "include"), src, "-o", out])
        L = C.CDLL(out)
        L.lfwemu_frame.argtypes = [C.c_char_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                   C.c_uint, C.POINTER(C.c_int)]
        L.lfw_edge_test.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int]
        _LIB = L
    return _LIB


def oracle_edge(taps, size, blimit, limit, thresh):
    """taps: (16, n) uint8, p7..q7 of n positions -> filtered copy (scalar reference filter)."""
    L = ref_decoder.lib()
    L.vp9ref_lpf_edge8.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
    n = taps.shape[1]
    assert n % 8 == 0
    buf = np.ascontiguousarray(taps.astype(np.uint16))      # rows = taps, columns = positions: a horizontal edge
    for x in range(0, n, 8):
        L.vp9ref_lpf_edge8(buf.ctypes.data + 2 * (8 * n + x), n, 0, size, blimit, limit, thresh, 8)
    return buf.astype(np.uint8)


def packed_edge(taps, size, blimit, limit, thresh):
    out = taps.copy()
    for x in range(0, taps.shape[1], 4):
        t = np.ascontiguousarray(out[:, x:x + 4])
        emu().lfw_edge_test(size, t.ctypes.data, blimit, limit, thresh)
        out[:, x:x + 4] = t
    return out


def tap_patterns(rng, n):
    """Columns of 16 taps: noise, smooth ramps (flat / flat2 regions), steps across the edge, extremes."""
    cols = []
    for i in range(n):
        kind = i % 6
        if kind == 0:
            c = rng.integers(0, 256, 16)
        elif kind == 1:        # nearly constant: flat2
            c = rng.integers(0, 256) + rng.integers(-1, 2, 16)
        elif kind == 2:        # flat inside, rough outside: flat but not flat2
            c = rng.integers(2, 254) + rng.integers(-1, 2, 16)
            c[:4] += rng.integers(-40, 41, 4)
            c[12:] += rng.integers(-40, 41, 4)
        elif kind == 3:        # a step across the edge on two smooth sides
            a, b = rng.integers(0, 256, 2)
            c = np.concatenate([a + rng.integers(-2, 3, 8), b + rng.integers(-2, 3, 8)])
        elif kind == 4:        # small gradients: mask passes, hev varies
            c = rng.integers(0, 256) + np.cumsum(rng.integers(-3, 4, 16))
        else:                  # saturating values
            c = rng.choice([0, 1, 2, 127, 128, 129, 253, 254, 255], 16)
        cols.append(np.clip(c, 0, 255))
    return np.stack(cols, axis=1).astype(np.uint8)


@pytest.mark.parametrize("size", [4, 8, 16])
def test_packed_filter_arithmetic_equals_scalar_reference(size):
    rng = np.random.default_rng(size)
    total = changed = 0
    for it in range(60):
        taps = tap_patterns(rng, 512)
        level = int(rng.integers(1, 64))
        sharp = int(rng.integers(0, 8))
        lim = level >> ((sharp > 0) + (sharp > 4))
        if sharp > 0:
            lim = min(lim, 9 - sharp)
        lim = max(lim, 1)
        blimit, thresh = 2 * (level + 2) + lim, level >> 4
        want = oracle_edge(taps, size, blimit, lim, thresh)
        got = packed_edge(taps, size, blimit, lim, thresh)
        if not np.array_equal(want, got):
            bad = np.argwhere(want != got)
            x = bad[0][1]
            pytest.fail("size %d blimit %d limit %d thresh %d position %d: taps %s -> packed %s, reference %s" %
                        (size, blimit, lim, thresh, x, taps[:, x], got[:, x], want[:, x]))
        total += taps.shape[1]
        changed += int((want != taps).any(axis=0).sum())
    assert changed > total // 10     # the filters really did something


def check_frame(packed, pre, post, schedules, tag):
    return base.check_frame(packed, pre, post, schedules, tag, emu_fn=emu().lfwemu_frame)


@pytest.mark.parametrize("name", ["odd_rt", "tiles_rt", "comp_lf", "scaled_rt", "aq", "bench_1440p"])
def test_emulated_warp_kernel_equals_oracle_on_streams(name):
    assert base.run_stream(name, emu_fn=emu().lfwemu_frame) > 0


@pytest.mark.parametrize("w,h", [(328, 200), (520, 136), (200, 264)])
@pytest.mark.parametrize("density", [2.0, 1.0, 0.5, 0.15])
def test_emulated_warp_kernel_equals_oracle_on_random_masks(w, h, density):
    base.random_mask_case(8, 1, w, h, density, emu_fn=emu().lfwemu_frame)
