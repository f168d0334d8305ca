"""Output scaling (vpxdec --scale -> libyuv I420Scale with kFilterBox, vpxdec.c:910-949).

Oracle: the reference tree's own libyuv, compiled from /root/reference/libvpx/third_party/libyuv by
`__graft_entry__.build()` into oracle/_ref/libyuv_scale_ref.so (C paths, LIBYUV_DISABLE_X86).
* CPU: a numpy restatement of the two bilinear paths the engine implements is compared with libyuv on
  random planes and size pairs (pins the step / rounding rules without a GPU);
* GPU: frames of the resolution-changing clips scaled by vp9gpu_get_frame_scaled_async == libyuv's
  scaling of the same (bit-exact) frames."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import ROOT, clip_path

import cuda_vp9_b200 as vb

LIB = os.path.join(ROOT, "oracle", "_ref", "libyuv_scale_ref.so")
needs_ref = pytest.mark.skipif(not os.path.exists(LIB), reason="oracle/_ref/libyuv_scale_ref.so not built (no reference tree here)")
K_FILTER_BOX = 3


def libyuv_i420_scale(planes, dw, dh):
    L = C.CDLL(LIB)
    hbd = planes[0].dtype == np.uint16
    fn = L.I420Scale_16 if hbd else L.I420Scale
    src = [np.ascontiguousarray(p) for p in planes]
    sh, sw = src[0].shape
    dt = src[0].dtype
    dst = [np.zeros((dh, dw), dt), np.zeros(((dh + 1) // 2, (dw + 1) // 2), dt), np.zeros(((dh + 1) // 2, (dw + 1) // 2), dt)]
    args = []
    for p in src:
        args += [C.c_void_p(p.ctypes.data), C.c_int(p.shape[1])]
    args += [C.c_int(sw), C.c_int(sh)]
    for p in dst:
        args += [C.c_void_p(p.ctypes.data), C.c_int(p.shape[1])]
    args += [C.c_int(dw), C.c_int(dh), C.c_int(K_FILTER_BOX)]
    assert fn(*args) == 0
    return dst


def scale_plan(sw, sh, dw, dh):
    """ScaleFilterReduce + ScalePlane dispatch + ScaleSlope (scale_common.cc:1162-1288): None = a box path."""
    if sw < 2 or not (dw * 2 >= sw and dh * 2 >= sh):
        return None
    linear = sh == 1 or dh == sh or dh * 3 == sh
    if (linear and (dw == sw or dw * 3 == sw)) or dw == sw:
        return None
    if dw <= sw and dh <= sh and ((4 * dw == 3 * sw and 4 * dh == 3 * sh) or (2 * dw == sw and 2 * dh == sh) or
                                   (8 * dw == 3 * sw and 8 * dh == 3 * sh)):
        return None
    fd = lambda n, d: (n << 16) // d                          # noqa: E731
    fd1 = lambda n, d: ((n << 16) - 0x00010001) // (d - 1)    # noqa: E731
    dx, x0 = (fd(sw, dw), (fd(sw, dw) >> 1) - 32768) if dw <= sw else (fd1(sw, dw), 0)
    if linear:
        dy = fd(sh, dh)
        y0 = dy >> 1
    elif dh <= sh:
        dy = fd(sh, dh)
        y0 = (dy >> 1) - 32768
    else:
        dy, y0 = fd1(sh, dh), 0
    return x0, dx, y0, dy, dh > sh, linear


def scale_plane_np(src, dw, dh):
    """The engine's kernel (k_scale_plane) in numpy."""
    sh, sw = src.shape
    plan = scale_plan(sw, sh, dw, dh)
    assert plan is not None
    x0, dx, y0, dy, up, linear = plan
    hbd = src.dtype == np.uint16
    s = np.pad(src.astype(np.int64), ((0, 1), (0, 1)), mode="edge")
    y = np.minimum(y0 + np.arange(dh, dtype=np.int64) * dy, (sh - 1) << 16)
    x = x0 + np.arange(dw, dtype=np.int64) * dx
    yi, yf = (y >> 16)[:, None], (np.zeros_like(y) if linear else (y >> 8) & 255)[:, None]
    xi, f = (x >> 16)[None, :], (x & 0xffff)[None, :]

    def hb(a, b):
        return a + ((((f >> 9) * (b - a)) + 0x40) >> 7) if not hbd else a + (((f * (b - a)) + 0x8000) >> 16)

    def vblend(a, b):
        g = (a * (256 - yf) + b * yf + (0 if hbd else 128)) >> 8
        g = np.where(yf == 128, (a + b + 1) >> 1, g)
        return np.where(yf == 0, a, g)
    a0, b0, a1, b1 = s[yi, xi], s[yi, xi + 1], s[yi + 1, xi], s[yi + 1, xi + 1]
    out = vblend(hb(a0, b0), hb(a1, b1)) if up else hb(vblend(a0, a1), vblend(b0, b1))
    return out.astype(src.dtype)


@needs_ref
@pytest.mark.parametrize("bd", [8, 10])
def test_bilinear_restatement_equals_libyuv(bd):
    rng = np.random.default_rng(bd)
    dt = np.uint16 if bd > 8 else np.uint8
    pairs = [((352, 288), (440, 360)), ((282, 230), (352, 288)), ((212, 173), (352, 288)), ((440, 360), (352, 288)),
             ((200, 136), (250, 170)), ((176, 144), (352, 288)), ((352, 288), (200, 150)), ((96, 64), (97, 65)),
             ((130, 98), (131, 99)), ((3072 // 8, 1728 // 8), (3840 // 8, 2160 // 8))]
    done = 0
    for (sw, sh), (dw, dh) in pairs:
        planes = [rng.integers(0, 1 << bd, size=(sh, sw)).astype(dt), rng.integers(0, 1 << bd, size=((sh + 1) // 2, (sw + 1) // 2)).astype(dt),
                  rng.integers(0, 1 << bd, size=((sh + 1) // 2, (sw + 1) // 2)).astype(dt)]
        if any(scale_plan(p.shape[1], p.shape[0], d[0], d[1]) is None for p, d in
               zip(planes, [(dw, dh), ((dw + 1) // 2, (dh + 1) // 2), ((dw + 1) // 2, (dh + 1) // 2)])):
            continue
        want = libyuv_i420_scale(planes, dw, dh)
        for p, w in zip(planes, want):
            got = scale_plane_np(p, w.shape[1], w.shape[0])
            assert np.array_equal(got, w), ((sw, sh), (dw, dh), np.argwhere(got != w)[:3])
        done += 1
    assert done >= 7


@needs_ref
@pytest.mark.gpu
@pytest.mark.parametrize("name", ["scaled_rt", "scaled", "scaled_hbd", "bench_4k"])
def test_gpu_output_scaling_equals_libyuv(name):
    w0, h0, chunks = vb.ivf_read(clip_path(name))
    parser, gpu = vb.HostParser(), vb.GpuRecon(0)
    scaled = 0
    for ch in chunks[:24]:
        for fr in vb.parse_superframe(ch):
            packed, show = parser.parse(fr)
            if packed is not None:
                gpu.submit(packed)
            if show < 0:
                continue
            w, h, bd, sx, sy = parser.slot_info(show)
            if (w, h) == (w0, h0):
                continue
            full = gpu.get_frame(show, w, h, bd, sx, sy)
            got = gpu.get_frame_scaled(show, w0, h0, bd)
            want = libyuv_i420_scale(full, w0, h0)
            for a, b in zip(got, want):
                assert np.array_equal(a, b), (name, (w, h), np.argwhere(a != b)[:3])
            scaled += 1
    assert scaled > 0
    # a ratio libyuv gives to a box kernel is refused, not approximated
    w, h, bd, sx, sy = parser.slot_info(show)
    with pytest.raises(vb.VP9BError):
        gpu.get_frame_scaled(show, w // 2, h // 2, bd)
    gpu.close()
    parser.close()
