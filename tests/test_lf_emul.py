"""The loop-filter kernel's decomposition, checked on the CPU: tests/emul/lf_emul.cc compiles the
kernel's own per-thread phase functions (vp9_lf_sb.cuh) for the host and runs them CTA by CTA in
adversarial but legal schedules; the result must equal the oracle's raster-order loop filter
(vp9_loopfilter.c:1440-1467) sample for sample, on real streams and on random masks."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import cuda_vp9_b200 as vb
from oracle import ref_decoder

import synth_packed as sp
from conftest import ROOT, clip_path

_LIB = None
# schedule bit 30: the threads of every phase run in reverse index order (the second thread of a luma
# row / column before the first one): a split of a row between two threads that is not independent
# would show
REV = 1 << 30


def emu():
    global _LIB
    if _LIB is None:
        src = os.path.join(ROOT, "tests", "emul", "lf_emul.cc")
        hdr = os.path.join(ROOT, "cuda-vp9_b200", "csrc", "vp9_lf_sb.cuh")
        out = os.path.join(ROOT, "tests", "emul", "_build_lfemu.so")
        if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wno-unknown-pragmas",
                                   "-I" + os.path.join(ROOT, "include"), src, "-o", out])
        L = C.CDLL(out)
        L.lfemu_frame.argtypes = [C.c_char_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                  C.c_uint, C.POINTER(C.c_int)]
        _LIB = L
    return _LIB


_LIB_MC = None


def emu_multicall():
    """The same emulator compiled with -DVP9B_LF_MULTICALL: the A/B variant of the edge walks (one
    inlined copy of the filter per size instead of one per walk) stays bit-exact too."""
    global _LIB_MC
    if _LIB_MC is None:
        src = os.path.join(ROOT, "tests", "emul", "lf_emul.cc")
        hdr = os.path.join(ROOT, "cuda-vp9_b200", "csrc", "vp9_lf_sb.cuh")
        out = os.path.join(ROOT, "tests", "emul", "_build_lfemu_mc.so")
        if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wno-unknown-pragmas", "-DVP9B_LF_MULTICALL",
                                   "-I" + os.path.join(ROOT, "include"), src, "-o", out])
        L = C.CDLL(out)
        L.lfemu_frame.argtypes = emu().lfemu_frame.argtypes
        _LIB_MC = L
    return _LIB_MC


def raw_planes(rec, slot):
    L = ref_decoder.lib()
    L.vp9ref_plane.restype = C.POINTER(C.c_uint16)
    L.vp9ref_plane.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    out = []
    for p in range(3):
        st, rows = C.c_int(), C.c_int()
        ptr = L.vp9ref_plane(rec.c, slot, p, C.byref(st), C.byref(rows))
        out.append(np.ctypeslib.as_array(ptr, shape=(rows.value, st.value)).copy())
    return out


def check_frame(packed, pre, post, schedules, tag, emu_fn=None):
    """pre / post: oracle planes before / after the loop filter (raw geometry)."""
    hdr = vb.FrameHdr.from_buffer_copy(packed[:C.sizeof(vb.FrameHdr)])
    chains = 0
    for sched in schedules:
        pl = [p.copy() for p in pre]
        strides = (C.c_int * 3)(*[p.shape[1] for p in pl])
        rows = (C.c_int * 3)(*[p.shape[0] for p in pl])
        nch = C.c_int()
        e = (emu_fn or emu().lfemu_frame)(packed, pl[0].ctypes.data, pl[1].ctypes.data, pl[2].ctypes.data, strides, rows, sched,
                              C.byref(nch))
        assert e == 0
        chains = max(chains, nch.value)
        for p in range(3):
            s = (hdr.ss_x if p else 0)
            ah, aw = (hdr.mi_rows * 8) >> (hdr.ss_y if p else 0), (hdr.mi_cols * 8) >> s
            a, b = pl[p][:ah, :aw], post[p][:ah, :aw]
            if not np.array_equal(a, b):
                bad = np.argwhere(a != b)
                pytest.fail("%s schedule %d plane %d: %d samples differ, first (y,x)=%s emu=%d oracle=%d" % (
                    tag, sched, p, len(bad), tuple(bad[0]), a[tuple(bad[0])], b[tuple(bad[0])]))
    return chains


def run_stream(name, max_frames=4, emu_fn=None):
    _, _, chunks = vb.ivf_read(clip_path(name))
    parser = vb.HostParser()
    a, b = ref_decoder.RefRecon(), ref_decoder.RefRecon()   # a: stops before the loop filter, b: complete
    n = 0
    for ch in chunks:
        for fr in vb.parse_superframe(ch):
            packed, _ = parser.parse(fr)
            if packed is None:
                continue
            hdr = vb.FrameHdr.from_buffer_copy(packed[:C.sizeof(vb.FrameHdr)])
            # reference frames of `a` must be the filtered ones: copy them over from b first
            b.submit(packed)
            a_slot_refs = [hdr.ref[i].slot for i in range(3)] if not hdr.intra_only else []
            for s in set(a_slot_refs):
                w, h, bd, sx, sy = parser.slot_info(s)
                a.set_frame(s, b.get_frame(s, w, h, bd, sx, sy), bd, sx, sy)
            a.submit(packed, stop_stage=3)
            if hdr.lf_level:
                check_frame(packed, raw_planes(a, hdr.dst_slot), raw_planes(b, hdr.dst_slot), (0, 1, 7, REV | 0, REV | 5), "%s frame %d" % (name, n), emu_fn)
                n += 1
            if n >= max_frames:
                return n
    return n


@pytest.mark.parametrize("name", ["odd_rt", "tiles_rt", "comp_lf", "hbd12", "scaled_hbd", "p444", "p444_hbd", "scaled_rt", "bench_1440p"])
def test_emulated_kernel_equals_oracle_on_streams(name):
    assert run_stream(name) > 0


def sparse_masks(rng, f, density):
    """Thin out the dense random masks: real streams have few inner 4x4 edges, so most superblocks
    take the no-chain path of the kernel."""
    for m in f.lfm:
        for name in ("left_y", "above_y", "left_uv", "above_uv"):
            for size in range(3):
                keep = 0
                for b in range(64):
                    if rng.random() < density:
                        keep |= 1 << b
                m[name][size] = int(m[name][size]) & keep
        m["int_4x4_y"] = int(m["int_4x4_y"]) & int(rng.integers(0, 1 << 63)) & (int(rng.integers(0, 1 << 63)) if density < 0.5 else ~0)
        m["int_4x4_uv"] = int(m["int_4x4_uv"]) & int(rng.integers(0, 1 << 16))


@pytest.mark.parametrize("bd,ss,w,h", [(8, 1, 328, 200), (8, 1, 520, 136), (10, 1, 200, 264), (8, 0, 200, 136), (12, 0, 136, 72)])
@pytest.mark.parametrize("density", [2.0, 1.0, 0.5, 0.15])
def test_emulated_kernel_equals_oracle_on_random_masks(bd, ss, w, h, density):
    random_mask_case(bd, ss, w, h, density)


def random_mask_case(bd, ss, w, h, density, emu_fn=None):
    chains, changed = [], 0
    for seed in range(3):
        rng = np.random.default_rng(1000 * seed + int(density * 100) + bd + w)
        f = sp.Frame(w, h, bd=bd, ss=ss, dst_slot=3, intra_only=True)
        sp.add_loop_filter(rng, f)
        if seed:
            f.lf_sharp = 0          # wide limits: most filters really modify samples
        if density < 1.0:
            sparse_masks(rng, f, density)
        if density > 1.0:           # every 8-aligned and every inner edge, all 4 wide: one long chain per row
            for sr in range(f.sb_rows):
                for sc in range(f.sb_cols):
                    m = f.lfm[sr * f.sb_cols + sc]
                    rows, cols = min(8, f.mi_rows - sr * 8), min(8, f.mi_cols - sc * 8)
                    full_y = sum(((1 << cols) - 1) << (8 * r) for r in range(rows))
                    full_uv = sum(((1 << ((cols + 1) >> 1)) - 1) << (4 * r) for r in range((rows + 1) >> 1))
                    int_uv = sum(((1 << (cols >> 1)) - 1) << (4 * r) for r in range(rows >> 1))
                    m["left_y"][0], m["left_y"][1], m["left_y"][2] = full_y & ~(0x0101010101010101 if sc == 0 else 0), 0, 0
                    m["left_uv"][0], m["left_uv"][1], m["left_uv"][2] = full_uv & ~(0x1111 if sc == 0 else 0), 0, 0
                    m["int_4x4_y"], m["int_4x4_uv"] = full_y, int_uv
        packed = f.pack()
        a, b = ref_decoder.RefRecon(), ref_decoder.RefRecon()
        planes = sp.random_planes(rng, (w + 7) & ~7, (h + 7) & ~7, bd, ss, smooth=True)
        # the frame content comes from set_frame on a frame of the aligned size; the packed frame has
        # no blocks, so reconstruction leaves it untouched and only the loop filter runs
        for r in (a, b):
            r.set_frame(3, planes, bd, ss, ss)
        hdr_w, hdr_h = (w + 7) & ~7, (h + 7) & ~7
        f2 = sp.Frame(hdr_w, hdr_h, bd=bd, ss=ss, dst_slot=3, intra_only=True)
        f2.lfm, f2.lf_level, f2.lf_sharp = f.lfm, f.lf_level, f.lf_sharp
        packed = f2.pack()
        b.submit(packed)
        pre = raw_planes(a, 3)
        post = raw_planes(b, 3)
        changed += sum(int((x != y).sum()) for x, y in zip(pre, post))
        chains.append(check_frame(packed, pre, post, (0, 1, 11, 12, REV | 0, REV | 13), "random bd%d ss%d %dx%d d%.2f seed %d" % (bd, ss, w, h, density, seed), emu_fn))
    assert changed > 1000            # the filters really did something
    if density > 1.0:
        assert max(chains) > 0      # all-4x4 masks: every superblock takes the serialised (chain) path


@pytest.mark.parametrize("bd,ss,w,h,density", [(8, 1, 328, 200, 1.0), (10, 1, 200, 264, 0.5), (8, 0, 200, 136, 2.0)])
def test_multicall_variant_equals_oracle(bd, ss, w, h, density):
    random_mask_case(bd, ss, w, h, density, emu_fn=emu_multicall().lfemu_frame)
