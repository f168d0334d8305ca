"""The reference tree's own vpxdec, built by tools/build_ref_vpxdec.py from /root/reference/libvpx
with integration/libvpx/vp9_b200_shim.inc in place of the fork's decode_tiles():

* VP9B_SHIM=cpu -- libvpx's own parser + its portable C DSP functions (the upstream decode_block
  flow the fork removed).  This is oracle/_ref: the reference C decoder.  It must print the golden
  MD5 lines (which were produced by libvpx v1.15.2), so reference C == libvpx 1.15.2 == our oracle.
* VP9B_SHIM=gpu -- the same binary with reconstruction handed to libvp9b200.so through
  vp9b_packer_* + vp9gpu_frame_submit: `vpxdec --md5` keeps working on the B200 engine, and the
  packed buffers come from libvpx's parser, not ours (an independent check of the packing contract).

The binary only exists where the reference tree was available at build time (it travels to the
GPU box inside oracle/_ref/)."""
import os
import subprocess

import pytest

from conftest import ROOT, clip_names, clip_path, golden_md5

import cuda_vp9_b200 as vb

VPXDEC = os.path.join(ROOT, "oracle", "_ref", "vpxdec_ref")
# the reference tree is configured without --enable-vp9-highbitdepth (the fork's HBD transform
# functions do not match their prototypes): 8-bit clips only
HBD = {"hbd10", "hbd12", "p444_hbd", "scaled_hbd", "bench_hbd1080", "p422_hbd"}
NON420 = set()
SMALL = [n for n in clip_names() if n not in HBD and not n.startswith("bench_")]
SMALL_GPU = [n for n in SMALL if n not in NON420]
ALL8 = [n for n in clip_names() if n not in HBD and n not in NON420]


def run_vpxdec(name, mode):
    env = dict(os.environ, VP9B_SHIM=mode, VP9B_LIB=os.path.join(ROOT, "cuda-vp9_b200", "libvp9b200.so"))
    r = subprocess.run([VPXDEC, "--md5", "--i420", "-o", "img-%w-%h-%4.i420", clip_path(name)], env=env,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    return [l.split()[0] for l in r.stdout.splitlines() if l.strip()]


needs_binary = pytest.mark.skipif(not os.path.exists(VPXDEC), reason="oracle/_ref/vpxdec_ref not built (no reference tree here)")


@needs_binary
@pytest.mark.parametrize("name", SMALL)
def test_reference_c_decoder_prints_the_golden_md5s(name):
    assert run_vpxdec(name, "cpu") == golden_md5(name)


FAKE = os.path.join(ROOT, "tests", "emul", "_build_fakegpu.so")


def fake_gpu_lib():
    """vp9gpu_* on top of the CPU oracle + the product packer: the shim's gpu mode without a GPU."""
    src = [os.path.join(ROOT, "tests", "emul", "fake_gpu.c"), os.path.join(ROOT, "oracle", "vp9_recon_ref.c"),
           os.path.join(ROOT, "cuda-vp9_b200", "csrc", "vp9_packer.cc"), os.path.join(ROOT, "include", "vp9_b200.h")]
    if not os.path.exists(FAKE) or os.path.getmtime(FAKE) < max(os.path.getmtime(f) for f in src):
        obj = FAKE[:-3] + ".o"
        subprocess.check_call(["gcc", "-O2", "-std=c99", "-fPIC", "-c", "-Wno-unused-function", src[0], "-o", obj])
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-I" + os.path.join(ROOT, "include"), obj, src[2],
                               "-o", FAKE])
    return FAKE


@needs_binary
@pytest.mark.parametrize("name", SMALL_GPU)
def test_libvpx_parser_through_the_packer_reproduces_the_golden_md5s(name):
    """gpu mode of the shim with vp9gpu_* served by the CPU oracle: libvpx's own parser feeds
    vp9b_packer_*, so the packing contract is checked with a parser that is not ours."""
    env = dict(os.environ, VP9B_SHIM="gpu", VP9B_LIB=fake_gpu_lib())
    r = subprocess.run([VPXDEC, "--md5", "--i420", "-o", "img-%w-%h-%4.i420", clip_path(name)], env=env,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    assert [l.split()[0] for l in r.stdout.splitlines() if l.strip()] == golden_md5(name)


@needs_binary
@pytest.mark.gpu
@pytest.mark.parametrize("name", ALL8)
def test_reference_vpxdec_md5_on_the_b200_engine(name):
    assert run_vpxdec(name, "gpu") == golden_md5(name)


def test_packer_reproduces_the_parsers_buffer():
    """Feed the packer the records of a parser-made packed frame: the GPU-visible content (sections,
    counters) must come out identical -- the packer and the parser implement one format."""
    import ctypes as C
    import numpy as np
    L = vb.lib_host()
    vp = C.c_void_p
    L.vp9b_packer_create.restype = vp
    L.vp9b_packer_destroy.argtypes = [vp]
    L.vp9b_packer_begin.argtypes = [vp, C.POINTER(vb.FrameHdr)]
    L.vp9b_packer_add_block.argtypes = [vp, vp, vp]
    L.vp9b_packer_add_txb.argtypes = [vp] + [C.c_int] * 10 + [vp, C.c_int, vp, C.c_int]
    L.vp9b_packer_set_lfm.argtypes = [vp, C.c_int, C.c_int, vp]
    L.vp9b_packer_finish.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_size_t)]
    L.vp9b_packer_last_error.restype = C.c_char_p
    L.vp9b_packer_last_error.argtypes = [vp]
    hs = C.sizeof(vb.FrameHdr)
    for clip in ("sub8", "comp_lf", "p444"):
        _, _, chunks = vb.ivf_read(clip_path(clip))
        parser = vb.HostParser()
        pk = L.vp9b_packer_create()
        n = 0
        for ch in chunks[:6]:
            for fr in vb.parse_superframe(ch):
                packed, _ = parser.parse(fr)
                if packed is None:
                    continue
                h = vb.FrameHdr.from_buffer_copy(packed[:hs])
                raw = np.frombuffer(packed, dtype=np.uint8)

                def sec(i, elem):
                    return raw[h.sec[i].offset:h.sec[i].offset + h.sec[i].count * elem].reshape(-1, elem)
                assert L.vp9b_packer_begin(pk, C.byref(h)) == 0
                blocks, submv = sec(0, 16), sec(1, 40)
                for b in blocks:
                    sub = None
                    if b[5] and b[4] < 3:
                        idx = int(np.frombuffer(b[8:12].tobytes(), dtype="<u4")[0])
                        sub = submv[idx][:32].copy()
                    bb = b.copy()
                    assert L.vp9b_packer_add_block(pk, bb.ctypes.data, sub.ctypes.data if sub is not None else None) == 0, \
                        L.vp9b_packer_last_error(pk)
                cb = h.coef_bytes
                coef = raw[h.sec[7].offset:h.sec[7].offset + h.sec[7].count * cb].view("<i2" if cb == 2 else "<i4")
                ident = np.arange(1024, dtype=np.int16)    # identity scan: "raster position" i holds coefficient i
                for si, inter in ((2, 1), (3, 0), (4, 0), (5, 0)):
                    for t in sec(si, 12):
                        off = int(np.frombuffer(t[0:4].tobytes(), dtype="<u4")[0])
                        x4, y4, eob = [int(v) for v in np.frombuffer(t[4:10].tobytes(), dtype="<u2")]
                        tx, aux = int(t[10]), int(t[11])
                        dq = np.zeros(1024, dtype="<i2" if cb == 2 else "<i4")
                        dq[:eob] = coef[off:off + eob]
                        e = L.vp9b_packer_add_txb(pk, (tx >> 4) & 3, inter, x4, y4, tx & 3, (tx >> 2) & 3, aux & 15, (aux >> 4) & 1,
                                                  (aux >> 5) & 1, (aux >> 6) & 1, dq.ctypes.data, cb, ident.ctypes.data, eob)
                        assert e == 0, L.vp9b_packer_last_error(pk)
                for i, m in enumerate(sec(8, 160)):
                    mm = m.copy()
                    assert L.vp9b_packer_set_lfm(pk, i // h.sb_cols, i % h.sb_cols, mm.ctypes.data) == 0
                out, nbytes = vp(), C.c_size_t()
                assert L.vp9b_packer_finish(pk, C.byref(out), C.byref(nbytes)) == 0
                got = C.string_at(out, nbytes.value)
                h2 = vb.FrameHdr.from_buffer_copy(got[:hs])
                g = np.frombuffer(got, dtype=np.uint8)
                for i in range(9):
                    assert h2.sec[i].count == h.sec[i].count, (clip, n, i)
                # intra records keep their order inside a superblock; coefficient offsets are re-assigned
                # in call order, so compare record fields other than coef_off, and the coefficients
                # each record points at
                coef2 = g[h2.sec[7].offset:h2.sec[7].offset + h2.sec[7].count * cb].view("<i2" if cb == 2 else "<i4")
                for si in (2, 3, 4, 5):
                    a = raw[h.sec[si].offset:h.sec[si].offset + h.sec[si].count * 12].reshape(-1, 12)
                    b = g[h2.sec[si].offset:h2.sec[si].offset + h2.sec[si].count * 12].reshape(-1, 12)
                    assert np.array_equal(a[:, 4:], b[:, 4:]), (clip, n, si)
                    for ta, tb in list(zip(a, b))[:: max(1, len(a) // 50)]:
                        oa, ob = [int(np.frombuffer(t[0:4].tobytes(), dtype="<u4")[0]) for t in (ta, tb)]
                        eob = int(np.frombuffer(ta[8:10].tobytes(), dtype="<u2")[0])
                        assert np.array_equal(coef[oa:oa + eob], coef2[ob:ob + eob])
                for si, elem in ((0, 16), (1, 40), (6, 8), (8, 160)):
                    assert np.array_equal(raw[h.sec[si].offset:h.sec[si].offset + h.sec[si].count * elem],
                                          g[h2.sec[si].offset:h2.sec[si].offset + h2.sec[si].count * elem]), (clip, n, si)
                assert list(h2.n_inter_px) == list(h.n_inter_px) and list(h2.n_compound_px) == list(h.n_compound_px)
                assert list(h2.n_txb_inter_size) == list(h.n_txb_inter_size)
                n += 1
        L.vp9b_packer_destroy(pk)
        assert n > 0
