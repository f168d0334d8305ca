"""TEST INFRASTRUCTURE ONLY -- CPU decoder = product host parser -> packed buffer -> C oracle
(oracle/vp9_recon_ref.c).  Used to (a) pin the oracle + packing contract against libvpx MD5s
and (b) give per-stage expected frames for the CUDA kernels."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _LIB
    if _LIB is None:
        p = os.path.join(_HERE, "_build", "libvp9ref.so")
        if not os.path.exists(p):
            build()
        L = C.CDLL(p)
        vp = C.c_void_p
        L.vp9ref_create.restype = vp
        L.vp9ref_destroy.argtypes = [vp]
        L.vp9ref_recon_frame.argtypes = [vp, C.c_char_p, C.c_size_t, C.c_int]
        L.vp9ref_get_frame.argtypes = [vp, C.c_int, vp, vp, vp, C.c_int, C.c_int]
        L.vp9ref_set_frame.argtypes = [vp, C.c_int, C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_int, vp, vp, vp,
                                       C.c_int, C.c_int]
        L.vp9ref_intra_predictor.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, vp, C.c_int, C.c_int]
        L.vp9ref_inverse_transform_add.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, C.c_int, C.c_int]
        L.vp9ref_lpf_edge8.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        _LIB = L
    return _LIB


class RefRecon:
    """Mirror of cuda_vp9_b200.GpuRecon over the C oracle."""

    def __init__(self):
        self.L = lib()
        self.c = self.L.vp9ref_create()

    def submit(self, packed, stop_stage=0):
        e = self.L.vp9ref_recon_frame(self.c, packed, len(packed), stop_stage)
        if e:
            raise RuntimeError("vp9ref_recon_frame: %d" % e)

    def get_frame(self, slot, w, h, bd, sx, sy):
        dt = np.uint16 if bd > 8 else np.uint8
        shapes = [(h, w), ((h + sy) >> sy, (w + sx) >> sx), ((h + sy) >> sy, (w + sx) >> sx)]
        pl = [np.empty(s, dtype=dt) for s in shapes]
        e = self.L.vp9ref_get_frame(self.c, slot, pl[0].ctypes.data, pl[1].ctypes.data, pl[2].ctypes.data,
                                    pl[0].strides[0], pl[1].strides[0])
        if e:
            raise RuntimeError("vp9ref_get_frame: %d" % e)
        return pl

    def set_frame(self, slot, planes, bd, sx, sy):
        y, u, v = [np.ascontiguousarray(p) for p in planes]
        e = self.L.vp9ref_set_frame(self.c, slot, y.shape[1], y.shape[0], bd, sx, sy, y.ctypes.data, u.ctypes.data,
                                    v.ctypes.data, y.strides[0], u.strides[0])
        if e:
            raise RuntimeError("vp9ref_set_frame: %d" % e)

    def close(self):
        if self.c:
            self.L.vp9ref_destroy(self.c)
            self.c = None


def decode_chunks(chunks, want="md5", stop_stage=0, parse_threads=1):
    """Decode IVF chunks with (host parser -> oracle).  Yields one entry per shown frame."""
    import cuda_vp9_b200 as vb
    parser = vb.HostParser()
    if parse_threads > 1:
        parser.set_threads(parse_threads)
    rec = RefRecon()
    out = []
    for ch in chunks:
        for fr in vb.parse_superframe(ch):
            packed, show = parser.parse(fr)
            if packed is not None:
                rec.submit(packed, stop_stage)
            if show >= 0:
                w, h, bd, sx, sy = parser.slot_info(show)
                pl = rec.get_frame(show, w, h, bd, sx, sy)
                out.append(vb.md5_planes(pl) if want == "md5" else pl)
    rec.close()
    parser.close()
    return out
