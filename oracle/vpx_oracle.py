"""TEST INFRASTRUCTURE ONLY -- ctypes harness around the prebuilt libvpx v1.15.2 that ships
inside the OpenCV wheel of this image (same image on the GPU box).

VP9 decoding is normative, so any conformant decoder (here: upstream libvpx, of which the
reference RomBik17/CUDA-VP9 is a fork of v1.9.0) is a bit-exact oracle for whole-frame output.
This module is used by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs only.  The product path never imports it.

What it mirrors from the reference:
  * decode loop + per-frame MD5: libvpx/vpxdec.c:833-1042, update_image_md5 vpxdec.c:285-302
  * IVF container: libvpx/ivfdec.c:40-110, libvpx/ivfenc.c
  * encoder driving (for synthetic clips): libvpx/vpxenc.c, test/resize_test.cc:349-398
"""
import ctypes as C
import glob
import hashlib
import os
import struct
import sys

import numpy as np

_LIB = None


def find_libvpx():
    import importlib.util
    cands = []
    for p in sys.path + [os.path.dirname(os.__file__) + "/site-packages"]:
        cands += glob.glob(os.path.join(p, "opencv_python_headless.libs", "libvpx*.so*"))
    spec = importlib.util.find_spec("cv2")
    if spec and spec.origin:
        root = os.path.dirname(os.path.dirname(spec.origin))
        cands += glob.glob(os.path.join(root, "opencv_python_headless.libs", "libvpx*.so*"))
    return cands[0] if cands else None


def lib():
    global _LIB
    if _LIB is None:
        path = find_libvpx()
        if path is None:
            raise RuntimeError("libvpx oracle (.so inside the OpenCV wheel) not found")
        L = C.CDLL(path)
        L.vpx_codec_version_str.restype = C.c_char_p
        L.vpx_codec_vp9_dx.restype = C.c_void_p
        L.vpx_codec_vp9_cx.restype = C.c_void_p
        L.vpx_codec_get_frame.restype = C.c_void_p
        L.vpx_codec_get_cx_data.restype = C.c_void_p
        L.vpx_codec_error.restype = C.c_char_p
        L.vpx_codec_error_detail.restype = C.c_char_p
        L.vpx_img_alloc.restype = C.c_void_p
        _LIB = L
    return _LIB


def available():
    try:
        lib()
        return True
    except Exception:
        return False


class VpxImage(C.Structure):
    # vpx/vpx_image.h (field order unchanged between v1.9 and v1.15)
    _fields_ = [("fmt", C.c_int), ("cs", C.c_int), ("range", C.c_int),
                ("w", C.c_uint), ("h", C.c_uint), ("bit_depth", C.c_uint),
                ("d_w", C.c_uint), ("d_h", C.c_uint), ("r_w", C.c_uint), ("r_h", C.c_uint),
                ("x_chroma_shift", C.c_uint), ("y_chroma_shift", C.c_uint),
                ("planes", C.c_void_p * 4), ("stride", C.c_int * 4), ("bps", C.c_int),
                ("user_priv", C.c_void_p), ("img_data", C.c_void_p),
                ("img_data_owner", C.c_int), ("self_allocd", C.c_int), ("fb_priv", C.c_void_p)]


VPX_IMG_FMT_I420 = 0x102
VPX_IMG_FMT_HIGHBITDEPTH = 0x800
VPX_IMG_FMT_I42016 = VPX_IMG_FMT_I420 | VPX_IMG_FMT_HIGHBITDEPTH
VPX_IMG_FMT_I444 = 0x106
VPX_IMG_FMT_I422 = 0x105
VPX_IMG_FMT_I440 = 0x107

# control ids (vpx/vp8dx.h:58-146, vpx/vp8cx.h:131-720)
VP9_SET_SKIP_LOOP_FILTER = 265
VP8E_SET_SCALEMODE = 11
VP8E_SET_CPUUSED = 13
VP8E_SET_ENABLEAUTOALTREF = 14
VP8E_SET_SHARPNESS = 16
VP8E_SET_STATIC_THRESHOLD = 17
VP8E_SET_ARNR_MAXFRAMES = 21
VP8E_SET_ARNR_STRENGTH = 22
VP8E_SET_CQ_LEVEL = 25
VP9E_SET_LOSSLESS = 32
VP9E_SET_TILE_COLUMNS = 33
VP9E_SET_TILE_ROWS = 34
VP9E_SET_FRAME_PARALLEL_DECODING = 35
VP9E_SET_AQ_MODE = 36


def plane_arrays(img):
    """Return cropped (Y,U,V) numpy arrays (copies) of a vpx_image_t."""
    hbd = bool(img.fmt & VPX_IMG_FMT_HIGHBITDEPTH)
    bps = 2 if hbd else 1
    out = []
    for p in range(3):
        w = img.d_w if p == 0 else (img.d_w + img.x_chroma_shift) >> img.x_chroma_shift
        h = img.d_h if p == 0 else (img.d_h + img.y_chroma_shift) >> img.y_chroma_shift
        stride = img.stride[p]
        buf = (C.c_uint8 * (stride * h)).from_address(img.planes[p])
        a = np.frombuffer(buf, dtype=np.uint8).reshape(h, stride)[:, : w * bps]
        a = a.copy()
        if hbd:
            a = a.view("<u2")
        out.append(a)
    return out


def md5_planes(planes):
    """Per-frame MD5 exactly as vpxdec --md5 (vpxdec.c:285-302): cropped rows of Y, U, V,
    2 bytes/sample little endian for high bit depth."""
    m = hashlib.md5()
    for a in planes:
        m.update(np.ascontiguousarray(a).tobytes())
    return m.hexdigest()


class Decoder:
    """vpx_codec_dec_init / vpx_codec_decode / vpx_codec_get_frame (vpxdec.c:771,847,888)."""

    def __init__(self, threads=1, skip_loop_filter=False):
        L = lib()
        self.L = L
        self.ctx = C.create_string_buffer(256)
        cfg = (C.c_uint * 3)(threads, 0, 0)
        iface = C.c_void_p(L.vpx_codec_vp9_dx())
        err = None
        for ver in (12, 11, 13, 10, 14, 9, 15, 8):
            err = L.vpx_codec_dec_init_ver(self.ctx, iface, cfg, C.c_long(0), ver)
            if err == 0:
                break
        if err != 0:
            raise RuntimeError("vpx_codec_dec_init_ver failed: %d" % err)
        if skip_loop_filter:
            r = L.vpx_codec_control_(self.ctx, VP9_SET_SKIP_LOOP_FILTER, C.c_int(1))
            if r != 0:
                raise RuntimeError("VP9_SET_SKIP_LOOP_FILTER failed")

    def decode(self, data, want="planes"):
        """Decode one compressed chunk; returns list of frames (plane lists or md5 strings)."""
        L = self.L
        buf = (C.c_uint8 * len(data)).from_buffer_copy(data)
        r = L.vpx_codec_decode(self.ctx, buf, C.c_uint(len(data)), None, C.c_long(0))
        if r != 0:
            raise RuntimeError("vpx_codec_decode: %s / %s" % (L.vpx_codec_error(self.ctx),
                                                              L.vpx_codec_error_detail(self.ctx)))
        frames = []
        it = C.c_void_p(0)
        while True:
            p = L.vpx_codec_get_frame(self.ctx, C.byref(it))
            if not p:
                break
            if want == "none":
                frames.append(None)
                continue
            img = VpxImage.from_address(p)
            pl = plane_arrays(img)
            frames.append(md5_planes(pl) if want == "md5" else pl)
        return frames

    def close(self):
        if self.ctx is not None:
            self.L.vpx_codec_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ----------------------------------------------------------------------------- IVF
def ivf_write(path, frames, w, h, fps=30):
    with open(path, "wb") as f:
        f.write(struct.pack("<4sHH4sHHIII4x", b"DKIF", 0, 32, b"VP90", w, h, fps, 1, len(frames)))
        for i, d in enumerate(frames):
            f.write(struct.pack("<IQ", len(d), i))
            f.write(d)


def ivf_read(path):
    with open(path, "rb") as f:
        b = f.read()
    assert b[:4] == b"DKIF", "not an IVF file"
    hl = struct.unpack_from("<H", b, 6)[0]
    w, h = struct.unpack_from("<HH", b, 12)
    pos = hl
    frames = []
    while pos + 12 <= len(b):
        sz = struct.unpack_from("<I", b, pos)[0]
        frames.append(b[pos + 12: pos + 12 + sz])
        pos += 12 + sz
    return w, h, frames


def decode_ivf_md5(path, threads=1, skip_loop_filter=False):
    _, _, chunks = ivf_read(path)
    d = Decoder(threads=threads, skip_loop_filter=skip_loop_filter)
    out = []
    for c in chunks:
        out += d.decode(c, want="md5")
    d.close()
    return out


# ----------------------------------------------------------------------------- encoder
class EncCfgPrefix(C.Structure):
    # vpx/vpx_encoder.h vpx_codec_enc_cfg: only the stable leading fields are touched; the
    # library fills the rest through vpx_codec_enc_config_default.
    _fields_ = [("g_usage", C.c_uint), ("g_threads", C.c_uint), ("g_profile", C.c_uint),
                ("g_w", C.c_uint), ("g_h", C.c_uint), ("g_bit_depth", C.c_int),
                ("g_input_bit_depth", C.c_uint), ("tb_num", C.c_int), ("tb_den", C.c_int),
                ("g_error_resilient", C.c_uint), ("g_pass", C.c_int),
                ("g_lag_in_frames", C.c_uint), ("rc_dropframe_thresh", C.c_uint),
                ("rc_resize_allowed", C.c_uint), ("rc_scaled_width", C.c_uint),
                ("rc_scaled_height", C.c_uint), ("rc_resize_up_thresh", C.c_uint),
                ("rc_resize_down_thresh", C.c_uint), ("rc_end_usage", C.c_int),
                ("stats_buf", C.c_void_p), ("stats_sz", C.c_size_t),
                ("mb_buf", C.c_void_p), ("mb_sz", C.c_size_t),
                ("rc_target_bitrate", C.c_uint), ("rc_min_quantizer", C.c_uint),
                ("rc_max_quantizer", C.c_uint), ("rc_undershoot_pct", C.c_uint),
                ("rc_overshoot_pct", C.c_uint), ("rc_buf_sz", C.c_uint),
                ("rc_buf_initial_sz", C.c_uint), ("rc_buf_optimal_sz", C.c_uint),
                ("rc_2pass_vbr_bias_pct", C.c_uint), ("rc_2pass_vbr_minsection_pct", C.c_uint),
                ("rc_2pass_vbr_maxsection_pct", C.c_uint), ("rc_2pass_corpus", C.c_uint),
                ("kf_mode", C.c_int), ("kf_min_dist", C.c_uint), ("kf_max_dist", C.c_uint),
                ("_rest", C.c_uint8 * 8192)]


class Encoder:
    """One-pass VP9 encoder driven through the public API; yields compressed chunks."""

    def __init__(self, w, h, bit_depth=8, cpu_used=4, lag=0, auto_alt_ref=0, tile_cols_log2=0,
                 tile_rows_log2=0, lossless=0, cq_level=None, bitrate=None, min_q=0, max_q=63,
                 kf_max_dist=9999, error_resilient=0, frame_parallel=0, threads=4, aq_mode=0,
                 end_usage=None, resize_allowed=0, sharpness=None, g_pass=0, stats=None,
                 arnr_maxframes=None, arnr_strength=None, chroma444=False, chroma=None):
        L = lib()
        self.L, self.w, self.h, self.bd = L, w, h, bit_depth
        iface = C.c_void_p(L.vpx_codec_vp9_cx())
        cfg = EncCfgPrefix()
        if L.vpx_codec_enc_config_default(iface, C.byref(cfg), 0) != 0:
            raise RuntimeError("enc_config_default failed")
        cfg.g_w, cfg.g_h = w, h
        cfg.g_threads = threads
        cfg.tb_num, cfg.tb_den = 1, 30
        cfg.g_lag_in_frames = lag
        cfg.g_error_resilient = error_resilient
        cfg.kf_max_dist = kf_max_dist
        cfg.rc_min_quantizer, cfg.rc_max_quantizer = min_q, max_q
        cfg.rc_resize_allowed = resize_allowed
        if bit_depth > 8:
            cfg.g_profile = 2
            cfg.g_bit_depth = bit_depth
            cfg.g_input_bit_depth = bit_depth
        if chroma444:
            chroma = "444"
        if chroma:  # profile 1 (8 bit) / 3 (high bit depth): non-4:2:0 sampling ("444", "422", "440")
            cfg.g_profile = 3 if bit_depth > 8 else 1
        if end_usage is not None:
            cfg.rc_end_usage = end_usage
        elif cq_level is not None:
            cfg.rc_end_usage = 3  # VPX_Q
        if bitrate is not None:
            cfg.rc_target_bitrate = bitrate
        cfg.g_pass = g_pass  # 0 one pass, 1 first pass, 2 last pass (vpx_encoder.h vpx_enc_pass)
        self._stats = None
        if stats is not None:
            self._stats = C.create_string_buffer(stats, len(stats))
            cfg.stats_buf = C.cast(self._stats, C.c_void_p)
            cfg.stats_sz = len(stats)
        self.cfg = cfg
        self.ctx = C.create_string_buffer(256)
        flags = 0x40000 if bit_depth > 8 else 0  # VPX_CODEC_USE_HIGHBITDEPTH
        err = None
        for ver in range(5, 64):
            err = L.vpx_codec_enc_init_ver(self.ctx, iface, C.byref(cfg), C.c_long(flags), ver)
            if err != 3:  # VPX_CODEC_ABI_MISMATCH
                break
        if err != 0:
            raise RuntimeError("vpx_codec_enc_init_ver failed: %d %s" % (err, L.vpx_codec_error_detail(self.ctx)))
        self.ctrl(VP8E_SET_CPUUSED, cpu_used)
        self.ctrl(VP8E_SET_ENABLEAUTOALTREF, auto_alt_ref)
        self.ctrl(VP9E_SET_TILE_COLUMNS, tile_cols_log2)
        self.ctrl(VP9E_SET_TILE_ROWS, tile_rows_log2)
        self.ctrl(VP9E_SET_FRAME_PARALLEL_DECODING, frame_parallel)
        self.ctrl(VP9E_SET_AQ_MODE, aq_mode)
        if lossless:
            self.ctrl(VP9E_SET_LOSSLESS, 1)
        if cq_level is not None:
            self.ctrl(VP8E_SET_CQ_LEVEL, cq_level)
        if sharpness is not None:
            self.ctrl(VP8E_SET_SHARPNESS, sharpness)
        if arnr_maxframes is not None:
            self.ctrl(VP8E_SET_ARNR_MAXFRAMES, arnr_maxframes)
        if arnr_strength is not None:
            self.ctrl(VP8E_SET_ARNR_STRENGTH, arnr_strength)
        self.stats_out = []
        fmt = {"444": VPX_IMG_FMT_I444, "422": VPX_IMG_FMT_I422, "440": VPX_IMG_FMT_I440}.get(chroma, VPX_IMG_FMT_I420)
        if bit_depth > 8:
            fmt |= VPX_IMG_FMT_HIGHBITDEPTH
        self.img_p = L.vpx_img_alloc(None, fmt, w, h, 32)
        self.img = VpxImage.from_address(self.img_p)
        self.pts = 0

    def ctrl(self, cid, val):
        r = self.L.vpx_codec_control_(self.ctx, cid, C.c_int(val))
        if r != 0:
            raise RuntimeError("control %d failed: %d" % (cid, r))

    def set_scalemode(self, hmode, vmode):
        """VP8E_SET_SCALEMODE (test/resize_test.cc:283): 0 normal, 1 4/5, 2 3/5, 3 1/2."""
        m = (C.c_int * 2)(hmode, vmode)
        r = self.L.vpx_codec_control_(self.ctx, VP8E_SET_SCALEMODE, C.byref(m))
        if r != 0:
            raise RuntimeError("SET_SCALEMODE failed")

    def _fill(self, planes):
        img = self.img
        bps = 2 if self.bd > 8 else 1
        for p in range(3):
            a = np.ascontiguousarray(planes[p])
            h, wbytes = a.shape[0], a.shape[1] * bps
            stride = img.stride[p]
            dst = np.frombuffer((C.c_uint8 * (stride * h)).from_address(img.planes[p]),
                                dtype=np.uint8).reshape(h, stride)
            dst[:, :wbytes] = a.view(np.uint8).reshape(h, wbytes)

    def _drain(self):
        out = []
        it = C.c_void_p(0)
        while True:
            p = self.L.vpx_codec_get_cx_data(self.ctx, C.byref(it))
            if not p:
                break
            kind = C.c_int.from_address(p).value
            if kind == 1:  # VPX_CODEC_STATS_PKT
                buf = C.c_void_p.from_address(p + 8).value
                sz = C.c_size_t.from_address(p + 16).value
                self.stats_out.append(C.string_at(buf, sz))
                continue
            if kind != 0:
                continue
            buf = C.c_void_p.from_address(p + 8).value
            sz = C.c_size_t.from_address(p + 16).value
            out.append(C.string_at(buf, sz))
        return out

    def encode(self, planes, flags=0, deadline=1000000):
        """planes: (Y,U,V) numpy arrays or None to flush. Returns list of chunks."""
        if planes is not None:
            self._fill(planes)
            r = self.L.vpx_codec_encode(self.ctx, C.c_void_p(self.img_p), C.c_int64(self.pts),
                                        C.c_ulong(1), C.c_long(flags), C.c_ulong(deadline))
            self.pts += 1
        else:
            r = self.L.vpx_codec_encode(self.ctx, None, C.c_int64(self.pts), C.c_ulong(1),
                                        C.c_long(0), C.c_ulong(deadline))
        if r != 0:
            raise RuntimeError("vpx_codec_encode: %s" % self.L.vpx_codec_error_detail(self.ctx))
        return self._drain()

    def flush(self):
        out = []
        while True:
            c = self.encode(None)
            if not c:
                break
            out += c
        return out

    def close(self):
        self.L.vpx_codec_destroy(self.ctx)
