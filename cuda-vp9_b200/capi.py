"""ctypes mirror of include/vp9_b200.h.  Loads the in-tree shared libraries; never falls back to
a CPU reconstruction path."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
NUM_SLOTS = 12
MAX_BATCH = 32   # VP9B_MAX_BATCH: frames of different contexts per launch (vp9gpu_frame_submit_device_multi)
SEC_COUNT = 12

(SEC_BLOCKS, SEC_SUBMV, SEC_TXB_INTER, SEC_TXB_INTRA_Y, SEC_TXB_INTRA_U, SEC_TXB_INTRA_V, SEC_SB_INTRA,
 SEC_COEF, SEC_LFM) = range(9)


class VP9BError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("vp9b error %d: %s" % (code, msg))
        self.code = code


class Section(C.Structure):
    _fields_ = [("offset", C.c_uint32), ("count", C.c_uint32)]


class RefDesc(C.Structure):
    _fields_ = [("slot", C.c_int32), ("width", C.c_uint32), ("height", C.c_uint32),
                ("x_scale_fp", C.c_int32), ("y_scale_fp", C.c_int32),
                ("x_step_q4", C.c_int32), ("y_step_q4", C.c_int32)]


class FrameHdr(C.Structure):
    _fields_ = [("magic", C.c_uint32), ("total_bytes", C.c_uint32),
                ("width", C.c_uint32), ("height", C.c_uint32),
                ("mi_cols", C.c_uint32), ("mi_rows", C.c_uint32),
                ("sb_cols", C.c_uint32), ("sb_rows", C.c_uint32),
                ("bit_depth", C.c_uint8), ("ss_x", C.c_uint8), ("ss_y", C.c_uint8), ("lossless", C.c_uint8),
                ("intra_only", C.c_uint8), ("show_frame", C.c_uint8), ("lf_level", C.c_uint8),
                ("lf_sharpness", C.c_uint8), ("coef_bytes", C.c_uint8), ("reserved", C.c_uint8 * 3),
                ("dst_slot", C.c_int32), ("ref", RefDesc * 3), ("sec", Section * SEC_COUNT),
                ("lf_lim", C.c_uint8 * 64), ("lf_mblim", C.c_uint8 * 64),
                ("frame_index", C.c_uint32), ("n_inter_px", C.c_uint32 * 3),
                ("n_compound_px", C.c_uint32 * 3), ("n_txb_inter_size", C.c_uint32 * 4),
                ("reserved2", C.c_uint32 * 4)]


class Image(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("bit_depth", C.c_int),
                ("ss_x", C.c_int), ("ss_y", C.c_int), ("slot", C.c_int)]


class FrameBuffer(C.Structure):      # vp9b_frame_buffer
    _fields_ = [("data", C.c_void_p), ("size", C.c_size_t), ("priv", C.c_void_p)]


class OutFrame(C.Structure):         # vp9b_out_frame
    _fields_ = [("img", Image), ("planes", C.c_void_p * 3), ("stride", C.c_int * 3), ("fb", FrameBuffer),
                ("ticket", C.c_uint64)]


GET_FB_CB = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_size_t, C.POINTER(FrameBuffer))
RELEASE_FB_CB = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(FrameBuffer))


_host = None
_gpu = None


def _proto_host(L):
    L.vp9b_parser_create.restype = C.c_void_p
    L.vp9b_parser_destroy.argtypes = [C.c_void_p]
    L.vp9b_parser_last_error.restype = C.c_char_p
    L.vp9b_parser_last_error.argtypes = [C.c_void_p]
    L.vp9b_parser_parse.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t, C.POINTER(C.c_void_p),
                                    C.POINTER(C.c_size_t), C.POINTER(C.c_int)]
    L.vp9b_parser_slot_info.argtypes = [C.c_void_p, C.c_int] + [C.POINTER(C.c_uint32)] * 2 + [C.POINTER(C.c_int)] * 3
    L.vp9b_parse_superframe.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(C.c_uint32)]


def lib_host():
    """libvp9b200_host.so: parser only (no CUDA).  The full library also exports these symbols."""
    global _host
    if _host is None:
        for name in ("libvp9b200_host.so", "libvp9b200.so"):
            p = os.path.join(_HERE, name)
            if os.path.exists(p):
                _host = C.CDLL(p)
                _proto_host(_host)
                break
        else:
            raise ImportError("libvp9b200_host.so not built: run `python -c 'import __graft_entry__ as g; g.build()'`")
    return _host


def lib_gpu():
    """libvp9b200.so: CUDA reconstruction + parser + decoder.  Raises if it was not built."""
    global _gpu
    if _gpu is None:
        p = os.environ.get("VP9B_LIB") or os.path.join(_HERE, "libvp9b200.so")   # VP9B_LIB: A/B builds
        if not os.path.exists(p):
            raise ImportError("libvp9b200.so (CUDA extension) not built -- there is no CPU fallback")
        # one hardware work queue per decoder context (effective only if CUDA has not started yet)
        os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
        L = C.CDLL(p)
        _proto_host(L)
        vp = C.c_void_p
        L.vp9gpu_create.restype = vp
        L.vp9gpu_create.argtypes = [C.c_int]
        L.vp9gpu_destroy.argtypes = [vp]
        L.vp9gpu_last_error.restype = C.c_char_p
        L.vp9gpu_last_error.argtypes = [vp]
        L.vp9gpu_frame_submit.argtypes = [vp, vp, C.c_size_t]
        L.vp9gpu_frame_submit_device.argtypes = [vp, vp, C.POINTER(FrameHdr)]
        L.vp9gpu_upload_packed.argtypes = [vp, vp, C.c_size_t, C.POINTER(vp)]
        L.vp9gpu_free_packed.argtypes = [vp, vp]
        L.vp9gpu_sync.argtypes = [vp]
        L.vp9gpu_frame_submit_device_n.argtypes = [vp, C.POINTER(C.c_void_p), C.POINTER(C.POINTER(FrameHdr)), C.c_int]
        L.vp9gpu_frame_submit_device_multi.argtypes = [C.POINTER(vp), C.POINTER(C.c_void_p),
                                                        C.POINTER(C.POINTER(FrameHdr)), C.c_int]
        L.vp9gpu_get_frame.argtypes = [vp, C.c_int, vp, vp, vp, C.c_int, C.c_int]
        L.vp9gpu_get_frame_async.argtypes = [vp, C.c_int, vp, vp, vp, C.c_int, C.c_int]
        L.vp9gpu_get_frame_8bit_async.argtypes = [vp, C.c_int, vp, vp, vp, C.c_int, C.c_int]
        L.vp9gpu_alloc_host.restype = vp
        L.vp9gpu_alloc_host.argtypes = [C.c_size_t]
        L.vp9gpu_free_host.argtypes = [vp]
        L.vp9b_decoder_read_frame_async.argtypes = [vp, C.POINTER(Image), vp, vp, vp, C.c_int, C.c_int]
        L.vp9b_decoder_sync.argtypes = [vp]
        L.vp9b_decoder_mark.argtypes = [vp, C.POINTER(C.c_uint64)]
        L.vp9b_decoder_wait.argtypes = [vp, C.c_uint64]
        L.vp9gpu_mark.argtypes = [vp, C.POINTER(C.c_uint64)]
        L.vp9gpu_wait.argtypes = [vp, C.c_uint64]
        L.vp9gpu_set_frame.argtypes = [vp, C.c_int, C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_int, vp, vp, vp,
                                       C.c_int, C.c_int]
        L.vp9gpu_set_stop_stage.argtypes = [vp, C.c_int]
        L.vp9gpu_stream.restype = vp
        L.vp9gpu_stream.argtypes = [vp]
        L.vp9gpu_launch_count.restype = C.c_uint64
        L.vp9gpu_launch_count.argtypes = [vp]
        L.vp9gpu_set_profiling.argtypes = [vp, C.c_int]
        L.vp9gpu_get_stage_ms.argtypes = [vp, C.POINTER(C.c_float), C.c_int]
        L.vp9b_decoder_create.restype = vp
        L.vp9b_decoder_create.argtypes = [C.c_int]
        L.vp9b_decoder_destroy.argtypes = [vp]
        L.vp9b_decoder_last_error.restype = C.c_char_p
        L.vp9b_decoder_last_error.argtypes = [vp]
        L.vp9b_decoder_decode.argtypes = [vp, C.c_char_p, C.c_size_t]
        L.vp9b_decoder_get_frame.argtypes = [vp, C.POINTER(Image)]
        L.vp9b_decoder_read_frame.argtypes = [vp, C.POINTER(Image), vp, vp, vp, C.c_int, C.c_int]
        L.vp9b_decoder_gpu.restype = vp
        L.vp9b_decoder_gpu.argtypes = [vp]
        L.vp9b_decoder_parser.restype = vp
        L.vp9b_decoder_parser.argtypes = [vp]
        _gpu = L
    return _gpu


def parse_superframe(chunk, L=None):
    """Split an IVF chunk into frames (vp9_parse_superframe_index)."""
    L = L or lib_host()
    sizes = (C.c_uint32 * 8)()
    n = L.vp9b_parse_superframe(chunk, len(chunk), sizes)
    if n < 0:
        raise VP9BError(n, "corrupt superframe index")
    out, off = [], 0
    for k in range(n):
        out.append(chunk[off: off + sizes[k]])
        off += sizes[k]
    return out


def plane_shapes(width, height, ss_x, ss_y):
    return [(height, width), ((height + ss_y) >> ss_y, (width + ss_x) >> ss_x),
            ((height + ss_y) >> ss_y, (width + ss_x) >> ss_x)]


class HostParser:
    """Bitstream -> packed per-frame buffer (include/vp9_b200.h section 2)."""

    def set_threads(self, n):
        """Tile-column worker threads per frame (vp9b_parser_set_threads)."""
        self.L.vp9b_parser_set_threads.argtypes = [C.c_void_p, C.c_int]
        if self.L.vp9b_parser_set_threads(self.p, int(n)):
            raise VP9BError(-1, "bad thread count")

    def __init__(self, L=None):
        self.L = L or lib_host()
        self.p = self.L.vp9b_parser_create()
        if not self.p:
            raise MemoryError("vp9b_parser_create")

    def parse(self, frame_bytes):
        """Returns (packed_bytes or None, show_slot)."""
        packed, nb, ss = C.c_void_p(), C.c_size_t(), C.c_int()
        e = self.L.vp9b_parser_parse(self.p, frame_bytes, len(frame_bytes), C.byref(packed), C.byref(nb), C.byref(ss))
        if e:
            raise VP9BError(e, self.L.vp9b_parser_last_error(self.p).decode())
        buf = C.string_at(packed.value, nb.value) if nb.value else None
        return buf, ss.value

    def slot_info(self, slot):
        w, h = C.c_uint32(), C.c_uint32()
        bd, sx, sy = C.c_int(), C.c_int(), C.c_int()
        e = self.L.vp9b_parser_slot_info(self.p, slot, C.byref(w), C.byref(h), C.byref(bd), C.byref(sx), C.byref(sy))
        if e:
            raise VP9BError(e, "bad slot")
        return w.value, h.value, bd.value, sx.value, sy.value

    def close(self):
        if self.p:
            self.L.vp9b_parser_destroy(self.p)
            self.p = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _alloc_planes(w, h, bd, sx, sy):
    dt = np.uint16 if bd > 8 else np.uint8
    return [np.empty(s, dtype=dt) for s in plane_shapes(w, h, sx, sy)]


class GpuRecon:
    """Reconstruction context on one GPU (include/vp9_b200.h section 1)."""

    def __init__(self, device=-1):
        self.L = lib_gpu()
        self.ctx = self.L.vp9gpu_create(device)
        if not self.ctx:
            raise VP9BError(-6, "vp9gpu_create failed (no CUDA device?)")

    def _chk(self, e):
        if e:
            raise VP9BError(e, self.L.vp9gpu_last_error(self.ctx).decode())

    def submit(self, packed):
        buf = (C.c_uint8 * len(packed)).from_buffer_copy(packed)
        self._chk(self.L.vp9gpu_frame_submit(self.ctx, buf, len(packed)))

    def upload(self, packed):
        buf = (C.c_uint8 * len(packed)).from_buffer_copy(packed)
        d = C.c_void_p()
        self._chk(self.L.vp9gpu_upload_packed(self.ctx, buf, len(packed), C.byref(d)))
        return d, FrameHdr.from_buffer_copy(packed[: C.sizeof(FrameHdr)])

    def submit_device(self, d_packed, hdr):
        self._chk(self.L.vp9gpu_frame_submit_device(self.ctx, d_packed, C.byref(hdr)))

    @staticmethod
    def make_multi(ctxs, frames):
        """One frame of each context in `ctxs` -> opaque argument pack for submit_multi.
        frames: [(device pointer, FrameHdr)], same order as ctxs."""
        n = len(ctxs)
        cs = (C.c_void_p * n)(*[c.ctx for c in ctxs])
        ptrs = (C.c_void_p * n)(*[d.value if isinstance(d, C.c_void_p) else d for d, _ in frames])
        hdrs = (C.POINTER(FrameHdr) * n)(*[C.pointer(h) for _, h in frames])
        return (cs, ptrs, hdrs, n, frames, ctxs)

    @staticmethod
    def submit_multi(pack):
        """Frames of several contexts through one launch per stage (vp9gpu_frame_submit_device_multi)."""
        lead = pack[5][0]
        lead._chk(lead.L.vp9gpu_frame_submit_device_multi(pack[0], pack[1], pack[2], pack[3]))

    def make_batch(self, frames):
        """frames: [(device pointer, FrameHdr)] -> opaque batch for submit_batch (one C call per pass)."""
        n = len(frames)
        ptrs = (C.c_void_p * n)(*[d.value if isinstance(d, C.c_void_p) else d for d, _ in frames])
        hdrs = (C.POINTER(FrameHdr) * n)(*[C.pointer(h) for _, h in frames])
        return (ptrs, hdrs, n, frames)          # keeps the headers alive

    def submit_batch(self, batch):
        self._chk(self.L.vp9gpu_frame_submit_device_n(self.ctx, batch[0], batch[1], batch[2]))

    def free(self, d_packed):
        self._chk(self.L.vp9gpu_free_packed(self.ctx, d_packed))

    def sync(self):
        self._chk(self.L.vp9gpu_sync(self.ctx))

    def set_stop_stage(self, stage):
        self._chk(self.L.vp9gpu_set_stop_stage(self.ctx, stage))

    def get_frame(self, slot, w, h, bd, sx, sy):
        pl = _alloc_planes(w, h, bd, sx, sy)
        self._chk(self.L.vp9gpu_get_frame(self.ctx, slot, pl[0].ctypes.data, pl[1].ctypes.data, pl[2].ctypes.data,
                                          pl[0].strides[0], pl[1].strides[0]))
        return pl

    def get_frame_8bit(self, slot, w, h, sx, sy):
        """Shown frame at 8 bits per sample (high-bit-depth slots are shifted down on the GPU)."""
        pl = _alloc_planes(w, h, 8, sx, sy)
        self._chk(self.L.vp9gpu_get_frame_8bit_async(self.ctx, slot, pl[0].ctypes.data, pl[1].ctypes.data,
                                                     pl[2].ctypes.data, pl[0].strides[0], pl[1].strides[0]))
        self.sync()
        return pl

    def get_frame_scaled(self, slot, dst_w, dst_h, bd):
        """Shown 4:2:0 frame scaled to dst_w x dst_h on the GPU (vpxdec --scale)."""
        pl = _alloc_planes(dst_w, dst_h, bd, 1, 1)
        self.L.vp9gpu_get_frame_scaled_async.argtypes = [C.c_void_p, C.c_int, C.c_uint32, C.c_uint32] + [C.c_void_p] * 3 + [C.c_int] * 2
        self._chk(self.L.vp9gpu_get_frame_scaled_async(self.ctx, slot, dst_w, dst_h, pl[0].ctypes.data, pl[1].ctypes.data,
                                                       pl[2].ctypes.data, pl[0].strides[0], pl[1].strides[0]))
        self.sync()
        return pl

    def set_frame(self, slot, planes, bd, sx, sy):
        y, u, v = [np.ascontiguousarray(p) for p in planes]
        self._chk(self.L.vp9gpu_set_frame(self.ctx, slot, y.shape[1], y.shape[0], bd, sx, sy, y.ctypes.data,
                                          u.ctypes.data, v.ctypes.data, y.strides[0], u.strides[0]))

    def launch_count(self):
        return int(self.L.vp9gpu_launch_count(self.ctx))

    def stream(self):
        return self.L.vp9gpu_stream(self.ctx)

    def set_profiling(self, on):
        self._chk(self.L.vp9gpu_set_profiling(self.ctx, int(on)))

    def stage_ms(self, reset=True):
        ms = (C.c_float * 8)()
        self._chk(self.L.vp9gpu_get_stage_ms(self.ctx, ms, int(reset)))
        return list(ms)

    def close(self):
        if self.ctx:
            self.L.vp9gpu_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class PinnedFrame:
    """Page-locked (Y,U,V) output planes allocated through the C ABI (vp9gpu_alloc_host)."""

    def __init__(self, w, h, bd=8, sx=1, sy=1):
        self.L = lib_gpu()
        bps = 2 if bd > 8 else 1
        shapes = plane_shapes(w, h, sx, sy)
        sizes = [s[0] * s[1] * bps for s in shapes]
        self.nbytes = sum(sizes)
        self.ptr = self.L.vp9gpu_alloc_host(self.nbytes)
        if not self.ptr:
            raise MemoryError("vp9gpu_alloc_host")
        dt = np.uint16 if bd > 8 else np.uint8
        self.planes, off = [], 0
        for s, n in zip(shapes, sizes):
            buf = (C.c_uint8 * n).from_address(self.ptr + off)
            self.planes.append(np.frombuffer(buf, dtype=dt).reshape(s))
            off += n

    def close(self):
        if self.ptr:
            self.planes = []
            self.L.vp9gpu_free_host(self.ptr)
            self.ptr = None


class Decoder:
    """Whole decoder on one GPU: vpx_codec_decode / vpx_codec_get_frame equivalents."""

    def __init__(self, device=-1):
        self.L = lib_gpu()
        self.d = self.L.vp9b_decoder_create(device)
        if not self.d:
            raise VP9BError(-6, "vp9b_decoder_create failed (no CUDA device?)")

    def _chk(self, e):
        if e < 0:
            raise VP9BError(e, self.L.vp9b_decoder_last_error(self.d).decode())
        return e

    def decode(self, chunk):
        self._chk(self.L.vp9b_decoder_decode(self.d, chunk, len(chunk)))

    def get_frame(self, read=True):
        """Returns None, or (planes | None) for the frame to show."""
        img = Image()
        r = self._chk(self.L.vp9b_decoder_get_frame(self.d, C.byref(img)))
        if r == 0:
            return None
        if not read:
            return img
        pl = _alloc_planes(img.width, img.height, img.bit_depth, img.ss_x, img.ss_y)
        self._chk(self.L.vp9b_decoder_read_frame(self.d, C.byref(img), pl[0].ctypes.data, pl[1].ctypes.data,
                                                 pl[2].ctypes.data, pl[0].strides[0], pl[1].strides[0]))
        return pl

    def get_frame_into(self, pinned, sync=False):
        """Asynchronous output into a PinnedFrame; returns False when there is no frame to show."""
        img = Image()
        r = self._chk(self.L.vp9b_decoder_get_frame(self.d, C.byref(img)))
        if r == 0:
            return False
        pl = pinned.planes
        fn = self.L.vp9b_decoder_read_frame if sync else self.L.vp9b_decoder_read_frame_async
        self._chk(fn(self.d, C.byref(img), pl[0].ctypes.data, pl[1].ctypes.data, pl[2].ctypes.data,
                     pl[0].strides[0], pl[1].strides[0]))
        return True

    def set_parse_threads(self, n):
        """Tile-column parser threads of this decoder (vp9b_parser_set_threads on its parser)."""
        self.L.vp9b_decoder_parser.restype = C.c_void_p
        self.L.vp9b_decoder_parser.argtypes = [C.c_void_p]
        self.L.vp9b_parser_set_threads.argtypes = [C.c_void_p, C.c_int]
        self._chk(self.L.vp9b_parser_set_threads(self.L.vp9b_decoder_parser(self.d), int(n)))

    def set_frame_buffer_functions(self, get, release):
        """External output buffers (vpx_codec_set_frame_buffer_functions): get(min_size) -> (address, size, tag),
        release(tag).  Must be called before the first decode()."""
        def _get(priv, min_size, fb):
            try:
                addr, size, tag = get(min_size)
            except Exception:
                return -1
            fb[0].data, fb[0].size, fb[0].priv = addr, size, tag
            return 0

        def _rel(priv, fb):
            try:
                release(fb[0].priv)
            except Exception:
                return -1
            return 0
        self._cbs = (GET_FB_CB(_get), RELEASE_FB_CB(_rel))     # keep the trampolines alive
        self.L.vp9b_decoder_set_frame_buffer_functions.argtypes = [C.c_void_p, GET_FB_CB, RELEASE_FB_CB, C.c_void_p]
        self._chk(self.L.vp9b_decoder_set_frame_buffer_functions(self.d, self._cbs[0], self._cbs[1], None))

    def get_frame_ext(self):
        """Shown frame delivered in an application buffer: None, or (OutFrame, [plane arrays]); the arrays are
        valid after wait(out.ticket) and until release_frame(out)."""
        out = OutFrame()
        self.L.vp9b_decoder_get_frame_ext.argtypes = [C.c_void_p, C.POINTER(OutFrame)]
        if self._chk(self.L.vp9b_decoder_get_frame_ext(self.d, C.byref(out))) == 0:
            return None
        im = out.img
        dt = np.uint16 if im.bit_depth > 8 else np.uint8
        planes = []
        for p, (h, w) in enumerate(plane_shapes(im.width, im.height, im.ss_x, im.ss_y)):
            st = out.stride[p]
            raw = np.frombuffer((C.c_uint8 * (st * h)).from_address(out.planes[p]), dtype=np.uint8).reshape(h, st)
            planes.append(raw.view(dt)[:, :w])
        return out, planes

    def release_frame(self, out):
        self.L.vp9b_decoder_release_frame.argtypes = [C.c_void_p, C.POINTER(OutFrame)]
        self._chk(self.L.vp9b_decoder_release_frame(self.d, C.byref(out)))

    def sync(self):
        self._chk(self.L.vp9b_decoder_sync(self.d))

    def mark(self):
        """Ticket for everything enqueued so far (e.g. the frame just handed to get_frame_into)."""
        t = C.c_uint64()
        self._chk(self.L.vp9b_decoder_mark(self.d, C.byref(t)))
        return t.value

    def wait(self, ticket):
        self._chk(self.L.vp9b_decoder_wait(self.d, ticket))

    def gpu_launch_count(self):
        return int(self.L.vp9gpu_launch_count(self.L.vp9b_decoder_gpu(self.d)))

    def close(self):
        if self.d:
            self.L.vp9b_decoder_destroy(self.d)
            self.d = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
