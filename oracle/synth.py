"""TEST INFRASTRUCTURE ONLY -- seeded synthetic video content + clip encoding through the libvpx
1.15.2 encoder API (clips C4/C5).  Content: box-blurred noise texture that
translates (3,2) px per frame over a slowly changing gradient, plus a few independently moving
patches and per-frame fresh noise in one corner, so that inter (single + compound), intra, skip
and all transform sizes occur.
"""
import numpy as np

from . import vpx_oracle as vo


def _blur(a, k):
    if k <= 1:
        return a
    c = np.cumsum(np.pad(a, ((k, 0), (k, 0)), mode="wrap").astype(np.float64), axis=0)
    a = (c[k:] - c[:-k]) / k
    c = np.cumsum(a, axis=1)
    a = (c[:, k:] - c[:, :-k]) / k
    return a


def make_frames(w, h, n, seed=1234, bit_depth=8, motion=(3, 2), detail=0.04, noise_corner=True,
                temporal_noise=0.0):
    """Returns list of (Y,U,V) arrays, 4:2:0."""
    rng = np.random.default_rng(seed)
    maxv = (1 << bit_depth) - 1
    W, H = w + 256, h + 256
    tex = _blur(rng.integers(0, 256, size=(H, W)).astype(np.float64), 9)
    tex = (tex - tex.mean()) * 6.0 + 128
    fine = rng.integers(0, 256, size=(H, W)).astype(np.float64)
    tex = (1.0 - detail) * tex + detail * fine
    texu = _blur(rng.integers(0, 256, size=(H // 2, W // 2)).astype(np.float64), 7)
    texv = _blur(rng.integers(0, 256, size=(H // 2, W // 2)).astype(np.float64), 9)
    texu = (texu - texu.mean()) * 4 + 128
    texv = (texv - texv.mean()) * 4 + 128
    yy, xx = np.mgrid[0:h, 0:w]
    frames = []
    scale = (maxv + 1) / 256.0
    for t in range(n):
        dx, dy = (motion[0] * t) % 128, (motion[1] * t) % 128
        y = tex[dy:dy + h, dx:dx + w].copy()
        grad = 20.0 * np.sin((xx + 5 * t) / 97.0) + 15.0 * np.cos((yy - 3 * t) / 61.0)
        y += grad
        # an independently moving patch (creates occlusion => intra / compound)
        ps = max(16, min(w, h) // 5) & ~1
        px = int((w - ps) * (0.5 + 0.45 * np.sin(t / 7.0))) & ~1
        py = int((h - ps) * (0.5 + 0.45 * np.cos(t / 5.0))) & ~1
        y[py:py + ps, px:px + ps] = tex[H - ps - 8:H - 8, 8:8 + ps] * 0.7 + 60
        # fresh noise corner (forces intra in inter frames)
        cs = max(16, min(w, h) // 8) & ~1
        if noise_corner:
            y[:cs, w - cs:] = rng.integers(0, 256, size=(cs, cs))
        if temporal_noise > 0:
            y += rng.normal(0.0, temporal_noise, size=y.shape)
        u = texu[dy // 2:dy // 2 + (h + 1) // 2, dx // 2:dx // 2 + (w + 1) // 2].copy()
        v = texv[dy // 2:dy // 2 + (h + 1) // 2, dx // 2:dx // 2 + (w + 1) // 2].copy()
        u[py // 2:(py + ps) // 2, px // 2:(px + ps) // 2] += 30
        dt = np.uint16 if bit_depth > 8 else np.uint8
        f = [np.clip(np.rint(p * scale), 0, maxv).astype(dt) for p in (y, u, v)]
        frames.append(f)
    return frames


def encode_clip(path, w, h, n, seed=1234, bit_depth=8, scale_at=None, detail=0.04, noise_corner=True,
                deadline=1000000, temporal_noise=0.0, force_lf=False, **enc_kw):
    """Encode n synthetic frames to an IVF.  scale_at: {frame_index: (hmode, vmode)} issues
    VP8E_SET_SCALEMODE mid-stream (scaled reference frames, test/resize_test.cc:283)."""
    two_pass = enc_kw.pop("two_pass", False)
    frames = make_frames(w, h, n, seed=seed, bit_depth=bit_depth, detail=detail,
                         noise_corner=noise_corner, temporal_noise=temporal_noise)
    if two_pass:  # first pass collects stats (vpxenc.c two-pass loop)
        e1 = vo.Encoder(w, h, bit_depth=bit_depth, g_pass=1, **enc_kw)
        for f in frames:
            e1.encode(f, deadline=deadline)
        e1.flush()
        stats = b"".join(e1.stats_out)
        e1.close()
        enc = vo.Encoder(w, h, bit_depth=bit_depth, g_pass=2, stats=stats, **enc_kw)
    else:
        enc = vo.Encoder(w, h, bit_depth=bit_depth, **enc_kw)
    chunks = []
    for i, f in enumerate(frames):
        if scale_at and i in scale_at:
            enc.set_scalemode(*scale_at[i])
        chunks += enc.encode(f, deadline=deadline)
    chunks += enc.flush()
    enc.close()
    if force_lf:
        chunks = force_loop_filter(chunks)
    vo.ivf_write(path, chunks, w, h)
    return chunks


# ----------------------------------------------------------------------------------------------
# Bitstream editing: libvpx's good-quality encoder searches the loop-filter level against the
# source and settles on 0 for this synthetic content, which would leave the loop filter (half of
# the hot path) idle in clips that also have alt-ref/compound prediction.  Rewriting the 6-bit
# filter_level field of the uncompressed header (vp9_decodeframe.c:1610-1612) yields a perfectly
# valid VP9 stream whose reference output is simply whatever libvpx decodes it to.
class _Bits:
    def __init__(self, data):
        self.d, self.pos = data, 0

    def bit(self):
        b = (self.d[self.pos >> 3] >> (7 - (self.pos & 7))) & 1
        self.pos += 1
        return b

    def lit(self, n):
        v = 0
        for _ in range(n):
            v = (v << 1) | self.bit()
        return v


def _lf_field(frame, state):
    """Returns (bit position of filter_level, current level, base_qindex) or None."""
    r = _Bits(frame)
    if r.lit(2) != 2:
        return None
    profile = r.bit() | (r.bit() << 1)
    if profile == 3:
        r.bit()
    if r.bit():  # show_existing_frame
        return None
    frame_type, show_frame, error_res = r.bit(), r.bit(), r.bit()

    def colorspace():
        if profile >= 2:
            r.bit()
        cs = r.lit(3)
        if cs != 7:
            r.bit()
            if profile in (1, 3):
                r.lit(3)
        elif profile in (1, 3):
            r.bit()

    if frame_type == 0:
        r.lit(24)
        colorspace()
        r.lit(32)
        if r.bit():
            r.lit(32)
    else:
        intra_only = 0 if show_frame else r.bit()
        if not error_res:
            r.lit(2)
        if intra_only:
            r.lit(24)
            if profile > 0:
                colorspace()
            r.lit(8)
            r.lit(32)
            if r.bit():
                r.lit(32)
        else:
            r.lit(8)
            for _ in range(3):
                r.lit(4)
            found = 0
            for _ in range(3):
                if r.bit():
                    found = 1
                    break
            if not found:
                r.lit(32)
            if r.bit():
                r.lit(32)
            r.bit()
            if not r.bit():
                r.lit(2)
    if not error_res:
        r.lit(2)
    r.lit(2)
    pos = r.pos
    level = r.lit(6)
    r.lit(3)
    if r.bit():
        if r.bit():
            for _ in range(6):
                if r.bit():
                    r.lit(7)
    q = r.lit(8)
    return pos, level, q


def force_loop_filter(chunks, level_fn=None):
    """Set a non-zero loop-filter level in every frame whose level is 0 (lossless frames kept)."""
    import cuda_vp9_b200 as vb
    if level_fn is None:
        level_fn = lambda q: max(1, min(63, (q + 6) // 5))  # noqa: E731
    out = []
    for ch in chunks:
        frames = vb.parse_superframe(ch)
        buf = bytearray(ch)
        off = 0
        for fr in frames:
            f = _lf_field(fr, None)
            if f is not None:
                pos, level, q = f
                if level == 0 and q > 0:
                    nl = level_fn(q)
                    for k in range(6):
                        bitpos = off * 8 + pos + k
                        bit = (nl >> (5 - k)) & 1
                        byte, sh = bitpos >> 3, 7 - (bitpos & 7)
                        buf[byte] = (buf[byte] & ~(1 << sh)) | (bit << sh)
            off += len(fr)
        out.append(bytes(buf))
    return out
