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
                deadline=1000000, temporal_noise=0.0, force_lf=False, show_existing_at=None, intra_only_at=None,
                force_kf_at=None, **enc_kw):
    """Encode n synthetic frames to an IVF.  scale_at: {frame_index: (hmode, vmode)} issues
    VP8E_SET_SCALEMODE mid-stream (scaled reference frames, test/resize_test.cc:283)."""
    two_pass = enc_kw.pop("two_pass", False)
    frames = make_frames(w, h, n, seed=seed, bit_depth=bit_depth, detail=detail,
                         noise_corner=noise_corner, temporal_noise=temporal_noise)
    if enc_kw.get("chroma444"):  # full-resolution chroma planes
        frames = [[f[0], np.repeat(np.repeat(f[1], 2, 0), 2, 1)[:h, :w], np.repeat(np.repeat(f[2], 2, 0), 2, 1)[:h, :w]]
                  for f in frames]
    elif enc_kw.get("chroma") == "422":  # full-height chroma
        frames = [[f[0], np.repeat(f[1], 2, 0)[:h], np.repeat(f[2], 2, 0)[:h]] for f in frames]
    elif enc_kw.get("chroma") == "440":  # full-width chroma
        frames = [[f[0], np.repeat(f[1], 2, 1)[:, :w], np.repeat(f[2], 2, 1)[:, :w]] for f in frames]
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
        flags = 1 if (force_kf_at and i in force_kf_at) else 0  # VPX_EFLAG_FORCE_KF
        chunks += enc.encode(f, flags=flags, deadline=deadline)
    chunks += enc.flush()
    enc.close()
    if force_lf:
        chunks = force_loop_filter(chunks)
    if intra_only_at:
        chunks = [keyframe_to_intra_only(c) if i in intra_only_at else c for i, c in enumerate(chunks)]
    if show_existing_at:
        out = []
        for i, c in enumerate(chunks):
            out.append(c)
            if i in show_existing_at:
                out.append(show_existing_frame(show_existing_at[i]))
        chunks = out
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


def show_existing_frame(ref_index, profile=0):
    """A one-byte (profile 0-2) frame: frame_marker 2, profile, show_existing_frame = 1, index
    (vp9_decodeframe.c:3130-3160).  The decoder outputs reference buffer `ref_index` again."""
    bits = [1, 0, profile & 1, profile >> 1]
    if profile == 3:
        bits.append(0)
    bits += [1, (ref_index >> 2) & 1, (ref_index >> 1) & 1, ref_index & 1]
    while len(bits) % 8:
        bits.append(0)
    return bytes(sum(b << (7 - j) for j, b in enumerate(bits[i:i + 8])) for i in range(0, len(bits), 8))


def keyframe_to_intra_only(chunk, refresh_mask=0xFF):
    """Rewrite the uncompressed header of a profile-0 key frame as an intra-only frame
    (vp9_decodeframe.c:3162-3260).  Both use the key-frame mode syntax and default probabilities
    (reset_frame_context = 3 is written), so the compressed header and tile data stay valid; only
    the header fields differ:
      key:         marker(2) profile(2) show_existing(1) frame_type=0 show_frame error_res
                   sync(24) [profile0: color_space(3) color_range(1)] size(32) render(1..)
      intra-only:  ... frame_type=1 show_frame error_res intra_only=1 reset_ctx(2)=3 sync(24)
                   refresh_frame_flags(8) size(32) render(1..)
    The rest of the header (loop filter, quantiser, segmentation, tiles, first-partition size) is
    copied bit for bit.  Requires error_res == 0 and BT.601-style colour config (profile 0
    intra-only frames imply 8-bit 4:2:0 BT.601)."""
    r = _Bits(chunk)
    assert r.lit(2) == 2
    profile = r.bit() | (r.bit() << 1)
    assert profile == 0, "only profile 0"
    assert r.bit() == 0          # show_existing_frame
    assert r.bit() == 0, "not a key frame"
    show_frame, error_res = r.bit(), r.bit()
    assert error_res == 0
    assert r.lit(24) == 0x498342
    color_space = r.lit(3)
    r.bit()                       # color_range
    assert color_space != 7
    # intra_only can only be signalled on a hidden frame (show_frame = 0, vp9_decodeframe.c:3182)
    out = [1, 0, 0, 0, 0, 1, 0, 0, 1, 1, 1]
    out += [(0x498342 >> (23 - i)) & 1 for i in range(24)]
    out += [(refresh_mask >> (7 - i)) & 1 for i in range(8)]
    # remaining header bits up to the end of the uncompressed header are copied unchanged; the
    # header ends with first_partition_size(16) followed by byte alignment, so find its length by
    # parsing the common tail
    start = r.pos
    r.lit(16); r.lit(16)                       # width-1, height-1
    if r.bit():
        r.lit(32)                               # render size
    # key frames: refresh_frame_context, frame_parallel, frame_context_idx follow
    r.bit(); r.bit(); r.lit(2)
    r.lit(6); r.lit(3)                          # lf level, sharpness
    if r.bit():                                 # mode_ref_delta_enabled
        if r.bit():
            for _ in range(4 + 2):
                if r.bit():
                    r.lit(7)
    r.lit(8)                                    # base_q_idx
    for _ in range(3):
        if r.bit():
            r.lit(5)
    if r.bit():                                 # segmentation enabled
        if r.bit():                             # update_map
            for _ in range(7):
                if r.bit():
                    r.lit(8)
            if r.bit():
                for _ in range(3):
                    if r.bit():
                        r.lit(8)
        if r.bit():                             # update_data
            r.bit()
            for _ in range(8):
                for bits_, sign in ((8, 1), (6, 1), (2, 0), (0, 0)):
                    if r.bit():
                        r.lit(bits_)
                        if sign:
                            r.bit()
    # tile info: increment bits depend on the frame width
    return _finish_intra_only(chunk, r, out, start)


def _finish_intra_only(chunk, r, out, start):
    # width is at `start`: recompute sb_cols to parse the tile-column increment bits
    rr = _Bits(chunk)
    rr.pos = start
    width = rr.lit(16) + 1
    sb_cols = (((width + 7) >> 3) + 7) >> 3
    min_log2 = 0
    while (64 << min_log2) < sb_cols:
        min_log2 += 1
    max_log2 = 1
    while (sb_cols >> max_log2) >= 4:
        max_log2 += 1
    max_log2 -= 1
    k = min_log2
    while k < max_log2 and r.bit():
        k += 1
    if r.bit():
        r.bit()                                 # tile rows
    r.lit(16)                                   # first partition size
    end = r.pos
    rr.pos = start
    out += [rr.bit() for _ in range(end - start)]
    while len(out) % 8:
        out.append(0)
    hdr = bytes(sum(b << (7 - j) for j, b in enumerate(out[i:i + 8])) for i in range(0, len(out), 8))
    return hdr + chunk[(end + 7) // 8:]
