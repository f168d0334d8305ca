"""Hand-built packed frame buffers (include/vp9_b200.h) for randomised GPU-vs-oracle tests.

The clips under tests/golden exercise what an encoder chooses; these buffers exercise what the
packed format can express: every transform size / type / eob class, every intra mode with every
edge-availability combination, every sub-pel phase and filter bank at every frame edge, scaled
references with arbitrary ratios, compound prediction, random (but well-formed) loop-filter masks
and levels -- in the spirit of the reference's randomised unit tests (test/convolve_test.cc,
test/partial_idct_test.cc, test/vp9_intrapred_test.cc, test/lpf_test.cc).
"""
import ctypes as C

import numpy as np

import cuda_vp9_b200 as vb

BLOCK_DT = np.dtype([("mi_col", "<u2"), ("mi_row", "<u2"), ("bsize", "u1"), ("ref0", "u1"), ("ref1", "u1"),
                     ("filter", "u1"), ("mv", "<i2", (2, 2))])
SUBMV_DT = np.dtype([("luma", "<i2", (4, 2, 2)), ("chroma", "<i2", (2, 2))])
TXB_DT = np.dtype([("coef_off", "<u4"), ("x4", "<u2"), ("y4", "<u2"), ("eob", "<u2"), ("tx", "u1"), ("aux", "u1")])
SBR_DT = np.dtype([("start", "<u4"), ("count", "<u4")])
LFM_DT = np.dtype([("left_y", "<u8", 4), ("above_y", "<u8", 4), ("int_4x4_y", "<u8"), ("left_uv", "<u2", 4),
                   ("above_uv", "<u2", 4), ("int_4x4_uv", "<u2"), ("pad", "<u2", 3), ("lfl_y", "u1", 64)])
assert BLOCK_DT.itemsize == 16 and SUBMV_DT.itemsize == 40 and TXB_DT.itemsize == 12 and LFM_DT.itemsize == 160

SEC_BLOCKS, SEC_SUBMV, SEC_TXB_INTER, SEC_TXB_INTRA_Y, SEC_SB_INTRA, SEC_COEF, SEC_LFM = 0, 1, 2, 3, 6, 7, 8
# BLOCK_4X4 .. BLOCK_64X64: width / height in 8-sample units (sub-8x8 blocks cover one 8x8)
BW8 = [1, 1, 1, 1, 1, 2, 2, 2, 4, 4, 4, 8, 8]
BH8 = [1, 1, 1, 1, 2, 1, 2, 4, 2, 4, 8, 4, 8]
SQUARE = {1: 3, 2: 6, 4: 9, 8: 12}          # size in mi -> BLOCK_NxN
HORZ = {1: 2, 2: 5, 4: 8, 8: 11}            # two N x N/2 blocks (for N = 8x8: BLOCK_8X4)
VERT = {1: 1, 2: 4, 4: 7, 8: 10}


class Frame:
    """Geometry + section lists of one packed frame."""

    def __init__(self, w, h, bd=8, ss=1, dst_slot=3, intra_only=False, lossless=False):
        self.w, self.h, self.bd, self.ss = w, h, bd, ss
        self.mi_cols, self.mi_rows = (w + 7) // 8, (h + 7) // 8
        self.sb_cols, self.sb_rows = (self.mi_cols + 7) // 8, (self.mi_rows + 7) // 8
        self.dst_slot, self.intra_only, self.lossless = dst_slot, intra_only, lossless
        self.refs = []           # (slot, w, h)
        self.blocks, self.submv, self.txb_inter = [], [], [[], [], [], []]
        self.txb_intra = [[], [], []]
        self.coefs = []
        self.lfm = None
        self.lf_level, self.lf_sharp = 0, 0

    def plane_dims(self, p):
        s = self.ss if p else 0
        return (self.mi_cols * 8) >> s, (self.mi_rows * 8) >> s

    def add_coefs(self, vals):
        off = len(self.coefs)
        self.coefs.extend(int(v) for v in vals)
        return off

    def pack(self):
        h = vb.FrameHdr()
        h.magic = 0x39425056
        h.width, h.height = self.w, self.h
        h.mi_cols, h.mi_rows, h.sb_cols, h.sb_rows = self.mi_cols, self.mi_rows, self.sb_cols, self.sb_rows
        h.bit_depth, h.ss_x, h.ss_y = self.bd, self.ss, self.ss
        h.lossless, h.intra_only, h.show_frame = int(self.lossless), int(self.intra_only), 1
        h.lf_level, h.lf_sharpness = self.lf_level, self.lf_sharp
        h.coef_bytes = 4 if self.bd > 8 else 2
        h.dst_slot = self.dst_slot
        for i, (slot, rw, rh) in enumerate(self.refs):
            r = h.ref[i]
            r.slot, r.width, r.height = slot, rw, rh
            # vp9_setup_scale_factors_for_frame (vp9_scale.c:52-66)
            r.x_scale_fp = (rw << 14) // self.w
            r.y_scale_fp = (rh << 14) // self.h
            r.x_step_q4 = (16 * r.x_scale_fp) >> 14
            r.y_step_q4 = (16 * r.y_scale_fp) >> 14
        for lvl in range(64):        # update_sharpness (vp9_loopfilter.c:212-231)
            sh = self.lf_sharp
            lim = lvl >> ((sh > 0) + (sh > 4))
            if sh > 0 and lim > 9 - sh:
                lim = 9 - sh
            lim = max(lim, 1)
            h.lf_lim[lvl] = lim
            h.lf_mblim[lvl] = 2 * (lvl + 2) + lim
        nsb = self.sb_cols * self.sb_rows
        # intra transform blocks: per plane, grouped by superblock (decode order inside)
        sbr = np.zeros(3 * nsb, dtype=SBR_DT)
        intra = []
        for p in range(3):
            n = 64 >> (self.ss if p else 0)
            key = lambda t: ((t["y4"] * 4) // n) * self.sb_cols + (t["x4"] * 4) // n  # noqa: E731
            lst = sorted(enumerate(self.txb_intra[p]), key=lambda it: (key(it[1]), it[0]))
            arr = np.array([t for _, t in lst], dtype=TXB_DT) if lst else np.zeros(0, dtype=TXB_DT)
            pos = 0
            for sb in range(nsb):
                cnt = sum(1 for _, t in lst if key(t) == sb) if lst else 0
                sbr[p * nsb + sb] = (pos, cnt)
                pos += cnt
            intra.append(arr)
        inter = [np.array(x, dtype=TXB_DT) if x else np.zeros(0, dtype=TXB_DT) for x in self.txb_inter]
        for i in range(4):
            h.n_txb_inter_size[i] = len(inter[i])
        secs = {SEC_BLOCKS: np.array(self.blocks, dtype=BLOCK_DT) if self.blocks else np.zeros(0, dtype=BLOCK_DT),
                SEC_SUBMV: np.array(self.submv, dtype=SUBMV_DT) if self.submv else np.zeros(0, dtype=SUBMV_DT),
                SEC_TXB_INTER: np.concatenate(inter),
                SEC_TXB_INTRA_Y: intra[0], SEC_TXB_INTRA_Y + 1: intra[1], SEC_TXB_INTRA_Y + 2: intra[2],
                SEC_SB_INTRA: sbr,
                SEC_COEF: np.array(self.coefs, dtype="<i4" if self.bd > 8 else "<i2"),
                SEC_LFM: self.lfm if (self.lfm is not None and self.lf_level) else np.zeros(0, dtype=LFM_DT)}
        off = (C.sizeof(h) + 15) & ~15
        blobs = []
        for i in range(12):
            a = secs.get(i)
            raw = a.tobytes() if a is not None else b""
            h.sec[i].offset = off
            h.sec[i].count = len(a) if a is not None else 0
            pad = (-len(raw)) & 15
            blobs.append(raw + b"\0" * pad)
            off += len(raw) + pad
        h.total_bytes = off
        head = bytes(h)
        return head + b"\0" * (((C.sizeof(h) + 15) & ~15) - len(head)) + b"".join(blobs)


def random_planes(rng, w, h, bd, ss, smooth=False):
    maxv = (1 << bd) - 1
    dt = np.uint16 if bd > 8 else np.uint8
    out = []
    for p in range(3):
        s = ss if p else 0
        pw, ph = (w + s) >> s, (h + s) >> s
        a = rng.integers(0, maxv + 1, size=(ph, pw))
        if smooth:   # low-pass so that flat / filter masks of the loop filter trigger
            a = a.astype(np.float64)
            for _ in range(3):
                a = (a + np.roll(a, 1, 0) + np.roll(a, 1, 1) + np.roll(a, -1, 0) + np.roll(a, -1, 1)) / 5
            a = np.clip(a + rng.integers(-2, 3, size=a.shape), 0, maxv)
        out.append(a.astype(dt))
    return out


def random_coefs(rng, n, tx_size, bd, eob=None):
    """Sparse dequantised coefficients in scan order (eob entries)."""
    if eob is None:
        cls = rng.integers(0, 5)
        eob = [1, int(rng.integers(1, min(n, 11) + 1)), int(rng.integers(1, min(n, 39) + 1)),
               int(rng.integers(1, min(n, 136) + 1)), int(rng.integers(1, n + 1))][cls]
    scale = 1 << (bd - 8)
    v = rng.integers(-300 * scale, 300 * scale + 1, size=eob)
    v[rng.random(eob) < 0.5] = 0
    v[0] = rng.integers(-1500 * scale, 1500 * scale + 1)
    if rng.random() < 0.1:
        v[rng.integers(0, eob)] = rng.choice([-1, 1]) * 6000 * scale    # clipping paths
    if v[-1] == 0:
        v[-1] = 1
    return v


def add_inter_blocks(rng, f, far=0.1, intra_sbs=()):
    """Random partition of the frame into inter prediction blocks with random references.
    Superblocks listed in intra_sbs (row, col) become one intra 64x64 block instead."""
    nref = len(f.refs)

    def rand_mv():
        r = rng.random()
        span = 8 * 48
        if r < far:
            span = 8 * max(f.w, f.h)
        mv = rng.integers(-span, span + 1, size=2)
        for k in range(2):
            if rng.random() < 0.25:
                mv[k] &= ~7          # integer position: copy / 1-D paths
        return mv

    def leaf(r, c, bsize):
        if r >= f.mi_rows or c >= f.mi_cols:
            return
        b = np.zeros((), dtype=BLOCK_DT)
        b["mi_col"], b["mi_row"], b["bsize"] = c, r, bsize
        b["ref0"] = rng.integers(1, nref + 1)
        b["ref1"] = rng.integers(1, nref + 1) if rng.random() < 0.35 else 0
        b["filter"] = rng.integers(0, 4)
        if bsize < 3:
            s = np.zeros((), dtype=SUBMV_DT)
            for k in range(4):
                for rr in range(2):
                    s["luma"][k][rr] = rand_mv()
            for rr in range(2):
                s["chroma"][rr] = rand_mv()
            idx = len(f.submv)
            f.submv.append(s)
            b["mv"][0] = np.array([idx & 0xFFFF, idx >> 16], dtype=np.uint16).view(np.int16)
        else:
            b["mv"][0], b["mv"][1] = rand_mv(), rand_mv()
        f.blocks.append(b)

    def part(r, c, n):   # n: size in mi
        if r >= f.mi_rows or c >= f.mi_cols:
            return
        k = rng.random()
        if n == 1:
            leaf(r, c, int(rng.integers(0, 4)))
        elif k < 0.25:
            leaf(r, c, SQUARE[n])
        elif k < 0.4:
            leaf(r, c, HORZ[n]); leaf(r + n // 2, c, HORZ[n])
        elif k < 0.55:
            leaf(r, c, VERT[n]); leaf(r, c + n // 2, VERT[n])
        else:
            h2 = n // 2
            for dr, dc in ((0, 0), (0, h2), (h2, 0), (h2, h2)):
                part(r + dr, c + dc, h2)

    for r in range(0, f.mi_rows, 8):
        for c in range(0, f.mi_cols, 8):
            if (r // 8, c // 8) in intra_sbs:
                b = np.zeros((), dtype=BLOCK_DT)
                b["mi_col"], b["mi_row"], b["bsize"] = c, r, 12
                f.blocks.append(b)
            else:
                part(r, c, 8)


def add_inter_residuals(rng, f, density=0.3, intra_sbs=()):
    for p in range(3):
        sbn = 64 >> (f.ss if p else 0)
        pw, ph = f.plane_dims(p)
        for tx_size in range(4):
            n = 4 << tx_size
            if f.lossless and tx_size:
                continue
            for y in range(0, ph, n):
                for x in range(0, pw, n):
                    if rng.random() > density / (1 + tx_size) or (y // sbn, x // sbn) in intra_sbs:
                        continue
                    v = random_coefs(rng, n * n, tx_size, f.bd)
                    t = np.zeros((), dtype=TXB_DT)
                    t["coef_off"], t["x4"], t["y4"], t["eob"] = f.add_coefs(v), x // 4, y // 4, len(v)
                    t["tx"] = tx_size | (p << 4)
                    f.txb_inter[tx_size].append(t)
    # overlapping transform blocks would make the result order dependent: keep one size per area
    # by construction (each size draws its own positions, so remove overlaps across sizes)
    taken = [np.zeros((f.plane_dims(p)[1] // 4 + 16, f.plane_dims(p)[0] // 4 + 16), dtype=bool) for p in range(3)]
    for tx_size in (3, 2, 1, 0):
        keep = []
        k = 1 << tx_size
        for t in f.txb_inter[tx_size]:
            p = int(t["tx"]) >> 4
            ys, xs = int(t["y4"]), int(t["x4"])
            if taken[p][ys:ys + k, xs:xs + k].any():
                continue
            taken[p][ys:ys + k, xs:xs + k] = True
            keep.append(t)
        f.txb_inter[tx_size] = keep


def add_intra_blocks(rng, f, tx_size_of_sb=None, coef_prob=0.7, tile_cols=(), only_sbs=None):
    """Every plane is covered with intra transform blocks, one transform size per superblock,
    raster order inside the superblock (a valid decode order for equal-size blocks)."""
    for p in range(3):
        pw, ph = f.plane_dims(p)
        sbn = 64 >> (f.ss if p else 0)
        for sy in range(0, ph, sbn):
            for sx in range(0, pw, sbn):
                if only_sbs is not None and (sy // sbn, sx // sbn) not in only_sbs:
                    continue
                tx_size = 0 if f.lossless else int(rng.integers(0, 4))
                if tx_size_of_sb is not None:
                    tx_size = tx_size_of_sb
                n = 4 << tx_size
                if n > sbn:
                    n, tx_size = sbn, {32: 3, 16: 2, 8: 1, 4: 0}[sbn]
                for y in range(sy, min(sy + sbn, ph), n):
                    for x in range(sx, min(sx + sbn, pw), n):
                        mode = int(rng.integers(0, 10))
                        have_top = int(y > 0)
                        have_left = int(x > 0 and (x not in tile_cols or rng.random() < 0.5))
                        if x > 0 and rng.random() < 0.05:
                            have_left = 0      # tile boundary inside the frame
                        have_right = int(y > 0 and (x - sx) + 2 * n <= sbn and rng.random() < 0.7)
                        t = np.zeros((), dtype=TXB_DT)
                        t["x4"], t["y4"] = x // 4, y // 4
                        tx_type = 0 if (tx_size == 3 or f.lossless or p) else int(rng.integers(0, 4))
                        t["tx"] = tx_size | (tx_type << 2) | (p << 4)
                        t["aux"] = mode | (have_top << 4) | (have_left << 5) | (have_right << 6)
                        if rng.random() < coef_prob:
                            v = random_coefs(rng, n * n, tx_size, f.bd)
                            t["coef_off"], t["eob"] = f.add_coefs(v), len(v)
                        f.txb_intra[p].append(t)


def add_loop_filter(rng, f, level=None):
    """Random well-formed masks: sizes disjoint per edge position, nothing on the frame's left
    column / top row, nothing outside the frame (vp9_adjust_mask, vp9_loopfilter.c:766-878)."""
    f.lf_level = int(level if level is not None else rng.integers(1, 64))
    f.lf_sharp = int(rng.integers(0, 8))
    nsb = f.sb_cols * f.sb_rows
    lfm = np.zeros(nsb, dtype=LFM_DT)
    for sr in range(f.sb_rows):
        for sc in range(f.sb_cols):
            m = lfm[sr * f.sb_cols + sc]
            rows = min(8, f.mi_rows - sr * 8)
            cols = min(8, f.mi_cols - sc * 8)
            valid_y = 0
            for r in range(rows):
                valid_y |= ((1 << cols) - 1) << (8 * r)
            rows_uv, cols_uv = (rows + 1) >> 1, (cols + 1) >> 1
            valid_uv = 0
            for r in range(rows_uv):
                valid_uv |= ((1 << cols_uv) - 1) << (4 * r)
            for name, valid, bits, edge_clear in (("left_y", valid_y, 64, 0x0101010101010101 if sc == 0 else 0),
                                                  ("above_y", valid_y, 64, 0xFF if sr == 0 else 0),
                                                  ("left_uv", valid_uv, 16, 0x1111 if sc == 0 else 0),
                                                  ("above_uv", valid_uv, 16, 0xF if sr == 0 else 0)):
                sel = [int(rng.integers(0, 4)) if rng.random() < 0.8 else 3 for _ in range(bits)]   # 3 = no edge
                for size in range(3):
                    v = 0
                    for b in range(bits):
                        if sel[b] == size:
                            v |= 1 << b
                    m[name][size] = v & valid & ~edge_clear
            m["int_4x4_y"] = int(rng.integers(0, 1 << 63)) & valid_y
            # Chroma of an odd number of 8x8 columns / rows ends in a half block: libvpx applies no
            # inner 4x4 edge there and replaces a wide filter by the 8-wide one, so that no filter
            # reads beyond the decoded area (vp9_adjust_mask, vp9_loopfilter.c:805-823, 846-870)
            int_uv = 0
            for r in range(rows >> 1):
                int_uv |= ((1 << (cols >> 1)) - 1) << (4 * r)
            m["int_4x4_uv"] = int(rng.integers(0, 1 << 16)) & int_uv
            if cols & 1:
                last_col = 0x1111 << (cols_uv - 1)
                m["left_uv"][1] |= m["left_uv"][2] & last_col
                m["left_uv"][2] &= ~last_col & 0xFFFF
            if rows & 1:
                last_row = 0xF << (4 * (rows_uv - 1))
                m["above_uv"][1] |= m["above_uv"][2] & last_row
                m["above_uv"][2] &= ~last_row & 0xFFFF
            m["lfl_y"] = rng.integers(1, 64, size=64)
    f.lfm = lfm
