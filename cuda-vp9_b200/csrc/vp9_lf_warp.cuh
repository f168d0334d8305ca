// Loop filter, 8-bit 4:2:0: one WARP per superblock column, four samples per lane.
//
// Same data flow as vp9_lf_sb.cuh (one persistent worker per superblock column, phases
// A / H7 / E / D / H ordered by the F_edge flags), but the worker is a single warp and a lane owns
// four sample rows (vertical edges) or four sample columns (horizontal edges) of a plane:
//   lanes 0-15  luma   rows / columns 4l .. 4l+3
//   lanes 16-23 U      rows / columns 4(l-16) ..
//   lanes 24-31 V
// Four neighbouring rows lie in one 8-row band and four neighbouring columns in one 8-column block,
// so the mask bits, the filter size and the level are uniform per lane; the samples of the four
// positions travel as the four bytes of a 32-bit register (one word load / store per tap instead
// of four byte accesses) and the filter arithmetic is done on packed data:
//   * thresholds (filter_mask, flat, flat2, hev; vpx_dsp/loopfilter.c:31-82) on u8x4 with
//     VABSDIFF4 and carry-free byte compares;
//   * filter4 (:84-116) on 16x2 lanes in biased-unsigned form (VIADD / VIMNMX.U16x2 only, no
//     cross-lane borrow); filter8 / filter16 (:201-216, :340-357 + :218-289) as running window sums
//     on 16x2 lanes;
//   * the variants are blended per byte with the masks (bit select), and a variant nobody in the
//     lane needs is skipped.
// Vertical edges need the taps of four rows transposed (4x4 byte transposes, 8 PRMT each way).
// Tiles are fetched one superblock ahead with cp.async straight into the second tile buffer.
//
// Like vp9_lf_sb.cuh the per-lane functions also compile for the host (VP9B_EMUL): the packed
// arithmetic is compared with the scalar reference filter sample by sample, and the whole
// decomposition runs on the CPU against the oracle (tests/emul/lfw_emul.cc).
#pragma once
#include "vp9_lf_sb.cuh"

namespace vp9b {

// ------------------------------------------------------------------------------------------
// packed primitives
#ifdef VP9B_EMUL
LF_HD uint32_t lfw_absdiff4(uint32_t a, uint32_t b) {
  uint32_t r = 0;
  for (int i = 0; i < 4; i++) {
    const int x = (a >> (8 * i)) & 0xff, y = (b >> (8 * i)) & 0xff;
    r |= (uint32_t)(x > y ? x - y : y - x) << (8 * i);
  }
  return r;
}
LF_HD uint32_t lfw_prmt(uint32_t a, uint32_t b, uint32_t sel) {   // prmt.b32, default mode (bit 3 of a selector: replicate the sign)
  const uint64_t v = ((uint64_t)b << 32) | a;
  uint32_t r = 0;
  for (int i = 0; i < 4; i++) {
    const uint32_t s = (sel >> (4 * i)) & 0xf;
    uint32_t byte = (uint32_t)(v >> (8 * (s & 7))) & 0xff;
    if (s & 8) byte = (byte & 0x80) ? 0xff : 0x00;
    r |= byte << (8 * i);
  }
  return r;
}
LF_HD uint32_t lfw_maxu2(uint32_t a, uint32_t b) {
  const uint32_t lo = max(a & 0xffffu, b & 0xffffu), hi = max(a >> 16, b >> 16);
  return lo | (hi << 16);
}
LF_HD uint32_t lfw_minu2(uint32_t a, uint32_t b) {
  const uint32_t lo = min(a & 0xffffu, b & 0xffffu), hi = min(a >> 16, b >> 16);
  return lo | (hi << 16);
}
#else
LF_HD uint32_t lfw_absdiff4(uint32_t a, uint32_t b) { return __vabsdiffu4(a, b); }
LF_HD uint32_t lfw_prmt(uint32_t a, uint32_t b, uint32_t sel) {
  uint32_t r;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel));
  return r;
}
LF_HD uint32_t lfw_maxu2(uint32_t a, uint32_t b) { return __vmaxu2(a, b); }
LF_HD uint32_t lfw_minu2(uint32_t a, uint32_t b) { return __vminu2(a, b); }
#endif

LF_HD uint32_t lfw_sel(uint32_t m, uint32_t a, uint32_t b) { return (a & m) | (b & ~m); }   // one LOP3
LF_HD uint32_t lfw_lo(uint32_t x) { return lfw_prmt(x, 0, 0x4140); }     // bytes 0,1 -> 16x2 lanes
LF_HD uint32_t lfw_hi(uint32_t x) { return lfw_prmt(x, 0, 0x4342); }     // bytes 2,3 -> 16x2 lanes
LF_HD uint32_t lfw_pack(uint32_t lo, uint32_t hi) { return lfw_prmt(lo, hi, 0x6420); }   // low bytes of the four lanes
LF_HD uint32_t lfw_expand(uint32_t flags) { return lfw_prmt(flags, 0, 0xba98); }          // bit 7 of a byte -> 0xff / 0x00
LF_HD uint32_t lfw_clamp2(uint32_t v, uint32_t lo, uint32_t hi) { return lfw_minu2(lfw_maxu2(v, lo), hi); }
// bit 7 of a byte of the result is set iff the byte of x is > L, K = (127 - L) in every byte (L <= 127)
LF_HD uint32_t lfw_gt(uint32_t x, uint32_t K) { return ((x & 0x7f7f7f7fu) + K) | x; }

// 4x4 byte transpose: in a[r] byte c  ->  out a[c] byte r (its own inverse)
LF_HD void lfw_transpose4(uint32_t &a0, uint32_t &a1, uint32_t &a2, uint32_t &a3) {
  const uint32_t t0 = lfw_prmt(a0, a1, 0x5140), t1 = lfw_prmt(a2, a3, 0x5140);
  const uint32_t t2 = lfw_prmt(a0, a1, 0x7362), t3 = lfw_prmt(a2, a3, 0x7362);
  a0 = lfw_prmt(t0, t1, 0x5410);
  a1 = lfw_prmt(t0, t1, 0x7632);
  a2 = lfw_prmt(t2, t3, 0x5410);
  a3 = lfw_prmt(t2, t3, 0x7632);
}

// filter4 on two samples (16x2 lanes holding 0..255), hev = 0xffff / 0 per lane
LF_HD void lfw_filter4_half(uint32_t &p1, uint32_t &p0, uint32_t &q0, uint32_t &q1, uint32_t hev) {
  const uint32_t C256 = 0x01000100u;
  uint32_t d = lfw_clamp2(p1 + C256 - q1, 0x00800080u, 0x017f017fu);          // clamp(ps1 - qs1) + 256
  const uint32_t f = lfw_sel(hev, d, C256);                                  // & hev
  const uint32_t e = q0 + C256 - p0;                                          // (qs0 - ps0) + 256
  uint32_t t = lfw_clamp2(f + e + e + e, 896u * 0x00010001u, 1151u * 0x00010001u);   // filter + 1024
  const uint32_t f1 = (lfw_minu2(t + 0x00040004u, 1151u * 0x00010001u) >> 3) & 0x1fff1fffu;   // filter1 + 128
  const uint32_t f2 = (lfw_minu2(t + 0x00030003u, 1151u * 0x00010001u) >> 3) & 0x1fff1fffu;   // filter2 + 128
  const uint32_t nq0 = lfw_clamp2(q0 + 0x008f008fu - f1, 0x000f000fu, 270u * 0x00010001u) - 0x000f000fu;
  const uint32_t np0 = lfw_clamp2(p0 + f2, 0x00800080u, 383u * 0x00010001u) - 0x00800080u;
  uint32_t h = ((f1 + 0x00010001u) >> 1) & 0x7fff7fffu;                       // ((filter1 + 1) >> 1) + 64
  h = lfw_sel(hev, 0x00400040u, h);                                           // & ~hev
  const uint32_t nq1 = lfw_clamp2(q1 + 0x00480048u - h, 0x00080008u, 263u * 0x00010001u) - 0x00080008u;
  const uint32_t np1 = lfw_clamp2(p1 + h, 0x00400040u, 319u * 0x00010001u) - 0x00400040u;
  p1 = np1; p0 = np0; q0 = nq0; q1 = nq1;
}

// One edge at four sample positions.  v[0..15] = p7 .. p0, q0 .. q7 (u8x4 each; sizes 4 and 8 only
// look at v[4..11]).  Modified taps are written back to v; returns false when nothing changed.
template <int SIZE>
LF_HD bool lfw_edge(uint32_t *v, int blimit, int limit, int thresh) {
  const uint32_t P3 = v[4], P2 = v[5], P1 = v[6], P0 = v[7], Q0 = v[8], Q1 = v[9], Q2 = v[10], Q3 = v[11];
  const uint32_t d10p = lfw_absdiff4(P1, P0), d10q = lfw_absdiff4(Q1, Q0);
  // filter_mask: every neighbour difference <= limit and 2|p0-q0| + |p1-q1|/2 <= blimit
  const uint32_t Kl = (uint32_t)(127 - limit) * 0x01010101u;
  uint32_t ex = lfw_gt(lfw_absdiff4(P3, P2), Kl) | lfw_gt(lfw_absdiff4(P2, P1), Kl) | lfw_gt(d10p, Kl) | lfw_gt(d10q, Kl) |
                lfw_gt(lfw_absdiff4(Q2, Q1), Kl) | lfw_gt(lfw_absdiff4(Q3, Q2), Kl);
  {
    const uint32_t a = lfw_absdiff4(P0, Q0), b = (lfw_absdiff4(P1, Q1) >> 1) & 0x7f7f7f7fu;
    const uint32_t Kb = (uint32_t)(0x7fff - blimit) * 0x00010001u;
    const uint32_t alo = lfw_lo(a), ahi = lfw_hi(a);
    const uint32_t tlo = alo + alo + lfw_lo(b) + Kb, thi = ahi + ahi + lfw_hi(b) + Kb;   // bit 15 of a lane: > blimit
    ex |= lfw_prmt(tlo, thi, 0x7531);
  }
  const uint32_t mask = ~lfw_expand(ex & 0x80808080u);
  if (!mask) return false;
  uint32_t flat = 0, flat2 = 0;
  if (SIZE >= 8) {
    uint32_t fl = (d10p | d10q | lfw_absdiff4(P2, P0) | lfw_absdiff4(Q2, Q0) | lfw_absdiff4(P3, P0) | lfw_absdiff4(Q3, Q0)) & 0xfefefefeu;
    flat = ~lfw_expand(lfw_gt(fl, 0x7f7f7f7fu) & 0x80808080u) & mask;        // every difference <= 1
    if (SIZE == 16 && flat) {
      fl = (lfw_absdiff4(v[3], P0) | lfw_absdiff4(v[12], Q0) | lfw_absdiff4(v[2], P0) | lfw_absdiff4(v[13], Q0) |
            lfw_absdiff4(v[1], P0) | lfw_absdiff4(v[14], Q0) | lfw_absdiff4(v[0], P0) | lfw_absdiff4(v[15], Q0)) & 0xfefefefeu;
      flat2 = ~lfw_expand(lfw_gt(fl, 0x7f7f7f7fu) & 0x80808080u) & flat;
    }
  }
  uint32_t o[16];
#pragma unroll
  for (int i = 0; i < 16; i++) o[i] = v[i];
  if (mask & ~flat) {   // filter4 somewhere
    const uint32_t Kt = (uint32_t)(127 - thresh) * 0x01010101u;
    const uint32_t hev = lfw_expand((lfw_gt(d10p, Kt) | lfw_gt(d10q, Kt)) & 0x80808080u);
    uint32_t p1l = lfw_lo(P1), p0l = lfw_lo(P0), q0l = lfw_lo(Q0), q1l = lfw_lo(Q1);
    uint32_t p1h = lfw_hi(P1), p0h = lfw_hi(P0), q0h = lfw_hi(Q0), q1h = lfw_hi(Q1);
    lfw_filter4_half(p1l, p0l, q0l, q1l, lfw_prmt(hev, 0, 0x1100));
    lfw_filter4_half(p1h, p0h, q0h, q1h, lfw_prmt(hev, 0, 0x3322));
    const uint32_t m4 = mask & ~flat;
    o[6] = lfw_sel(m4, lfw_pack(p1l, p1h), o[6]);
    o[7] = lfw_sel(m4, lfw_pack(p0l, p0h), o[7]);
    o[8] = lfw_sel(m4, lfw_pack(q0l, q0h), o[8]);
    o[9] = lfw_sel(m4, lfw_pack(q1l, q1h), o[9]);
  }
  if (SIZE >= 8 && (flat & ~flat2)) {   // filter8 somewhere
    uint32_t r[6][2];
#pragma unroll
    for (int hh = 0; hh < 2; hh++) {
      uint32_t t[8];
#pragma unroll
      for (int i = 0; i < 8; i++) t[i] = hh ? lfw_hi(v[4 + i]) : lfw_lo(v[4 + i]);   // p3 p2 p1 p0 q0 q1 q2 q3
      uint32_t s = t[0] + t[0] + t[0] + t[1] + t[1] + t[2] + t[3] + t[4] + 0x00040004u;   // window of op2 (+ rounding)
      r[0][hh] = (s >> 3) & 0x00ff00ffu;
      s = s - t[0] - t[1] + t[2] + t[5]; r[1][hh] = (s >> 3) & 0x00ff00ffu;   // op1
      s = s - t[0] - t[2] + t[3] + t[6]; r[2][hh] = (s >> 3) & 0x00ff00ffu;   // op0
      s = s - t[0] - t[3] + t[4] + t[7]; r[3][hh] = (s >> 3) & 0x00ff00ffu;   // oq0
      s = s - t[1] - t[4] + t[5] + t[7]; r[4][hh] = (s >> 3) & 0x00ff00ffu;   // oq1
      s = s - t[2] - t[5] + t[6] + t[7]; r[5][hh] = (s >> 3) & 0x00ff00ffu;   // oq2
    }
    const uint32_t m8 = flat & ~flat2;
#pragma unroll
    for (int i = 0; i < 6; i++) o[5 + i] = lfw_sel(m8, lfw_pack(r[i][0], r[i][1]), o[5 + i]);
  }
  if (SIZE == 16 && flat2) {   // filter16 somewhere
    uint32_t r[14][2];
#pragma unroll
    for (int hh = 0; hh < 2; hh++) {
      uint32_t t[16];
#pragma unroll
      for (int i = 0; i < 16; i++) t[i] = hh ? lfw_hi(v[i]) : lfw_lo(v[i]);            // p7 .. p0 q0 .. q7
      // window of op6 without its doubled centre: 7 p7 + p6 + ... + p0 + q0 (+ rounding)
      uint32_t s = t[0] * 7u + t[1] + t[2] + t[3] + t[4] + t[5] + t[6] + t[7] + t[8] + 0x00080008u;
#pragma unroll
      for (int k = 1; k <= 14; k++) {      // output k = tap index of the modified sample (p6 = 1 .. q6 = 14)
        r[k - 1][hh] = ((s + t[k]) >> 4) & 0x00ff00ffu;
        // slide: the sample leaving on the p side is p7 until the window has moved past it, the one
        // entering on the q side is q7 once the window reaches beyond it
        const int out_i = k - 7 > 0 ? k - 7 : 0, in_i = k + 8 < 15 ? k + 8 : 15;
        s = s - t[out_i] + t[in_i];
      }
    }
#pragma unroll
    for (int k = 1; k <= 14; k++) o[k] = lfw_sel(flat2, lfw_pack(r[k - 1][0], r[k - 1][1]), o[k]);
  }
  constexpr int a = SIZE == 16 ? 1 : (SIZE == 8 ? 5 : 6), b = SIZE == 16 ? 14 : (SIZE == 8 ? 10 : 9);
#pragma unroll
  for (int i = a; i <= b; i++) v[i] = o[i];
  return true;
}

// ------------------------------------------------------------------------------------------
// geometry: same tile layout as LfGeom<uint8_t, true> (rows -8..n-1, columns 0..n+7, odd word
// pitch), 32 lanes
struct LfwGeom {
  static constexpr int NT = 32;
  static constexpr int pitchY = 76, pitchC = 44;
  static constexpr int offU = 72 * pitchY, offV = offU + 40 * pitchC;
  static constexpr int kTileBytes = (offV + 40 * pitchC + 8 + 15) & ~15;
};

struct LfwCtx {
  uint8_t *tile;
  uint8_t *g[3];
  int stride[3];
  const vp9b_lfm *M, *Mn;
  const uint8_t *lf_lim, *lf_mblim;
  uint16_t *dset;
  int sbr, sbc, sb_cols, mi_rows;
  bool has_next, has_top;
};

// lane -> (plane, group of four rows / columns)
LF_HD bool lfw_role(int lane, int &plane, int &grp) {
  if (lane < 16) { plane = 0; grp = lane; }
  else if (lane < 24) { plane = 1; grp = lane - 16; }
  else { plane = 2; grp = lane - 24; }
  return true;
}
LF_HD int lfw_plane_off(int plane) { return plane == 0 ? 0 : (plane == 1 ? LfwGeom::offU : LfwGeom::offV); }
LF_HD uint8_t *lfw_at(const LfwCtx &C, int plane, int y, int x) {
  return C.tile + lfw_plane_off(plane) + (y + 8) * (plane ? LfwGeom::pitchC : LfwGeom::pitchY) + x;
}
LF_HD uint32_t lfw_ldw(const uint8_t *p) { return *(const uint32_t *)p; }
LF_HD void lfw_stw(uint8_t *p, uint32_t v) { *(uint32_t *)p = v; }

// One edge of width SIZE at four sample positions of the tile, out of line: every phase of the
// kernel calls this one copy (inlined at each call site the kernel was 160 KB of code and lived
// in instruction-cache misses).  vertical: s = (first of the lane's four rows, column of q0), taps
// along x -- the rows are transposed on the way in and out; horizontal: s = (row of q0, first of
// the lane's four columns), taps along y.
#ifdef VP9B_EMUL
#define LFW_NOINLINE inline
#else
#define LFW_NOINLINE __device__ __noinline__
#endif
template <int SIZE>
LF_HD void lfw_apply_size(uint8_t *s, int pitch, bool vertical, int blimit, int limit, int thresh) {
  uint32_t v[16];
  if (vertical) {
    constexpr int W = SIZE == 16 ? 4 : 2;         // words per row: taps x - 2W .. x + 2W - 1
    uint8_t *t = s - 2 * W;
#pragma unroll
    for (int w = 0; w < W; w++) {
      uint32_t a0 = lfw_ldw(t + 4 * w), a1 = lfw_ldw(t + pitch + 4 * w), a2 = lfw_ldw(t + 2 * pitch + 4 * w), a3 = lfw_ldw(t + 3 * pitch + 4 * w);
      lfw_transpose4(a0, a1, a2, a3);
      const int base = 8 - 2 * W + 4 * w;
      v[base] = a0; v[base + 1] = a1; v[base + 2] = a2; v[base + 3] = a3;
    }
    if (!lfw_edge<SIZE>(v, blimit, limit, thresh)) return;
#pragma unroll
    for (int w = 0; w < W; w++) {
      const int base = 8 - 2 * W + 4 * w;
      uint32_t a0 = v[base], a1 = v[base + 1], a2 = v[base + 2], a3 = v[base + 3];
      lfw_transpose4(a0, a1, a2, a3);
      lfw_stw(t + 4 * w, a0); lfw_stw(t + pitch + 4 * w, a1); lfw_stw(t + 2 * pitch + 4 * w, a2); lfw_stw(t + 3 * pitch + 4 * w, a3);
    }
  } else {
    constexpr int N = SIZE == 16 ? 8 : 4;         // taps per side
#pragma unroll
    for (int i = -N; i < N; i++) v[8 + i] = lfw_ldw(s + i * pitch);
    if (!lfw_edge<SIZE>(v, blimit, limit, thresh)) return;
    constexpr int a = SIZE == 16 ? 1 : (SIZE == 8 ? 5 : 6), b = SIZE == 16 ? 14 : (SIZE == 8 ? 10 : 9);
#pragma unroll
    for (int i = a; i <= b; i++) lfw_stw(s + (i - 8) * pitch, v[i]);
  }
}
LFW_NOINLINE void lfw_apply(uint8_t *s, int pitch, int size, bool vertical, int blimit, int limit, int thresh) {
  if (size == 16) lfw_apply_size<16>(s, pitch, vertical, blimit, limit, thresh);
  else if (size == 8) lfw_apply_size<8>(s, pitch, vertical, blimit, limit, thresh);
  else lfw_apply_size<4>(s, pitch, vertical, blimit, limit, thresh);
}
template <int SIZE>
LF_HD void lfw_v_edge(uint8_t *row0, int pitch, int x, int lvl, const uint8_t *lf_lim, const uint8_t *lf_mblim) {
  lfw_apply(row0 + x, pitch, SIZE, true, lf_mblim[lvl], lf_lim[lvl], lvl >> 4);
}
template <int SIZE>
LF_HD void lfw_h_edge(uint8_t *s, int pitch, int lvl, const uint8_t *lf_lim, const uint8_t *lf_mblim) {
  lfw_apply(s, pitch, SIZE, false, lf_mblim[lvl], lf_lim[lvl], lvl >> 4);
}

// Vertical edges of the lane's four rows (same selection rules as lf_v_row): sel_edge / sel_inner
// select the block columns whose 8-aligned / inner 4x4 edge is applied now.  row0 = tile address of
// (first row, x = 0 of the superblock described by M).
LF_HD void lfw_v_rows(const vp9b_lfm &M, bool uvsub, int y, uint8_t *row0, int pitch, uint32_t sel_edge, uint32_t sel_inner,
                      const uint8_t *lf_lim, const uint8_t *lf_mblim) {
  const int r8 = y >> 3, pr = r8 & ~1;
  uint32_t m16, m8, m4, m4i, p16, p8, p4, p4i;
  lf_band_masks_v(M, uvsub, r8, m16, m8, m4, m4i);
  lf_band_masks_v(M, uvsub, r8 ^ 1, p16, p8, p4, p4i);
  const uint32_t m16n = m16 & p16;   // set in both bands of the pair: the first band's level for both
  m16 &= sel_edge; m8 &= sel_edge; m4 &= sel_edge; m4i &= sel_inner;
  uint32_t all = m16 | m8 | m4 | m4i;
  if (!all) return;
  const int lrow = uvsub ? 2 * r8 : r8, prow = uvsub ? 2 * pr : pr, lsh = uvsub ? 16 : 8;
  uint64_t lv_own, lv_pair;
  memcpy(&lv_own, &M.lfl_y[lrow * 8], 8);
  memcpy(&lv_pair, &M.lfl_y[prow * 8], 8);
  while (all) {
    const int c = lf_ffs(all) - 1;
    all &= all - 1;
    const int lo = (int)(lv_own >> (c * lsh)) & 0xff;
    if ((m16 >> c) & 1) lfw_v_edge<16>(row0, pitch, c * 8, ((m16n >> c) & 1) ? ((int)(lv_pair >> (c * lsh)) & 0xff) : lo, lf_lim, lf_mblim);
    if ((m8 >> c) & 1) lfw_v_edge<8>(row0, pitch, c * 8, lo, lf_lim, lf_mblim);
    if ((m4 >> c) & 1) lfw_v_edge<4>(row0, pitch, c * 8, lo, lf_lim, lf_mblim);
    if ((m4i >> c) & 1) lfw_v_edge<4>(row0, pitch, c * 8 + 4, lo, lf_lim, lf_mblim);
  }
}

// Horizontal edges of the lane's four columns x0..x0+3 (lf_h_col / lf_h_unit with four positions)
LF_HD void lfw_h_cols(const vp9b_lfm &M, bool uvsub, int x0, uint8_t *col0, int pitch, int mi_row, int mi_rows, int nb,
                      const uint8_t *lf_lim, const uint8_t *lf_mblim) {
  const LfHMasks H = lf_h_masks(M, uvsub, mi_row, mi_rows, nb);
  const int c = x0 >> 3;
  const uint64_t colsel = uvsub ? 0x1111ull : 0x0101010101010101ull;
  uint64_t todo = ((H.a16 | H.a8 | H.a4 | H.ai) >> c) & colsel;
  if (H.nrow < nb) todo &= ((uint64_t)1 << (H.nrow << H.rsh)) - 1;
  while (todo) {
    const int bit0 = (todo & 0xffffffffu) ? lf_ffs((uint32_t)todo) - 1 : 32 + lf_ffs((uint32_t)(todo >> 32)) - 1;
    todo &= todo - 1;
    const int r = bit0 >> H.rsh, bit = (r << H.rsh) + c;
    const uint32_t k16 = (uint32_t)(H.a16 >> bit) & 1, k8 = (uint32_t)(H.a8 >> bit) & 1, k4 = (uint32_t)(H.a4 >> bit) & 1,
                   ki = (uint32_t)(H.ai >> bit) & 1;
    uint8_t *s = col0 + r * 8 * pitch;
    const uint8_t *lfl = &M.lfl_y[(r * H.lstep) * 8];
    const int lvl = lfl[c * H.lstep];
    if (k16) {
      const uint32_t row16 = (uint32_t)(H.a16 >> (r << H.rsh)) & 0xff;
      const uint32_t below = ~row16 & ((1u << c) - 1);
      const int run = below ? c - 1 - (31 - lf_clz(below)) : c;
      lfw_h_edge<16>(s, pitch, (run & 1) ? lfl[(c - 1) * H.lstep] : lvl, lf_lim, lf_mblim);
    } else if (k8) {
      lfw_h_edge<8>(s, pitch, lvl, lf_lim, lf_mblim);
    } else if (k4) {
      lfw_h_edge<4>(s, pitch, lvl, lf_lim, lf_mblim);
    }
    if (!k16 && ki) lfw_h_edge<4>(s + 4 * pitch, pitch, lvl, lf_lim, lf_mblim);
  }
}

// ------------------------------------------------------------------------------------------
// phase functions, called by every lane (0..31) between warp barriers

// dependent sets of one superblock: 16 (plane, band) units -> dset[plane * 8 + band]; chain flag
LF_HD bool lfw_plan_unit(const vp9b_lfm &M, bool has_left, int u, uint16_t *dset) { return lf_plan_unit<uint8_t, true>(M, has_left, u, dset); }
constexpr int kLfwPlanUnits = LfPlanUnits<uint8_t, true>::value;    // 16

// Fetch of a superblock's interior rows (+ the 8 look-ahead columns of the next superblock) into a
// tile buffer.  Device: cp.async, 4 bytes per copy (tile rows are only word aligned), completion is
// awaited by lfw_fetch_wait(); host emulation: plain copies.  The lane -> (row, word) mapping is
// fixed per loop so that an iteration costs one copy and two pointer increments.
LF_HD void lfw_copy4(uint8_t *d, const uint8_t *s) {
#ifdef VP9B_EMUL
  memcpy(d, s, 4);
#else
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(d)), "l"(s) : "memory");
#endif
}
// rows r0, r0 + dr, ... (n_it of them) of one plane: word `w` of each row
LF_HD void lfw_fetch_rows(uint8_t *tile_plane, int pitch, const uint8_t *g, int stride, int r0, int dr, int n_it, int w) {
  uint8_t *d = tile_plane + (r0 + 8) * pitch + 4 * w;
  const uint8_t *s = g + (size_t)r0 * stride + 4 * w;
#pragma unroll 4
  for (int k = 0; k < n_it; k++, d += dr * pitch, s += (size_t)dr * stride) lfw_copy4(d, s);
}
LF_HD void lfw_phase_fetch(uint8_t *tile, uint8_t *const g[3], const int stride[3], bool has_next, int lane) {
  // luma: 64 rows x 16 words (two rows per iteration) + 2 look-ahead words (16 rows per iteration)
  lfw_fetch_rows(tile, LfwGeom::pitchY, g[0], stride[0], lane >> 4, 2, 32, lane & 15);
  if (has_next) lfw_fetch_rows(tile, LfwGeom::pitchY, g[0], stride[0], lane >> 1, 16, 4, 16 + (lane & 1));
  // chroma: 32 rows x 8 words (four rows per iteration) + 2 look-ahead words per plane
#pragma unroll
  for (int p = 1; p < 3; p++) {
    uint8_t *tp = tile + lfw_plane_off(p);
    lfw_fetch_rows(tp, LfwGeom::pitchC, g[p], stride[p], lane >> 3, 4, 8, lane & 7);
    if (has_next) lfw_fetch_rows(tp, LfwGeom::pitchC, g[p], stride[p], lane >> 1, 16, 2, 8 + (lane & 1));
  }
#ifndef VP9B_EMUL
  asm volatile("cp.async.commit_group;" ::: "memory");
#endif
}
LF_HD void lfw_fetch_wait() {
#ifndef VP9B_EMUL
  asm volatile("cp.async.wait_group 0;" ::: "memory");
#endif
}

// bottom 8 rows of the finished superblock -> top rows (-8..-1) of the next tile buffer
LF_HD void lfw_phase_carry(const LfwCtx &C, uint8_t *next_tile, int lane) {
  {   // luma: 8 rows x 16 words, two rows per iteration
    const uint8_t *s = C.tile + (56 + 8 + (lane >> 4)) * LfwGeom::pitchY + 4 * (lane & 15);
    uint8_t *d = next_tile + (lane >> 4) * LfwGeom::pitchY + 4 * (lane & 15);
#pragma unroll
    for (int k = 0; k < 4; k++) lfw_stw(d + 2 * k * LfwGeom::pitchY, lfw_ldw(s + 2 * k * LfwGeom::pitchY));
  }
#pragma unroll
  for (int p = 1; p < 3; p++) {   // chroma: 8 rows x 8 words, four rows per iteration
    const int off = lfw_plane_off(p);
    const uint8_t *s = C.tile + off + (24 + 8 + (lane >> 3)) * LfwGeom::pitchC + 4 * (lane & 7);
    uint8_t *d = next_tile + off + (lane >> 3) * LfwGeom::pitchC + 4 * (lane & 7);
#pragma unroll
    for (int k = 0; k < 2; k++) lfw_stw(d + 4 * k * LfwGeom::pitchC, lfw_ldw(s + 4 * k * LfwGeom::pitchC));
  }
}

LF_HD void lfw_phase_vertical(const LfwCtx &C, int lane, bool dependent) {
  int p, g;
  lfw_role(lane, p, g);
  const int y = 4 * g;
  const bool uvsub = p != 0;
  if (!lf_row_valid(uvsub, y, C.sbr * 8, C.mi_rows)) return;
  const uint32_t d = C.dset[p * 8 + (y >> 3)];
  const uint32_t de = d & 0xff, di = d >> 8;
  const uint32_t se = (dependent ? de : ~de) & 0xfe, si = (dependent ? di : ~di) & 0xff;
  lfw_v_rows(*C.M, uvsub, y, lfw_at(C, p, y, 0), p ? LfwGeom::pitchC : LfwGeom::pitchY, se, si, C.lf_lim, C.lf_mblim);
}

LF_HD void lfw_phase_next_left_edge(const LfwCtx &C, int lane) {
  int p, g;
  lfw_role(lane, p, g);
  const int y = 4 * g;
  const bool uvsub = p != 0;
  if (!lf_row_valid(uvsub, y, C.sbr * 8, C.mi_rows)) return;
  lfw_v_rows(*C.Mn, uvsub, y, lfw_at(C, p, y, p ? 32 : 64), p ? LfwGeom::pitchC : LfwGeom::pitchY, 0x01, 0x00, C.lf_lim, C.lf_mblim);
}

// take over the samples the left neighbour's worker changed with our left edge (after F_edge(r, c-1))
LF_HD void lfw_phase_merge_left(const LfwCtx &C, int lane) {
  int p, g;
  lfw_role(lane, p, g);
  const int y = 4 * g;
  const bool uvsub = p != 0;
  uint32_t m16, m8, m4, m4i;
  lf_band_masks_v(*C.M, uvsub, y >> 3, m16, m8, m4, m4i);
  const int reach = lf_left_reach(m16, m8, m4);
  if (!reach) return;
  // bytes 0..reach-1 of the row: reach is 7, 3 or 2
  const uint32_t k0 = reach >= 4 ? 0xffffffffu : (reach == 3 ? 0x00ffffffu : 0x0000ffffu), k1 = reach == 7 ? 0x00ffffffu : 0u;
  for (int i = 0; i < 4; i++) {
    const lf_u2 v = lf_ldg<lf_u2>(C.g[p] + (size_t)(y + i) * C.stride[p]);
    uint8_t *t = lfw_at(C, p, y + i, 0);
    lfw_stw(t, lfw_sel(k0, v.x, lfw_ldw(t)));
    if (k1) lfw_stw(t + 4, lfw_sel(k1, v.y, lfw_ldw(t + 4)));
  }
}

// horizontal edges: part 0 = last block column, part 1 = all the others, part 2 = block columns
// 1 .. nb-2, part 3 = block column 0 (parts 2 + 3 = part 1: the left edge of the superblock and the
// vertical edges that depend on it normally stay inside block column 0, so the other columns need
// not wait for the left neighbour)
LF_HD void lfw_phase_horizontal(const LfwCtx &C, int lane, int part) {
  int p, g;
  lfw_role(lane, p, g);
  const int x = 4 * g;
  const bool uvsub = p != 0;
  const int n = p ? 32 : 64, nb = n >> 3, c = x >> 3;
  const bool mine = part == 0 ? c == nb - 1 : (part == 1 ? c < nb - 1 : (part == 2 ? (c > 0 && c < nb - 1) : c == 0));
  if (!mine) return;
  lfw_h_cols(*C.M, uvsub, x, lfw_at(C, p, 0, x), p ? LfwGeom::pitchC : LfwGeom::pitchY, C.sbr * 8, C.mi_rows, nb, C.lf_lim, C.lf_mblim);
}
// does a dependent set reach beyond block column 0 (operations k >= 2)?
LF_HD bool lfw_plan_deep(const uint16_t *dset) {
  uint32_t any = 0;
  for (int i = 0; i < 24; i++) any |= dset[i];
  return (any & 0xfefe) != 0;   // low byte: 8-aligned edges, high byte: inner edges, one bit per block column
}

// Stores (same split as lf_phase_store): part 0 = last block column of every plane (rows r0 .. n-1-keep)
// + the look-ahead columns (rows 0..n-1); part 1 = block columns 0..nb-2.  8 bytes per store; the
// lane -> (row, vector) mapping is fixed so that a row step is two pointer increments.
LF_HD void lfw_store_col(const uint8_t *t, int pitch, uint8_t *g, int stride, int r0, int r1, int dr) {
  // one 8-sample vector per row: rows r0, r0 + dr, ... < r1; t / g = tile / global address of that vector in row 0
  t += (r0 + 8) * pitch;
  g += (ptrdiff_t)r0 * stride;
  for (int r = r0; r < r1; r += dr, t += dr * pitch, g += (ptrdiff_t)dr * stride) {
    lf_u2 val;
    val.x = lfw_ldw(t); val.y = lfw_ldw(t + 4);
    lf_stg<lf_u2>(g, val);
  }
}
LF_HD void lfw_phase_store(const LfwCtx &C, int lane, int part, int keep) {
  const int top = C.has_top ? -8 : 0;
  if (part == 0) {
    // last block column: luma on lanes 0-15, U on 16-23, V on 24-31, one row per lane and step
    int p, g;
    lfw_role(lane, p, g);
    const int n = p ? 32 : 64, lanes = p ? 8 : 16, pitch = p ? LfwGeom::pitchC : LfwGeom::pitchY;
    const uint8_t *tp = C.tile + lfw_plane_off(p);
    lfw_store_col(tp + n - 8, pitch, C.g[p] + n - 8, C.stride[p], top + g, n - keep, lanes);
    if (C.has_next) lfw_store_col(tp + n, pitch, C.g[p] + n, C.stride[p], g, n, lanes);
  } else {
    // luma: 7 vectors per row, four rows per step on lanes 0-27
    if (lane < 28) {
      const int rl = lane / 7, v = lane - 7 * rl;
      lfw_store_col(C.tile + 8 * v, LfwGeom::pitchY, C.g[0] + 8 * v, C.stride[0], top + rl, 64 - keep, 4);
    }
    // chroma: 3 vectors per row, ten rows per step on lanes 0-29, U then V
    if (lane < 30) {
      const int rl = lane / 3, v = lane - 3 * rl;
#pragma unroll
      for (int p = 1; p < 3; p++)
        lfw_store_col(C.tile + lfw_plane_off(p) + 8 * v, LfwGeom::pitchC, C.g[p] + 8 * v, C.stride[p], top + rl, 32 - keep, 10);
    }
  }
}

}  // namespace vp9b
