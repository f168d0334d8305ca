// Loop filter: one persistent CTA per superblock COLUMN, superblocks ordered by data flow.
//
// libvpx filters superblock by superblock in raster order: all vertical edges of a superblock,
// then all its horizontal edges (vp9_loopfilter.c:1440-1467, vp9_filter_block_plane_ss00/ss11
// :1241-1422).  Two filter operations only have to keep that order when their sample footprints
// overlap.  Vertical-edge operations touch one sample row each, horizontal ones one sample column,
// so the order that matters is
//   * along a sample row: vertical edges left to right,
//   * down a sample column: horizontal edges top to bottom,
//   * vertical before horizontal wherever they overlap -- EXCEPT that the horizontal edges of the
//     last 8 columns of superblock c run before the left-most vertical edge of c+1 (which reaches
//     8 samples back into c).
// The horizontal edges chain down the frame (the top edge of every superblock needs the final
// bottom rows of the one above), so a worker owns a whole superblock column and keeps those 8 rows
// in shared memory from one superblock to the next: no global round trip on the long dependency.
// For superblock (r, c) -- tile = 64x64 (+ chroma) plus 8 look-ahead columns of (r, c+1) -- it runs
//   A   every vertical edge of c that does not (transitively) overlap the left edge of c
//       [that left edge is applied by worker c-1 on its look-ahead columns]
//   H7  horizontal edges of the last block column of c
//   E   the left vertical edge of c+1 on the look-ahead columns
//       -> store last block column + look-ahead, release F_edge(r, c)
//   D   after F_edge(r, c-1): take over the samples the left edge changed, then the vertical edges
//       that depend on them (rarely any: an 8/16-wide edge every 8+ samples does not chain)
//   H   horizontal edges of the other block columns -> store
// and carries the bottom 8 rows over as the next superblock's top rows.  The next superblock's
// samples are fetched while the current one is filtered.  If the dependent set D reaches the last
// block column ("chain"), D runs before H7 and the workers of that row serialise through F_edge;
// otherwise a worker publishes its edge before it looks at its left neighbour, and the columns
// advance side by side.  Workers take their column from a ticket counter, so a worker only ever
// waits for one that started before it.
//
// Filter arithmetic: vpx_dsp/loopfilter.c:18-357 (8-bit) == :360-743 evaluated at bd; selection
// and level quirks: filter_selectively_vert_row2 / _horiz (vp9_loopfilter.c:297-650).
//
// The per-thread functions are compiled for the device and, with VP9B_EMUL, for the host: the CPU
// emulation in tests/emul runs the same phase functions thread by thread in adversarial but legal
// schedules and is compared with the oracle (no GPU needed to validate the decomposition).
#pragma once
#include <stdint.h>

#include "vp9_b200.h"

#ifdef VP9B_EMUL
#include <algorithm>
#include <cstdlib>
#include <cstring>
#define LF_HD inline
namespace vp9b {
struct lf_u2 { uint32_t x, y; };
struct lf_u4 { uint32_t x, y, z, w; };
using std::max;
using std::min;
inline int lf_ffs(uint32_t v) { return v ? __builtin_ctz(v) + 1 : 0; }
inline int lf_clz(uint32_t v) { return v ? __builtin_clz(v) : 32; }
inline int lf_absd(int a, int b) { return a < b ? b - a : a - b; }
template <typename V> inline V lf_ldg(const void *p) { V v; memcpy(&v, p, sizeof(V)); return v; }
template <typename V> inline void lf_stg(void *p, const V &v) { memcpy(p, &v, sizeof(V)); }
}  // namespace vp9b
#else
#define LF_HD __device__ __forceinline__
namespace vp9b {
typedef uint2 lf_u2;
typedef uint4 lf_u4;
LF_HD int lf_ffs(uint32_t v) { return __ffs(v); }
LF_HD int lf_clz(uint32_t v) { return __clz(v); }
#ifdef VP9B_LF_NOSAD
LF_HD int lf_absd(int a, int b) { return a < b ? b - a : a - b; }
#else
LF_HD int lf_absd(int a, int b) { return (int)__sad(a, b, 0u); }   // |a - b| in one instruction
#endif
template <typename V> LF_HD V lf_ldg(const void *p) { return __ldcg((const V *)p); }   // L2: written by other SMs
template <typename V> LF_HD void lf_stg(void *p, const V &v) { __stcg((V *)p, v); }
}  // namespace vp9b
#endif

namespace vp9b {

LF_HD int lf_clampi(int v, int lo, int hi) { return min(max(v, lo), hi); }
LF_HD int lf_abs(int v) { return v < 0 ? -v : v; }
LF_HD int lf_sc_clamp(int t, int bd) {   // signed_char_clamp / signed_char_clamp_high
  const int lim = 128 << (bd - 8);
  return lf_clampi(t, -lim, lim - 1);
}

// One sample position across an edge.  s points at q0; step is the distance between taps.
template <typename Px>
LF_HD void lf_px(Px *s, int step, int size, int blimit, int limit, int thresh, int bd_) {
  const int bd = sizeof(Px) == 1 ? 8 : bd_;   // 8-bit samples: every shift below folds away
  const int sh = bd - 8;
  const int p3 = s[-4 * step], p2 = s[-3 * step], p1 = s[-2 * step], p0 = s[-step];
  const int q0 = s[0], q1 = s[step], q2 = s[2 * step], q3 = s[3 * step];
  const int lim16 = limit << sh, blim16 = blimit << sh, one = 1 << sh;
  // masks as max chains: no short-circuit branches, |p1-p0| and |q1-q0| shared by all of them
  const int d10p = lf_absd(p1, p0), d10q = lf_absd(q1, q0);
  const int dmax = max(max(max(lf_absd(p3, p2), lf_absd(p2, p1)), max(lf_absd(q2, q1), lf_absd(q3, q2))), max(d10p, d10q));
  const bool fm = dmax <= lim16 && lf_absd(p0, q0) * 2 + (lf_absd(p1, q1) >> 1) <= blim16;
  if (!fm) return;  // filter_mask == 0: every variant leaves the samples untouched
  bool flat = false;
  if (size >= 8)
    flat = max(max(max(d10p, d10q), max(lf_absd(p2, p0), lf_absd(q2, q0))), max(lf_absd(p3, p0), lf_absd(q3, q0))) <= one;
  if (size == 16 && flat) {
    const int p7 = s[-8 * step], p6 = s[-7 * step], p5 = s[-6 * step], p4 = s[-5 * step];
    const int q4 = s[4 * step], q5 = s[5 * step], q6 = s[6 * step], q7 = s[7 * step];
    const bool flat2 = max(max(max(lf_absd(p4, p0), lf_absd(q4, q0)), max(lf_absd(p5, p0), lf_absd(q5, q0))),
                           max(max(lf_absd(p6, p0), lf_absd(q6, q0)), max(lf_absd(p7, p0), lf_absd(q7, q0)))) <= one;
    if (flat2) {
      // 15-tap [1,1,1,1,1,1,1,2,1,1,1,1,1,1,1] with edge replication, evaluated from prefix
      // sums so that the 14 outputs are independent of each other (short dependency chain)
      const int v[16] = { p7, p6, p5, p4, p3, p2, p1, p0, q0, q1, q2, q3, q4, q5, q6, q7 };
      int S[17];
      S[0] = 0;
#pragma unroll
      for (int i = 0; i < 16; i++) S[i + 1] = S[i] + v[i];
#pragma unroll
      for (int k = 1; k <= 14; k++) {
        const int lo = k - 7 > 0 ? k - 7 : 0, hi = k + 7 < 15 ? k + 7 : 15;
        const int box = S[hi + 1] - S[lo] + (7 - k > 0 ? (7 - k) * p7 : 0) + (k - 8 > 0 ? (k - 8) * q7 : 0);
        s[(k - 8) * step] = (Px)((box + v[k] + 8) >> 4);
      }
      return;
    }
  }
  if (flat) {
    s[-3 * step] = (Px)((p3 + p3 + p3 + 2 * p2 + p1 + p0 + q0 + 4) >> 3);
    s[-2 * step] = (Px)((p3 + p3 + p2 + 2 * p1 + p0 + q0 + q1 + 4) >> 3);
    s[-1 * step] = (Px)((p3 + p2 + p1 + 2 * p0 + q0 + q1 + q2 + 4) >> 3);
    s[0] = (Px)((p2 + p1 + p0 + 2 * q0 + q1 + q2 + q3 + 4) >> 3);
    s[1 * step] = (Px)((p1 + p0 + q0 + 2 * q1 + q2 + q3 + q3 + 4) >> 3);
    s[2 * step] = (Px)((p0 + q0 + q1 + 2 * q2 + q3 + q3 + q3 + 4) >> 3);
    return;
  }
  // filter4 (mask is all ones here)
  const int off = 0x80 << sh;
  const int ps1 = p1 - off, ps0 = p0 - off, qs0 = q0 - off, qs1 = q1 - off;
  const int t16 = thresh << sh;
  const int hev = max(d10p, d10q) > t16 ? -1 : 0;
  int f = lf_sc_clamp(ps1 - qs1, bd) & hev;
  f = lf_sc_clamp(f + 3 * (qs0 - ps0), bd);
  const int f1 = lf_sc_clamp(f + 4, bd) >> 3, f2 = lf_sc_clamp(f + 3, bd) >> 3;
  s[0] = (Px)(lf_sc_clamp(qs0 - f1, bd) + off);
  s[-step] = (Px)(lf_sc_clamp(ps0 + f2, bd) + off);
  f = ((f1 + 1) >> 1) & ~hev;
  s[step] = (Px)(lf_sc_clamp(qs1 - f, bd) + off);
  s[-2 * step] = (Px)(lf_sc_clamp(ps1 + f, bd) + off);
}

// ------------------------------------------------------------------------------------------
// geometry of the shared-memory tile: per plane rows -8..n-1 (8 halo rows above), columns 0..n+7
// (8 look-ahead columns of the next superblock); row pitch in 32-bit words is odd so that a
// thread-per-row or thread-per-column walk is free of bank conflicts
template <typename Px> struct LfVecT;
template <> struct LfVecT<uint8_t> { typedef lf_u2 T; };    // 8 samples
template <> struct LfVecT<uint16_t> { typedef lf_u4 T; };

LF_HD void lf_smem_put(void *d, lf_u2 v) { uint32_t *w = (uint32_t *)d; w[0] = v.x; w[1] = v.y; }
LF_HD void lf_smem_put(void *d, lf_u4 v) { uint32_t *w = (uint32_t *)d; w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w; }
LF_HD void lf_smem_get(const void *s, lf_u2 &v) { const uint32_t *w = (const uint32_t *)s; v.x = w[0]; v.y = w[1]; }
LF_HD void lf_smem_get(const void *s, lf_u4 &v) { const uint32_t *w = (const uint32_t *)s; v.x = w[0]; v.y = w[1]; v.z = w[2]; v.w = w[3]; }

template <typename Px, bool SUB> struct LfGeom {
  static constexpr int nC = SUB ? 32 : 64;                              // chroma superblock size in samples
  static constexpr int NT = SUB ? 128 : 192;                            // threads = sample rows (columns) of the three planes
  static constexpr int pitchY = sizeof(Px) == 1 ? 76 : 74;              // samples per tile row
  static constexpr int pitchC = SUB ? (sizeof(Px) == 1 ? 44 : 42) : pitchY;
  static constexpr int offU = 72 * pitchY, offV = offU + (nC + 8) * pitchC;
  static constexpr int kTileSamples = offV + (nC + 8) * pitchC + 8;
  static constexpr size_t kTileBytes = (((size_t)kTileSamples * sizeof(Px)) + 15) & ~(size_t)15;
  static constexpr int kInterior = 64 * 8 + 2 * nC * (nC >> 3);       // 8-sample vectors of the three planes' interiors
  static constexpr int kPre = kInterior / NT + 1;                       // per thread: interior slots + one look-ahead vector
  static_assert(kInterior % NT == 0, "interior vectors must divide evenly over the threads");
};

// everything one CTA needs to know about its superblock
template <typename Px, bool SUB> struct LfCtx {
  Px *tile;                 // shared-memory tile (all planes)
  Px *g[3];                 // global address of sample (0,0) of this superblock in each plane
  int stride[3];            // plane strides in samples
  const vp9b_lfm *M, *Mn;   // masks of this superblock and of the next one in the row
  const uint8_t *lf_lim, *lf_mblim;
  uint16_t *dset;           // [3][8] per plane, per 8-row band: vertical operations that depend on the left edge (lf_plan_unit)
  int sbr, sbc, sb_cols, mi_rows, bd;
  bool has_next, has_top;
};

template <typename Px, bool SUB>
LF_HD void lf_role(int tid, int &plane, int &idx) {   // thread -> (plane, sample row or column)
  constexpr int nC = LfGeom<Px, SUB>::nC;
  if (tid < 64) { plane = 0; idx = tid; }
  else if (tid < 64 + nC) { plane = 1; idx = tid - 64; }
  else { plane = 2; idx = tid - 64 - nC; }
}
template <typename Px, bool SUB>
LF_HD Px *lf_tile_at(const LfCtx<Px, SUB> &C, int plane, int y, int x) {   // y in -8..n-1, x in 0..n+7
  typedef LfGeom<Px, SUB> G;
  const int off = plane == 0 ? 0 : (plane == 1 ? G::offU : G::offV);
  const int pitch = plane ? G::pitchC : G::pitchY;
  return C.tile + off + (y + 8) * pitch + x;
}

// ------------------------------------------------------------------------------------------
// Which vertical operations of one 8-row band have to wait for the left edge of the superblock.
// Operation k sits at x = 4k: even k = the 8-aligned edge of block column k/2 (16-, 8- or 4-wide by
// mask), odd k = the inner 4x4 edge.  Sample extents per size (q side / p side symmetric):
// reads 8 / 4 / 4, writes 7 / 3 / 2.  An operation joins the set when its reads overlap what the
// set has written or its writes overlap what the set has read (conservative extents).
// m16/m8/m4/m4i: mask bits of this band, one per block column; nb block columns.
LF_HD uint32_t lf_dependent_set(uint32_t m16, uint32_t m8, uint32_t m4, uint32_t m4i, int nb, bool left_applies) {
  if (!left_applies) return 0;
  const int s0 = (m16 & 1) ? 16 : ((m8 & 1) ? 8 : ((m4 & 1) ? 4 : 0));
  if (!s0) return 0;
  int max_w = s0 == 16 ? 6 : (s0 == 8 ? 2 : 1);   // last sample written / read by the set so far
  int max_r = s0 == 16 ? 7 : 3;
  uint32_t d = 1;
  for (int k = 1; k < 2 * nb; k++) {
    const int c = k >> 1, x = 4 * k;
    int sz;
    if (k & 1) sz = ((m4i >> c) & 1) ? 4 : 0;
    else sz = ((m16 >> c) & 1) ? 16 : (((m8 >> c) & 1) ? 8 : (((m4 >> c) & 1) ? 4 : 0));
    if (!sz) continue;
    const int rlo = x - (sz == 16 ? 8 : 4), wlo = x - (sz == 16 ? 7 : (sz == 8 ? 3 : 2));
    if (rlo <= max_w || wlo <= max_r) {
      d |= 1u << k;
      max_w = max(max_w, x + (sz == 16 ? 6 : (sz == 8 ? 2 : 1)));
      max_r = max(max_r, x + (sz == 16 ? 7 : 3));
    }
  }
  return d;
}
// samples of a row that the left edge of a superblock may have changed on the q side (columns 0..n-1)
LF_HD int lf_left_reach(uint32_t m16, uint32_t m8, uint32_t m4) {
  return (m16 & 1) ? 7 : ((m8 & 1) ? 3 : ((m4 & 1) ? 2 : 0));
}

// mask bits of band r8 of a plane (uvsub: 4:2:0 chroma, 4 block columns per band)
LF_HD void lf_band_masks_v(const vp9b_lfm &M, bool uvsub, int r8, uint32_t &m16, uint32_t &m8, uint32_t &m4, uint32_t &m4i) {
  if (!uvsub) {
    m16 = (uint32_t)(M.left_y[2] >> (r8 * 8)) & 0xff;
    m8 = (uint32_t)(M.left_y[1] >> (r8 * 8)) & 0xff;
    m4 = (uint32_t)(M.left_y[0] >> (r8 * 8)) & 0xff;
    m4i = (uint32_t)(M.int_4x4_y >> (r8 * 8)) & 0xff;
  } else {
    m16 = (uint32_t)(M.left_uv[2] >> (r8 * 4)) & 0xf;
    m8 = (uint32_t)(M.left_uv[1] >> (r8 * 4)) & 0xf;
    m4 = (uint32_t)(M.left_uv[0] >> (r8 * 4)) & 0xf;
    m4i = (uint32_t)(M.int_4x4_uv >> (r8 * 4)) & 0xf;
  }
}

// Vertical edges of sample row y of one plane of the superblock described by M; `row` = sample
// x = 0 of that superblock in this row.  sel_edge / sel_inner select the block columns whose
// 8-aligned / inner 4x4 edge is applied now (filter_selectively_vert_row2, vp9_loopfilter.c:297-451:
// 16, then 8, then 4, then the inner edge; a 16-wide edge set in both bands of a pair takes the
// first band's level for both).
template <typename Px>
LF_HD void lf_v_row(const vp9b_lfm &M, bool uvsub, int y, Px *row, uint32_t sel_edge, uint32_t sel_inner,
                    const uint8_t *lf_lim, const uint8_t *lf_mblim, int bd) {
  const int r8 = y >> 3, pr = r8 & ~1;
  uint32_t m16, m8, m4, m4i, p16, p8, p4, p4i;
  lf_band_masks_v(M, uvsub, r8, m16, m8, m4, m4i);
  lf_band_masks_v(M, uvsub, r8 ^ 1, p16, p8, p4, p4i);
  const uint32_t m16n = m16 & p16;   // set in both bands of the pair
  m16 &= sel_edge; m8 &= sel_edge; m4 &= sel_edge; m4i &= sel_inner;
  uint32_t all = m16 | m8 | m4 | m4i;
  if (!all) return;
  const int lrow = uvsub ? 2 * r8 : r8, prow = uvsub ? 2 * pr : pr, lsh = uvsub ? 16 : 8;
  uint64_t lv_own, lv_pair;
  memcpy(&lv_own, &M.lfl_y[lrow * 8], 8);
  memcpy(&lv_pair, &M.lfl_y[prow * 8], 8);
#ifndef VP9B_LF_MULTICALL
  while (all) {
    const int c = lf_ffs(all) - 1;
    all &= all - 1;
    const int lo = (int)(lv_own >> (c * lsh)) & 0xff;
    const int lv16 = ((m16n >> c) & 1) ? ((int)(lv_pair >> (c * lsh)) & 0xff) : lo;
    Px *s = row + c * 8;
    // the operations of this block column in the reference's order (16, 8, 4, inner) through ONE
    // inlined copy of the filter (rows of different bands reconverge in it; a third less code than a copy
    // per size at the same speed; VP9B_LF_MULTICALL keeps the copies for A/B)
    uint32_t ops = ((m16 >> c) & 1) | (((m8 >> c) & 1) << 1) | (((m4 >> c) & 1) << 2) | (((m4i >> c) & 1) << 3);
#pragma unroll 1
    while (ops) {
      const int k = lf_ffs(ops) - 1;
      ops &= ops - 1;
      const int lv = k == 0 ? lv16 : lo;
      lf_px<Px>(s + (k == 3 ? 4 : 0), 1, k == 0 ? 16 : (k == 1 ? 8 : 4), lf_mblim[lv], lf_lim[lv], lv >> 4, bd);
    }
  }
#else
  while (all) {
    const int c = lf_ffs(all) - 1;
    all &= all - 1;
    const int lo = (int)(lv_own >> (c * lsh)) & 0xff;
    Px *s = row + c * 8;
    if ((m16 >> c) & 1) {
      const int lv = ((m16n >> c) & 1) ? ((int)(lv_pair >> (c * lsh)) & 0xff) : lo;
      lf_px<Px>(s, 1, 16, lf_mblim[lv], lf_lim[lv], lv >> 4, bd);
    }
    if ((m8 >> c) & 1) lf_px<Px>(s, 1, 8, lf_mblim[lo], lf_lim[lo], lo >> 4, bd);
    if ((m4 >> c) & 1) lf_px<Px>(s, 1, 4, lf_mblim[lo], lf_lim[lo], lo >> 4, bd);
    if ((m4i >> c) & 1) lf_px<Px>(s + 4, 1, 4, lf_mblim[lo], lf_lim[lo], lo >> 4, bd);
  }
#endif
}
// is the band pair of sample row y inside the frame (the vertical pass walks band pairs)
LF_HD bool lf_row_valid(bool uvsub, int y, int mi_row, int mi_rows) {
  const int pr = (y >> 3) & ~1;
  return mi_row + (uvsub ? 2 * pr : pr) < mi_rows;
}

// Horizontal edges (filter_selectively_horiz, vp9_loopfilter.c:453-650: priority 16 > 8 > 4, inner
// edge unless 16; a run of 16-wide edges is paired greedily from the left and the second block of a
// pair is filtered with the first block's thresholds).
struct LfHMasks {
  uint64_t a16, a8, a4, ai;   // above masks of the plane, frame-edge rules applied
  int rsh, lstep, nrow;       // log2(block columns per mask row), level step, block rows inside the frame
};
LF_HD LfHMasks lf_h_masks(const vp9b_lfm &M, bool uvsub, int mi_row, int mi_rows, int nb) {
  LfHMasks H;
  if (!uvsub) {
    H.a16 = M.above_y[2]; H.a8 = M.above_y[1]; H.a4 = M.above_y[0]; H.ai = M.int_4x4_y;
    H.rsh = 3; H.lstep = 1;
  } else {
    H.a16 = M.above_uv[2]; H.a8 = M.above_uv[1]; H.a4 = M.above_uv[0]; H.ai = M.int_4x4_uv;
    H.rsh = 2; H.lstep = 2;
  }
  if (mi_row == 0) {   // no edge on the frame's top row (the inner 4x4 edge stays)
    const uint64_t keep = ~(uint64_t)((1u << (1 << H.rsh)) - 1);
    H.a16 &= keep; H.a8 &= keep; H.a4 &= keep;
  }
  // rows of this superblock inside the frame; 4:2:0 drops the inner edge of the last odd row
  const int rows_left = mi_rows - mi_row;
  H.nrow = uvsub ? min(nb, (rows_left + 1) >> 1) : min(nb, rows_left);
  if (uvsub && (rows_left & 1) && rows_left < 8) H.ai &= ~((uint64_t)0xf << (4 * (rows_left >> 1)));  // skip_border_4x4_r
  return H;
}
// the edges of block row r in sample column x (block column c = x >> 3); col = sample (0, x)
template <typename Px>
LF_HD void lf_h_unit(const LfHMasks &H, const vp9b_lfm &M, int x, int r, Px *col, int ts, const uint8_t *lf_lim,
                     const uint8_t *lf_mblim, int bd) {
  const int c = x >> 3, bit = (r << H.rsh) + c;
  const uint32_t k16 = (uint32_t)(H.a16 >> bit) & 1, k8 = (uint32_t)(H.a8 >> bit) & 1, k4 = (uint32_t)(H.a4 >> bit) & 1,
                 ki = (uint32_t)(H.ai >> bit) & 1;
  if (!(k16 | k8 | k4 | ki)) return;
  Px *s = col + r * 8 * ts;
  const uint8_t *lfl = &M.lfl_y[(r * H.lstep) * 8];
  const int lvl = lfl[c * H.lstep];
#ifndef VP9B_LF_MULTICALL
  int lv = lvl, size = k16 ? 16 : (k8 ? 8 : (k4 ? 4 : 0));
  if (k16) {
    const uint32_t row16 = (uint32_t)(H.a16 >> (r << H.rsh)) & 0xff;
    const uint32_t below = ~row16 & ((1u << c) - 1);          // zero bits below c
    const int run = below ? c - 1 - (31 - lf_clz(below)) : c;
    if (run & 1) lv = lfl[(c - 1) * H.lstep];
  }
  // edge, then the inner edge (unless 16-wide), through one inlined copy of the filter
  uint32_t ops = (size ? 1u : 0u) | ((!k16 && ki) ? 2u : 0u);
#pragma unroll 1
  while (ops) {
    const int k = lf_ffs(ops) - 1;
    ops &= ops - 1;
    const int l = k ? lvl : lv;
    lf_px<Px>(s + (k ? 4 * ts : 0), ts, k ? 4 : size, lf_mblim[l], lf_lim[l], l >> 4, bd);
  }
#else
  if (k16) {
    const uint32_t row16 = (uint32_t)(H.a16 >> (r << H.rsh)) & 0xff;
    const uint32_t below = ~row16 & ((1u << c) - 1);          // zero bits below c
    const int run = below ? c - 1 - (31 - lf_clz(below)) : c;
    const int lv = (run & 1) ? lfl[(c - 1) * H.lstep] : lvl;
    lf_px<Px>(s, ts, 16, lf_mblim[lv], lf_lim[lv], lv >> 4, bd);
  } else if (k8) {
    lf_px<Px>(s, ts, 8, lf_mblim[lvl], lf_lim[lvl], lvl >> 4, bd);
  } else if (k4) {
    lf_px<Px>(s, ts, 4, lf_mblim[lvl], lf_lim[lvl], lvl >> 4, bd);
  }
  if (!k16 && ki) lf_px<Px>(s + 4 * ts, ts, 4, lf_mblim[lvl], lf_lim[lvl], lvl >> 4, bd);
#endif
}
// one thread walks a whole sample column: only the block rows that have an edge are visited
template <typename Px>
LF_HD void lf_h_col(const vp9b_lfm &M, bool uvsub, int x, Px *col, int ts, int mi_row, int mi_rows, int nb,
                    const uint8_t *lf_lim, const uint8_t *lf_mblim, int bd) {
  const LfHMasks H = lf_h_masks(M, uvsub, mi_row, mi_rows, nb);
  const int c = x >> 3;
  const uint64_t colsel = uvsub ? 0x1111ull : 0x0101010101010101ull;
  uint64_t todo = ((H.a16 | H.a8 | H.a4 | H.ai) >> c) & colsel;
  if (H.nrow < nb) todo &= ((uint64_t)1 << (H.nrow << H.rsh)) - 1;
  while (todo) {
    const int bit0 = (todo & 0xffffffffu) ? lf_ffs((uint32_t)todo) - 1 : 32 + lf_ffs((uint32_t)(todo >> 32)) - 1;
    todo &= todo - 1;
    lf_h_unit<Px>(H, M, x, bit0 >> H.rsh, col, ts, lf_lim, lf_mblim, bd);
  }
}
// ------------------------------------------------------------------------------------------
// phase functions: each is called by every thread of the CTA (tid = 0..NT-1) between barriers

// Dependent sets of one superblock: unit u = 0 .. 8 + 2 * nbC - 1 is a (plane, band); writes
// dset[plane * 8 + band] and returns whether the set reaches the last block column (chain).
template <typename Px, bool SUB>
LF_HD bool lf_plan_unit(const vp9b_lfm &M, bool has_left, int u, uint16_t *dset) {
  constexpr int nbC = LfGeom<Px, SUB>::nC >> 3;
  int plane, band;
  if (u < 8) { plane = 0; band = u; }
  else if (u < 8 + nbC) { plane = 1; band = u - 8; }
  else { plane = 2; band = u - 8 - nbC; }
  const bool uvsub = plane && SUB;
  const int nb = plane ? nbC : 8;
  uint32_t m16, m8, m4, m4i;
  lf_band_masks_v(M, uvsub, band, m16, m8, m4, m4i);
  const uint32_t d = lf_dependent_set(m16, m8, m4, m4i, nb, has_left);
  // stored per block column: low byte = 8-aligned edges, high byte = inner edges (the phase functions
  // of every sample row of the band read it: split once here, not once per row and phase)
  uint32_t de = 0, di = 0;
  for (int c = 0; c < 8; c++) {
    de |= ((d >> (2 * c)) & 1) << c;
    di |= ((d >> (2 * c + 1)) & 1) << c;
  }
  dset[plane * 8 + band] = (uint16_t)(de | (di << 8));
  return (d >> (2 * nb - 2)) != 0;
}
template <typename Px, bool SUB> struct LfPlanUnits { static constexpr int value = 8 + 2 * (LfGeom<Px, SUB>::nC >> 3); };

// Fetch of a superblock's interior (rows 0..n-1, columns 0..n-1, + the look-ahead columns) into
// registers, and its later commit to a tile buffer: issued one superblock ahead so that the loads
// fly while the current superblock is filtered.  Nobody else writes these samples before this
// worker is done with them, except the left edge of the superblock (merged later).
// g[p] = global address of sample (0,0) of that superblock.
template <typename Px, bool SUB>
LF_HD void lf_phase_prefetch(Px *const g[3], const int stride[3], bool has_next, int tid,
                             typename LfVecT<Px>::T (&pre)[LfGeom<Px, SUB>::kPre]) {
  typedef LfGeom<Px, SUB> G;
  typedef typename LfVecT<Px>::T Vec;
  constexpr int vY = 64 * 8, vC = G::nC * (G::nC >> 3);
#pragma unroll
  for (int k = 0; k < G::kPre - 1; k++) {
    int p = 0, i = tid + k * G::NT;
    if (i >= vY) { i -= vY; p = 1; if (i >= vC) { i -= vC; p = 2; } }
    const int lg = (p && SUB) ? 2 : 3;                 // log2(vectors per row)
    const int r = i >> lg, v = i & ((1 << lg) - 1);
    pre[k] = lf_ldg<Vec>(g[p] + (size_t)r * stride[p] + v * 8);
  }
  if (has_next) {
    int p, r;
    lf_role<Px, SUB>(tid, p, r);
    pre[G::kPre - 1] = lf_ldg<Vec>(g[p] + (size_t)r * stride[p] + (p ? G::nC : 64));
  }
}
template <typename Px, bool SUB>
LF_HD void lf_phase_commit(const LfCtx<Px, SUB> &C, int tid, const typename LfVecT<Px>::T (&pre)[LfGeom<Px, SUB>::kPre]) {
  typedef LfGeom<Px, SUB> G;
  constexpr int vY = 64 * 8, vC = G::nC * (G::nC >> 3);
#pragma unroll
  for (int k = 0; k < G::kPre - 1; k++) {
    int p = 0, i = tid + k * G::NT;
    if (i >= vY) { i -= vY; p = 1; if (i >= vC) { i -= vC; p = 2; } }
    const int lg = (p && SUB) ? 2 : 3;
    lf_smem_put(lf_tile_at(C, p, i >> lg, (i & ((1 << lg) - 1)) * 8), pre[k]);
  }
  if (C.has_next) {
    int p, r;
    lf_role<Px, SUB>(tid, p, r);
    lf_smem_put(lf_tile_at(C, p, r, p ? G::nC : 64), pre[G::kPre - 1]);
  }
}

// bottom 8 rows of the finished superblock -> top rows (-8..-1) of the next one's tile buffer
template <typename Px, bool SUB>
LF_HD void lf_phase_carry(const LfCtx<Px, SUB> &C, Px *next_tile, int tid) {
  typedef LfGeom<Px, SUB> G;
  typedef typename LfVecT<Px>::T Vec;
  constexpr int hY = 8 * 8, hC = 8 * (G::nC >> 3);    // vectors per plane
  LfCtx<Px, SUB> N = C;
  N.tile = next_tile;
  for (int i = tid; i < hY + 2 * hC; i += G::NT) {
    int p = 0, k = i;
    if (k >= hY) { k -= hY; p = 1; if (k >= hC) { k -= hC; p = 2; } }
    const int lg = (p && SUB) ? 2 : 3;
    const int r = k >> lg, v = k & ((1 << lg) - 1), n = p ? G::nC : 64;
    Vec t;
    lf_smem_get(lf_tile_at(C, p, n - 8 + r, v * 8), t);
    lf_smem_put(lf_tile_at(N, p, r - 8, v * 8), t);
  }
}

// vertical edges of this superblock: the independent ones (dependent = false) or the ones that
// had to wait for the left edge (dependent = true); never the left edge itself
template <typename Px, bool SUB>
LF_HD void lf_phase_vertical(const LfCtx<Px, SUB> &C, int tid, bool dependent) {
  int p, y;
  lf_role<Px, SUB>(tid, p, y);
  const bool uvsub = p && SUB;
  if (!lf_row_valid(uvsub, y, C.sbr * 8, C.mi_rows)) return;
  const uint32_t d = C.dset[p * 8 + (y >> 3)];
  const uint32_t de = d & 0xff, di = d >> 8;   // dependent set: 8-aligned / inner edges per block column
  const uint32_t se = (dependent ? de : ~de) & 0xfe, si = (dependent ? di : ~di) & 0xff;
  lf_v_row<Px>(*C.M, uvsub, y, lf_tile_at(C, p, y, 0), se, si, C.lf_lim, C.lf_mblim, C.bd);
}

// the left edge of the next superblock, on the look-ahead columns
template <typename Px, bool SUB>
LF_HD void lf_phase_next_left_edge(const LfCtx<Px, SUB> &C, int tid) {
  typedef LfGeom<Px, SUB> G;
  int p, y;
  lf_role<Px, SUB>(tid, p, y);
  const bool uvsub = p && SUB;
  if (!lf_row_valid(uvsub, y, C.sbr * 8, C.mi_rows)) return;
  lf_v_row<Px>(*C.Mn, uvsub, y, lf_tile_at(C, p, y, p ? G::nC : 64), 0x01, 0x00, C.lf_lim, C.lf_mblim, C.bd);
}

// take over the samples the left neighbour's CTA changed with our left edge (after F_edge(r, c-1))
template <typename Px, bool SUB>
LF_HD void lf_phase_merge_left(const LfCtx<Px, SUB> &C, int tid) {
  typedef typename LfVecT<Px>::T Vec;
  int p, y;
  lf_role<Px, SUB>(tid, p, y);
  const bool uvsub = p && SUB;
  uint32_t m16, m8, m4, m4i;
  lf_band_masks_v(*C.M, uvsub, y >> 3, m16, m8, m4, m4i);
  const int reach = lf_left_reach(m16, m8, m4);
  if (!reach) return;
  const Vec v = lf_ldg<Vec>(C.g[p] + (size_t)y * C.stride[p]);
  Px tmp[8];
  memcpy(tmp, &v, sizeof(v));
  Px *t = lf_tile_at(C, p, y, 0);
  for (int i = 0; i < reach; i++) t[i] = tmp[i];
}

// horizontal edges, one thread per sample column: part 0 = last block column, part 1 = the others
template <typename Px, bool SUB>
LF_HD void lf_phase_horizontal(const LfCtx<Px, SUB> &C, int tid, int part) {
  typedef LfGeom<Px, SUB> G;
  int p, x;
  lf_role<Px, SUB>(tid, p, x);
  const bool uvsub = p && SUB;
  const int n = p ? G::nC : 64, nb = n >> 3;
  if (((x >> 3) == nb - 1) != (part == 0)) return;
  lf_h_col<Px>(*C.M, uvsub, x, lf_tile_at(C, p, 0, x), p ? G::pitchC : G::pitchY, C.sbr * 8, C.mi_rows, nb, C.lf_lim,
               C.lf_mblim, C.bd);
}

// stores of rows r0 .. n-1-keep (r0 = -8 with rows above, else 0; keep = 8 while the bottom rows
// still wait for the next superblock's top edge, 0 in the last superblock row):
// part 0 = last block column, plus the look-ahead columns (all rows 0..n-1: the left edge of c+1 is
// done with them); part 1 = block columns 0..nb-2.  A thread owns one 8-sample column of a plane
// and walks down the rows with a fixed step.
template <typename Px, bool SUB>
LF_HD void lf_store_plane(const LfCtx<Px, SUB> &C, int plane, int lg, int idx, int nthreads, int part, int keep) {
  typedef LfGeom<Px, SUB> G;
  typedef typename LfVecT<Px>::T Vec;
  const int n = plane ? G::nC : 64, vpr = 1 << lg;        // vectors per row
  const int v = idx & (vpr - 1), step = nthreads >> lg;
  if ((v == vpr - 1) != (part == 0)) return;
  const int pitch = plane ? G::pitchC : G::pitchY, stride = C.stride[plane];
  int r = (C.has_top ? -8 : 0) + (idx >> lg);
  const Px *t = lf_tile_at(C, plane, r, v * 8);
  Px *g = C.g[plane] + (ptrdiff_t)r * stride + v * 8;
  for (; r < n - keep; r += step, t += step * pitch, g += (ptrdiff_t)step * stride) {
    Vec val;
    lf_smem_get(t, val);
    lf_stg<Vec>(g, val);
  }
}
template <typename Px, bool SUB>
LF_HD void lf_phase_store(const LfCtx<Px, SUB> &C, int tid, int part, int keep) {
  typedef LfGeom<Px, SUB> G;
  typedef typename LfVecT<Px>::T Vec;
  if (SUB) {
    lf_store_plane<Px, SUB>(C, 0, 3, tid, G::NT, part, keep);
    // the two chroma planes share the threads: 64 threads each, 4 vectors per row
    lf_store_plane<Px, SUB>(C, 1 + (tid >> 6), 2, tid & 63, 64, part, keep);
  } else {
    int p, i;
    lf_role<Px, SUB>(tid, p, i);
    lf_store_plane<Px, SUB>(C, p, 3, i, 64, part, keep);
  }
  if (part == 0 && C.has_next) {
    int p, r;
    lf_role<Px, SUB>(tid, p, r);
    const int n = p ? G::nC : 64;
    Vec v;
    lf_smem_get(lf_tile_at(C, p, r, n), v);
    lf_stg<Vec>(C.g[p] + (size_t)r * C.stride[p] + n, v);
  }
}

}  // namespace vp9b
