// VP9 reconstruction kernels for sm_100a (B200) + the vp9gpu_* C ABI (include/vp9_b200.h).
//
// One packed buffer per frame is uploaded; reference frames stay resident in HBM.  Stages per
// frame (or per batch of frames of up to 32 contexts), all enqueued on the context stream:
//   0  k_expand_blocks    block records -> dense per-8x8 motion map (L2 resident)
//   1  k_inter_pred_sb    8-tap sub-pel MC incl. compound + scaled refs, one CTA per superblock,
//                         reference windows staged by TMA                  (a1-a3)
//   2  k_residual         inverse transform + add for inter blocks          (a6-a7)
//   3  k_intra_residual   inverse transforms of intra blocks into residual planes
//      k_intra            intra prediction + residual, superblock-row wavefront (a4-a7)
//   4  k_lf_sb            lpf 4/8/16 in libvpx's normative order, one worker per superblock column,
//                         launched programmatically behind k_intra and following it one superblock
//                         row behind (a9-a10; k_lf_warp: packed-arithmetic variant, opt-in)
// Integer byte work, HBM/L2 bound -- no tensor cores by design (bit exactness).
#include <cuda.h>
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <mutex>
#include <string>
#include <vector>

#include "vp9_b200.h"
#include "vp9_tables.h"
#include "vp9_txfm.cuh"
#include "vp9_lf_sb.cuh"
#include "vp9_lf_warp.cuh"

// occupancy targets of the two register-hungry kernels (A/B builds override them)
#ifndef VP9B_IRES_MINB
#define VP9B_IRES_MINB 8
#endif
#ifndef VP9B_MC_MINB
#define VP9B_MC_MINB 5
#endif
#ifndef VP9B_INTRA_MINB
#define VP9B_INTRA_MINB 1
#endif
#ifndef VP9B_INTRA_DENSE
#define VP9B_INTRA_DENSE 4      // CTAs per SM of the k_intra variant used for large batches
#endif
#ifndef VP9B_LF_MINB
#define VP9B_LF_MINB 5
#endif

namespace vp9b {

// ------------------------------------------------------------------------------------------
// device-side constant tables
__device__ int16_t d_scan[(16 + 64 + 256) * 3 + 1024];  // [4x4: def,row,col][8x8..][16x16..][32x32 def]
__device__ int16_t d_filters16[4 * 16 * 8];

__host__ __device__ inline int scan_offset(int tx_size, int tx_type) {
  // tx_type: 0 default, 1 row scan, 2 col scan, 3 default
  const int t = (tx_type == 1) ? 1 : (tx_type == 2 ? 2 : 0);
  switch (tx_size) {
    case 0: return t * 16;
    case 1: return 48 + t * 64;
    case 2: return 48 + 192 + t * 256;
    default: return 48 + 192 + 768;
  }
}

struct MiEntry {  // 16 bytes, dense per 8x8
  int16_t mv[2][2];
  uint8_t ref0, ref1, filter, bsize;
  uint32_t aux;  // sub-8x8: index into submv table
};

struct PlaneGeom {
  void *ptr;
  int stride;  // in samples
};

struct RefGeom {
  PlaneGeom pl[3];
  int slot;        // frame slot (selects the TMA descriptors)
  int w[3], h[3];  // crop size per plane
  int x_scale_fp, y_scale_fp, x_step_q4, y_step_q4;
};

struct FrameParams {
  PlaneGeom dst[3];
  RefGeom ref[3];
  int mi_cols, mi_rows, sb_cols, sb_rows;
  int ss_x, ss_y, bd, lossless;
  const vp9b_block *blocks;
  int n_blocks;
  const vp9b_submv *submv;
  const vp9b_txb *txb_inter;
  int n_txb_inter;
  int n_txb_intra[3];
  void *resid[3];      // residual planes of the intra blocks (int16 / int32), same stride as dst
  int n_txb_size[4];   // TXB_INTER is grouped by transform size
  const vp9b_txb *txb_intra[3];
  const vp9b_sbrange *sb_intra;  // [3][nsb]
  const void *coef;
  const vp9b_lfm *lfm;
  const vp9b_lfm_uv *lfm_uv;    // chroma masks of 4:2:2 / 4:4:0 frames (nullptr otherwise)
  MiEntry *mi;
  const CUtensorMap *tmaps;     // [slot][plane][class] TMA descriptors of the frame pool (nullptr: no staging)
  uint32_t *prog_intra;         // [3][sb_rows] progress counters: prog_base + superblocks done in the row
  uint32_t *prog_done;          // [3][sb_rows] the same, counted once the superblock is completely in memory (read by the loop filter)
  uint32_t prog_base;           // per-frame epoch offset: counters never need clearing between frames
  // Work tickets: a wavefront CTA / warp takes its row from a counter instead of blockIdx, so a
  // unit that waits for a lower-numbered one can only be running if that one was started before
  // it -- forward progress does not depend on the order in which the hardware dispatches CTAs.
  uint32_t *ticket_lf, *ticket_intra;
  uint32_t *lf_flags;           // [superblocks] loop-filter data-flow flags F_edge (k_lf_sb)
  uint32_t lf_epoch;            // value a flag takes when set in this frame
  uint32_t ticket_lf_base, ticket_intra_base;
  uint8_t lf_lim[64], lf_mblim[64];
  // which stages this frame takes part in, and its own motion-compensation grid (a launch covers a
  // batch of frames of different contexts; its grid is the maximum over the batch)
  int do_inter, do_res, do_intra, do_lf;
};

// Frames of up to kMaxBatch different contexts that go through one launch per stage: the hardware
// has 32 work queues, so more than 32 concurrent streams need several frames per launch.
constexpr int kMaxBatch = VP9B_MAX_BATCH;   // 32 x 736 bytes of kernel parameters (limit: 32764)
// Two parameter-block sizes: a launch copies the whole block (736 bytes per frame), and 23 KB of
// parameters cost ~10 us of host time per launch -- single frames and pairs use the small block.
template <int NB> struct FrameBatchT {
  static constexpr int kN = NB;
  FrameParams f[NB];
};
typedef FrameBatchT<kMaxBatch> FrameBatch;
typedef FrameBatchT<2> FrameBatchSmall;

__device__ __constant__ uint8_t d_bw8[13] = { 1, 1, 1, 1, 1, 2, 2, 2, 4, 4, 4, 8, 8 };
__device__ __constant__ uint8_t d_bh8[13] = { 1, 1, 1, 1, 2, 1, 2, 4, 2, 4, 8, 4, 8 };

#define VP9B_HD __host__ __device__ __forceinline__
VP9B_DEV int clampi(int v, int lo, int hi) { return min(max(v, lo), hi); }

// ------------------------------------------------------------------------------------------
// stage 0: block list -> dense mi map
template <typename FB>
__global__ void k_expand_blocks(const __grid_constant__ FB B) {
  const FrameParams &P = B.f[blockIdx.z];
  if (!P.do_inter) return;
  const vp9b_block *__restrict__ blocks = P.blocks;
  MiEntry *__restrict__ mi = P.mi;
  const int n = P.n_blocks, mi_cols = P.mi_cols, mi_rows = P.mi_rows;
  // 8 threads per block: thread l fills row l of the block's footprint in the 8x8 map
  const int t = blockIdx.x * blockDim.x + threadIdx.x, i = t >> 3, l = t & 7;
  if (i >= n) return;
  const vp9b_block b = blocks[i];
  MiEntry e;
  e.mv[0][0] = b.mv[0][0]; e.mv[0][1] = b.mv[0][1]; e.mv[1][0] = b.mv[1][0]; e.mv[1][1] = b.mv[1][1];
  e.ref0 = b.ref0; e.ref1 = b.ref1; e.filter = b.filter; e.bsize = b.bsize;
  e.aux = 0;
  if (b.ref0 && b.bsize < 3) {
    uint32_t idx;
    memcpy(&idx, &b.mv[0][0], 4);
    e.aux = idx;
  }
  const int bw = min((int)d_bw8[b.bsize], mi_cols - b.mi_col), bh = min((int)d_bh8[b.bsize], mi_rows - b.mi_row);
  if (l < bh)
    for (int x = 0; x < bw; x++) mi[(size_t)(b.mi_row + l) * mi_cols + b.mi_col + x] = e;
}

// ------------------------------------------------------------------------------------------
// stage 1: inter prediction.  Position maths of dec_build_inter_predictors
// (vp9_decodeframe.c:556-690): unscaled blocks take mv*2^(1-ss) in 1/16 pel; scaled blocks
// clamp the MV (clamp_mv_to_umv_border_sb, vp9_reconinter.c:90-110), scale the block origin and
// the MV (vp9_scale.c:16-43) and step by x_step_q4.  Every fetch clamps to the reference crop
// size (== build_mc_border, :432-504).  Filters: vp9_filter.c:15-82; rounding and clip after each
// 1-D pass: vpx_convolve.c:22-82; compound average (a+b+1)>>1: :58-59.
struct PredPos {
  int x0, y0;      // integer sample position of the unit origin in the reference plane
  int sx, sy;      // sub-pel phase at the origin (1/16)
  int xs, ys;      // step per output sample (16 = unscaled)
};

VP9B_DEV PredPos pred_position(const FrameParams &P, const RefGeom &R, int plane, int mi_row0, int mi_col0, int bw8,
                               int bh8, int xoff, int yoff, int mv_row, int mv_col) {
  const int ss_x = plane ? P.ss_x : 0, ss_y = plane ? P.ss_y : 0;
  const int mb_to_left = -(mi_col0 * 64), mb_to_top = -(mi_row0 * 64);
  PredPos p;
  int mvr, mvc;
  if (R.x_step_q4 != 16 || R.y_step_q4 != 16) {
    const int mb_to_right = (P.mi_cols - bw8 - mi_col0) * 64, mb_to_bottom = (P.mi_rows - bh8 - mi_row0) * 64;
    const int bw_px = 4 * ((bw8 << 1) >> ss_x), bh_px = 4 * ((bh8 << 1) >> ss_y);
    const int spel_left = (4 + bw_px) << 4, spel_right = spel_left - 16;
    const int spel_top = (4 + bh_px) << 4, spel_bottom = spel_top - 16;
    int qrow = (int16_t)(mv_row * (1 << (1 - ss_y))), qcol = (int16_t)(mv_col * (1 << (1 - ss_x)));
    qcol = clampi(qcol, mb_to_left * (1 << (1 - ss_x)) - spel_left, mb_to_right * (1 << (1 - ss_x)) + spel_right);
    qrow = clampi(qrow, mb_to_top * (1 << (1 - ss_y)) - spel_top, mb_to_bottom * (1 << (1 - ss_y)) + spel_bottom);
    const int x_start = (-mb_to_left >> (3 + ss_x)), y_start = (-mb_to_top >> (3 + ss_y));
    p.x0 = (int)(((long long)(x_start + xoff) * R.x_scale_fp) >> 14);
    p.y0 = (int)(((long long)(y_start + yoff) * R.y_scale_fp) >> 14);
    const int x_off_q4 = (int)(((long long)((mi_col0 * 8 + xoff) << 4) * R.x_scale_fp) >> 14) & 15;
    const int y_off_q4 = (int)(((long long)((mi_row0 * 8 + yoff) << 4) * R.y_scale_fp) >> 14) & 15;
    mvr = (int)(((long long)qrow * R.y_scale_fp) >> 14) + y_off_q4;
    mvc = (int)(((long long)qcol * R.x_scale_fp) >> 14) + x_off_q4;
    p.xs = R.x_step_q4;
    p.ys = R.y_step_q4;
  } else {
    p.x0 = (-mb_to_left >> (3 + ss_x)) + xoff;
    p.y0 = (-mb_to_top >> (3 + ss_y)) + yoff;
    mvr = mv_row * (1 << (1 - ss_y));
    mvc = mv_col * (1 << (1 - ss_x));
    p.xs = p.ys = 16;
  }
  p.sx = mvc & 15;
  p.sy = mvr & 15;
  p.x0 += mvc >> 4;
  p.y0 += mvr >> 4;
  return p;
}

// Generic (any block size / scaling / bit depth) per-sample predictor: the exact path for scaled
// references (x_step_q4 != 16); unscaled prediction goes through the 4 x H unit kernels below.
template <typename Px>
VP9B_DEV int predict_sample(const FrameParams &P, const RefGeom &R, int plane, const PredPos &pp, int dx, int dy,
                            int filter) {
  const Px *__restrict__ ref = (const Px *)R.pl[plane].ptr;
  const int rstride = R.pl[plane].stride, rw = R.w[plane], rh = R.h[plane];
  const int xq = pp.sx + dx * pp.xs, yq = pp.sy + dy * pp.ys;
  const int16_t *fx = d_filters16 + (filter * 16 + (xq & 15)) * 8;
  const int16_t *fy = d_filters16 + (filter * 16 + (yq & 15)) * 8;
  const int bx = pp.x0 - 3 + (xq >> 4), by = pp.y0 - 3 + (yq >> 4);
  const int maxv = (1 << P.bd) - 1;
  int cx[8];
#pragma unroll
  for (int k = 0; k < 8; k++) cx[k] = clampi(bx + k, 0, rw - 1);
  int acc = 0;
#pragma unroll
  for (int j = 0; j < 8; j++) {
    const Px *row = ref + (size_t)clampi(by + j, 0, rh - 1) * rstride;
    int s = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) s += (int)row[cx[k]] * fx[k];
    s = clampi((s + 64) >> 7, 0, maxv);
    acc += s * fy[j];
  }
  return clampi((acc + 64) >> 7, 0, maxv);
}

// ------------------------------------------------------------------------------------------
// stage 1 fast path: 8-bit, unscaled references.  One thread predicts a 4-wide x 8-tall strip
// (two 4x4 units when the two halves belong to different prediction units).  Reference rows
// are fetched as aligned 32-bit words through L1 (neighbouring threads overlap heavily), the
// 8-tap filters run as two DP4A (u8 samples x s8 taps) per output with byte windows built by
// funnel shifts, the horizontally filtered rows are clipped and kept packed, transposed 4x4 by
// PRMT so that the vertical pass is DP4A too, and results are packed with cvt.pack.sat.  The
// compound average is a packed rounding byte average.  Units that touch the frame edge run the
// same code with clamped row loads; only scaled references fall back to predict_sample().
__device__ __constant__ uint32_t d_taps4[4 * 16 * 2];  // [filter][phase][lo,hi] packed s8 taps

VP9B_DEV int dp4a_us(uint32_t a, uint32_t b, int c) {
  int d;
  asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
VP9B_DEV uint32_t pack_sat_u8x4(int a, int b, int c, int d) {  // bytes (a,b,c,d), a lowest, each >>7 done by caller
  uint32_t lo, r;
  asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, 0;" : "=r"(lo) : "r"(d), "r"(c));
  asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(b), "r"(a), "r"(lo));
  return r;
}
VP9B_DEV uint32_t avg_u8x4(uint32_t a, uint32_t b) {  // per byte (a + b + 1) >> 1
  return (a | b) - (((a ^ b) >> 1) & 0x7f7f7f7fu);
}

// Predict a 4 x H unit (H = 4 or 8) from one reference; returns H packed rows in out[].
// One code path for all sub-pel classes: lanes whose phase is 0 in a direction simply sit out
// that direction's filter, so a warp pays for the most expensive class it contains, not the sum.
// 16 bytes of reference row y starting at (4-aligned, possibly negative) column xa, every sample
// position clamped to the reference crop size -- the frame-edge emulation of build_mc_border
// (vp9_decodeframe.c:432-466).  Out of line: only units that touch the frame edge come here.
__device__ __noinline__ uint4 mc_load_row_clamped(const uint8_t *__restrict__ ref, int rstride, int y, int xa, int rw,
                                                  int rh) {
  const uint8_t *row = ref + (ptrdiff_t)clampi(y, 0, rh - 1) * rstride;
  uint32_t w[4];
#pragma unroll
  for (int j = 0; j < 4; j++) {
    const int x = xa + 4 * j;
    if (x >= 0 && x + 3 < rw) {
      w[j] = __ldg((const uint32_t *)(row + x));
    } else {
      uint32_t v = 0;
#pragma unroll
      for (int k = 0; k < 4; k++) v |= (uint32_t)__ldg(row + clampi(x + k, 0, rw - 1)) << (8 * k);
      w[j] = v;
    }
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}

// `ref` + rstride describe either a reference plane in global memory or a window of it staged in
// shared memory (generic address); (xs, ys) = position of the first tap (x0 - 3, y0 - 3) relative
// to it.  The clamped loads of `edge` units only exist for global planes.
template <int H>
VP9B_DEV void mc_unit_fast(const uint8_t *ref, int rstride, int xs, int ys, int sx, int sy, int filter,
                           bool edge, int rw, int rh, uint32_t *out) {
  constexpr int R = H + 7;  // rows y0-3 .. y0+H+3
  const uint32_t txl = d_taps4[(filter * 16 + sx) * 2], txh = d_taps4[(filter * 16 + sx) * 2 + 1];
  const uint32_t tyl = d_taps4[(filter * 16 + sy) * 2], tyh = d_taps4[(filter * 16 + sy) * 2 + 1];
  const int o = (xs & 3) * 8;
  const uint32_t *p = (const uint32_t *)(ref + (ptrdiff_t)ys * rstride + (xs & ~3));
  uint32_t im[R + 1];
#pragma unroll
  for (int r = 0; r < R; r++) {
    uint32_t v = 0;
    if (sy || (r >= 3 && r < 3 + H)) {
      uint32_t a0, a1, a2, a3;
      if (!edge) {
        a0 = p[0]; a1 = p[1]; a2 = p[2]; a3 = p[3];
      } else {
        const uint4 a = mc_load_row_clamped(ref, rstride, ys + r, xs & ~3, rw, rh);
        a0 = a.x; a1 = a.y; a2 = a.z; a3 = a.w;
      }
      const uint32_t v0 = __funnelshift_r(a0, a1, o), v1 = __funnelshift_r(a1, a2, o), v2 = __funnelshift_r(a2, a3, o);
      if (sx) {
        int s[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const uint32_t w0 = __funnelshift_r(v0, v1, 8 * j), w1 = __funnelshift_r(v1, v2, 8 * j);
          s[j] = dp4a_us(w1, txh, dp4a_us(w0, txl, 64)) >> 7;
        }
        v = pack_sat_u8x4(s[0], s[1], s[2], s[3]);
      } else {
        v = __funnelshift_r(v0, v1, 24);  // samples x0 .. x0+3
      }
    }
    im[r] = v;
    p = (const uint32_t *)((const uint8_t *)p + rstride);
  }
  im[R] = 0;
  if (sy) {
    // transpose 4x4 byte blocks: col[c][g] = samples of column c, rows 4g..4g+3
    constexpr int G = (R + 3) / 4;
    uint32_t col[4][G];
#pragma unroll
    for (int g = 0; g < G; g++) {
      const uint32_t r0 = im[4 * g], r1 = im[4 * g + 1], r2 = (4 * g + 2 <= R) ? im[4 * g + 2] : 0,
                     r3 = (4 * g + 3 <= R) ? im[4 * g + 3] : 0;
      const uint32_t t0 = __byte_perm(r0, r1, 0x5140), t1 = __byte_perm(r2, r3, 0x5140);
      const uint32_t t2 = __byte_perm(r0, r1, 0x7362), t3 = __byte_perm(r2, r3, 0x7362);
      col[0][g] = __byte_perm(t0, t1, 0x5410);
      col[1][g] = __byte_perm(t0, t1, 0x7632);
      col[2][g] = __byte_perm(t2, t3, 0x5410);
      col[3][g] = __byte_perm(t2, t3, 0x7632);
    }
#pragma unroll
    for (int y = 0; y < H; y++) {
      int s[4];
#pragma unroll
      for (int c = 0; c < 4; c++) {
        const int g = y >> 2, sh = (y & 3) * 8;
        const uint32_t w0 = __funnelshift_r(col[c][g], col[c][g + 1], sh);
        const uint32_t w1 = __funnelshift_r(col[c][g + 1], (g + 2 < G) ? col[c][g + 2] : 0u, sh);
        s[c] = dp4a_us(w1, tyh, dp4a_us(w0, tyl, 64)) >> 7;
      }
      out[y] = pack_sat_u8x4(s[0], s[1], s[2], s[3]);
    }
  } else {
#pragma unroll
    for (int y = 0; y < H; y++) out[y] = im[y + 3];
  }
}

// is the whole fetch footprint of a 4xH unit (incl. the aligned word reads) inside the plane?
VP9B_DEV bool mc_inside(int x0, int y0, int H, int rw, int rh) {
  return x0 - 3 >= 0 && y0 - 3 >= 0 && x0 + 8 <= rw && y0 + H + 4 <= rh;
}

// Exact per-sample path for one 4-sample row of a unit (frame edge / scaled reference).  Kept out
// of line and rolled: it is rare, and inlining it bloats the kernel past the instruction cache.
__device__ __noinline__ uint32_t mc_row_slow(const FrameParams &P, const RefGeom &R, int plane, const MiEntry &e,
                                             int mi_row, int mi_col, int x, int y, int yy, int mvr, int mvc) {
  const int ss_x = plane ? P.ss_x : 0, ss_y = plane ? P.ss_y : 0;
  const bool sub8 = e.bsize < 3;
  const int bw8 = d_bw8[e.bsize], bh8 = d_bh8[e.bsize];
  const int mi_col0 = mi_col & ~(bw8 - 1), mi_row0 = mi_row & ~(bh8 - 1);
  const int bx = x - ((mi_col0 * 8) >> ss_x), by = y - ((mi_row0 * 8) >> ss_y);
  const int xoff = sub8 ? (bx & ~3) : 0, yoff = sub8 ? (by & ~3) : 0;
  const PredPos pp = pred_position(P, R, plane, mi_row0, mi_col0, bw8, bh8, xoff, yoff, mvr, mvc);
  uint32_t w = 0;
#pragma unroll 1
  for (int xx = 0; xx < 4; xx++)
    w |= (uint32_t)predict_sample<uint8_t>(P, R, plane, pp, bx - xoff + xx, by - yoff + yy, e.filter) << (8 * xx);
  return w;
}

// ------------------------------------------------------------------------------------------
// stage 1 fast path, high bit depth (vpx_highbd_convolve8*, vpx_convolve.c:292-535): the same
// 4 x H units on u16 samples.  Reference rows are fetched as aligned 32-bit words (two samples),
// the window is normalised to an even start with one variable funnel shift per word, and the
// 8-tap filters run as four DP2A (2 x u16 samples times 2 x s8 taps) per output.  The vertical
// pass pairs rows with PRMT so that it is DP2A too.
VP9B_DEV int dp2a_lo(uint32_t a, uint32_t b, int c) {
  int d;
  asm("dp2a.lo.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
VP9B_DEV int dp2a_hi(uint32_t a, uint32_t b, int c) {
  int d;
  asm("dp2a.hi.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
VP9B_DEV uint32_t avg_u16x2(uint32_t a, uint32_t b) {  // per half (a + b + 1) >> 1
  return (a | b) - (((a ^ b) >> 1) & 0x7fff7fffu);
}
struct Row6 { uint32_t w[6]; };

// 12 samples of reference row y starting at (even, possibly negative) column xa, positions
// clamped to the reference crop size (highbd build_mc_border, vp9_decodeframe.c:469-504)
__device__ __noinline__ Row6 mc16_load_row_clamped(const uint16_t *__restrict__ ref, int rstride, int y, int xa, int rw,
                                                   int rh) {
  const uint16_t *row = ref + (ptrdiff_t)clampi(y, 0, rh - 1) * rstride;
  Row6 r;
#pragma unroll
  for (int j = 0; j < 6; j++) {
    const int x = xa + 2 * j;
    const uint32_t lo = __ldg(row + clampi(x, 0, rw - 1)), hi = __ldg(row + clampi(x + 1, 0, rw - 1));
    r.w[j] = lo | (hi << 16);
  }
  return r;
}

template <int H>
VP9B_DEV void mc_unit_fast16(const uint16_t *ref, int rstride, int xs, int ys, int sx, int sy, int filter,
                             bool edge, int rw, int rh, int maxv, uint2 *out) {
  constexpr int R = H + 7;  // rows y0-3 .. y0+H+3; (xs, ys) = first tap relative to `ref` (plane or staged window)
  const uint32_t txl = d_taps4[(filter * 16 + sx) * 2], txh = d_taps4[(filter * 16 + sx) * 2 + 1];
  const uint32_t tyl = d_taps4[(filter * 16 + sy) * 2], tyh = d_taps4[(filter * 16 + sy) * 2 + 1];
  const int par = xs & 1, xa = xs - par;
  const uint16_t *p = ref + (ptrdiff_t)ys * rstride + xa;
  uint2 im[R];
#pragma unroll
  for (int r = 0; r < R; r++) {
    uint2 v = make_uint2(0, 0);
    if (sy || (r >= 3 && r < 3 + H)) {
      uint32_t w[6];
      if (!edge) {
#pragma unroll
        for (int i = 0; i < 6; i++) w[i] = ((const uint32_t *)p)[i];
      } else {
        const Row6 t = mc16_load_row_clamped(ref, rstride, ys + r, xa, rw, rh);
#pragma unroll
        for (int i = 0; i < 6; i++) w[i] = t.w[i];
      }
      // n[i] = samples (xs + 2i, xs + 2i + 1)
      uint32_t n[6];
#pragma unroll
      for (int i = 0; i < 5; i++) n[i] = __funnelshift_r(w[i], w[i + 1], par * 16);
      n[5] = w[5] >> (par * 16);
      if (sx) {
        uint32_t o[5];  // o[i] = samples (xs + 2i + 1, xs + 2i + 2)
#pragma unroll
        for (int i = 0; i < 5; i++) o[i] = __funnelshift_r(n[i], n[i + 1], 16);
        int s[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const uint32_t *a = (j & 1) ? o : n;
          const int b = j >> 1;
          int acc = dp2a_lo(a[b], txl, 64);
          acc = dp2a_hi(a[b + 1], txl, acc);
          acc = dp2a_lo(a[b + 2], txh, acc);
          acc = dp2a_hi(a[b + 3], txh, acc);
          s[j] = clampi(acc >> 7, 0, maxv);
        }
        v = make_uint2((uint32_t)s[0] | ((uint32_t)s[1] << 16), (uint32_t)s[2] | ((uint32_t)s[3] << 16));
      } else {
        v = make_uint2(__funnelshift_r(n[1], n[2], 16), __funnelshift_r(n[2], n[3], 16));  // samples x0 .. x0+3
      }
    }
    im[r] = v;
    p += rstride;
  }
  if (sy) {
#pragma unroll
    for (int y = 0; y < H; y++) {
      int s[4];
#pragma unroll
      for (int c = 0; c < 4; c++) {
        const uint32_t sel = (c & 1) ? 0x7632u : 0x5410u;
        uint32_t pr[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
          const uint2 a = im[y + 2 * k], b = im[y + 2 * k + 1];
          pr[k] = __byte_perm(c < 2 ? a.x : a.y, c < 2 ? b.x : b.y, sel);
        }
        int acc = dp2a_lo(pr[0], tyl, 64);
        acc = dp2a_hi(pr[1], tyl, acc);
        acc = dp2a_lo(pr[2], tyh, acc);
        acc = dp2a_hi(pr[3], tyh, acc);
        s[c] = clampi(acc >> 7, 0, maxv);
      }
      out[y] = make_uint2((uint32_t)s[0] | ((uint32_t)s[1] << 16), (uint32_t)s[2] | ((uint32_t)s[3] << 16));
    }
  } else {
#pragma unroll
    for (int y = 0; y < H; y++) out[y] = im[y + 3];
  }
}

__device__ __noinline__ uint2 mc16_row_slow(const FrameParams &P, const RefGeom &R, int plane, const MiEntry &e,
                                            int mi_row, int mi_col, int x, int y, int yy, int mvr, int mvc) {
  const int ss_x = plane ? P.ss_x : 0, ss_y = plane ? P.ss_y : 0;
  const bool sub8 = e.bsize < 3;
  const int bw8 = d_bw8[e.bsize], bh8 = d_bh8[e.bsize];
  const int mi_col0 = mi_col & ~(bw8 - 1), mi_row0 = mi_row & ~(bh8 - 1);
  const int bx = x - ((mi_col0 * 8) >> ss_x), by = y - ((mi_row0 * 8) >> ss_y);
  const int xoff = sub8 ? (bx & ~3) : 0, yoff = sub8 ? (by & ~3) : 0;
  const PredPos pp = pred_position(P, R, plane, mi_row0, mi_col0, bw8, bh8, xoff, yoff, mvr, mvc);
  uint32_t s[4];
#pragma unroll 1
  for (int xx = 0; xx < 4; xx++)
    s[xx] = (uint32_t)predict_sample<uint16_t>(P, R, plane, pp, bx - xoff + xx, by - yoff + yy, e.filter);
  return make_uint2(s[0] | (s[1] << 16), s[2] | (s[3] << 16));
}

// ------------------------------------------------------------------------------------------
// stage 1 driver: one CTA per 64x64 superblock, reference windows of the big prediction blocks
// staged in shared memory by TMA.
//
// A thread predicts one 4 x 8 luma unit and one chroma unit of the superblock.  Prediction blocks of
// 16x16 and larger (most of the area of real streams) are cut into square staging units of
// S = min(width, height) samples; for each unit and reference one thread issues three
// cp.async.bulk.tensor.2d copies -- the (S + 7)^2 window of the luma plane and of both chroma
// planes, at the motion vector's integer offset (rounded down to a 16-byte column) -- into a bump-allocated region of the CTA's
// staging buffer, all completing on one mbarrier.  The 4 x H units then read their taps from the
// staged window (generic loads: the same code path reads global memory for small blocks, for units
// at the frame edge -- TMA fills out-of-range samples with zeros, libvpx replicates the border,
// vp9_decodeframe.c:432-504 -- and for scaled references).  The window is read from HBM / L2 once,
// 128 bytes at a time, instead of once per 4-wide unit through L1.
template <typename Px> struct McStage {
  static constexpr int kBudget = 24576;     // staging buffer, samples
  // TMA wants the first column of a box on a 16-byte boundary (measured on B200: an unaligned inner
  // coordinate raises an illegal-instruction fault), so a window starts at the aligned column at or
  // before its first tap: box width = alignment slack (15 / 7 samples) + S + 7, rounded up to 16 bytes
  static constexpr int kAlign = 16 / (int)sizeof(Px);     // samples
  VP9B_HD static int boxw(int S) {
    return sizeof(Px) == 1 ? (S == 64 ? 96 : (S == 32 ? 64 : (S == 16 ? 48 : 32))) : (S == 64 ? 80 : (S == 32 ? 48 : (S == 16 ? 32 : 24)));
  }
  VP9B_DEV static int cls(int S) { return S == 64 ? 3 : (S == 32 ? 2 : (S == 16 ? 1 : 0)); }   // tensor-map class
  VP9B_DEV static int region(int S) { return (boxw(S) * (S + 7) + 127) & ~127; }                // samples, 128-aligned
};

VP9B_DEV uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
VP9B_DEV void mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
VP9B_DEV void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
VP9B_DEV void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
VP9B_DEV void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
VP9B_DEV void tma_load_2d(void *dst, const CUtensorMap *map, int x, int y, uint64_t *bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(smem_u32(dst)), "l"(map), "r"(x), "r"(y), "r"(smem_u32(bar)) : "memory");
}

// Scaled reference, 8-bit, one 4 x H unit (vpx_scaled_2d == vpx_convolve8_c with x/y_step_q4 != 16,
// vpx_dsp/vpx_convolve.c:22-82,153-203): the unit's own intermediate block -- every reference row
// between the first and the last row the vertical pass touches, filtered horizontally at the four
// columns' positions and phases, clipped -- goes to a per-thread column of shared memory (the
// staging buffer, unused in frames with scaled references), then the vertical pass picks eight
// consecutive intermediate rows per output row.  Same arithmetic as predict_sample(), but a reference
// row is filtered once instead of once per vertical tap (8x fewer loads and horizontal MACs).
// Returns false (caller takes the per-sample path) when the window leaves the reference or the
// step is larger than the scratch allows.  pp = position of the BLOCK origin, (dx0, dy0) = the unit
// inside the block.  scr: this thread's column, rows 128 words apart.
constexpr int kScaledRows = 48;
template <int H>
VP9B_DEV bool mc_unit_scaled(const RefGeom &R, int plane, const PredPos &pp, int dx0, int dy0, int filter, uint32_t *scr,
                             uint32_t *out) {
  const int rw = R.w[plane], rh = R.h[plane], rstride = R.pl[plane].stride;
  int ix[4], ph[4];
#pragma unroll
  for (int c = 0; c < 4; c++) {
    const int xq = pp.sx + (dx0 + c) * pp.xs;
    ix[c] = pp.x0 - 3 + (xq >> 4);
    ph[c] = xq & 15;
  }
  const int yq0 = pp.sy + dy0 * pp.ys, yql = pp.sy + (dy0 + H - 1) * pp.ys;
  const int iy0 = pp.y0 - 3 + (yq0 >> 4), nrows = (yql >> 4) - (yq0 >> 4) + 8;
  if (nrows > kScaledRows || ix[0] < 0 || ix[3] + 8 > rw || iy0 < 0 || iy0 + nrows > rh) return false;
  // horizontal pass: three aligned words per column and row (neighbouring lanes share them in L1)
  uint32_t txl[4], txh[4];
  int sh[4];
  const uint8_t *col[4];
#pragma unroll
  for (int c = 0; c < 4; c++) {
    txl[c] = d_taps4[(filter * 16 + ph[c]) * 2];
    txh[c] = d_taps4[(filter * 16 + ph[c]) * 2 + 1];
    sh[c] = (ix[c] & 3) * 8;
    col[c] = (const uint8_t *)R.pl[plane].ptr + (ptrdiff_t)iy0 * rstride + (ix[c] & ~3);
  }
#pragma unroll 2
  for (int j = 0; j < nrows; j++) {
    int v[4];
#pragma unroll
    for (int c = 0; c < 4; c++) {
      const uint32_t *p = (const uint32_t *)(col[c] + (ptrdiff_t)j * rstride);
      const uint32_t a0 = __ldg(p), a1 = __ldg(p + 1), a2 = __ldg(p + 2);
      const uint32_t lo = __funnelshift_r(a0, a1, sh[c]), hi = __funnelshift_r(a1, a2, sh[c]);
      // phase 0 is the tap 128, stored as 127 in the packed table: add the centre sample once more
      int sum = dp4a_us(lo, txl[c], dp4a_us(hi, txh[c], 64)) + (ph[c] ? 0 : (int)(lo >> 24));
      v[c] = sum >> 7;
    }
    scr[j * 128] = pack_sat_u8x4(v[0], v[1], v[2], v[3]);
  }
  // vertical pass
#pragma unroll
  for (int r = 0; r < H; r++) {
    const int yq = pp.sy + (dy0 + r) * pp.ys, jr = (yq >> 4) - (yq0 >> 4), py = yq & 15;
    const uint32_t tyl = d_taps4[(filter * 16 + py) * 2], tyh = d_taps4[(filter * 16 + py) * 2 + 1];
    uint32_t m[8];
#pragma unroll
    for (int k = 0; k < 8; k++) m[k] = scr[(jr + k) * 128];
    // 4x4 byte transposes: m[0..3] / m[4..7] -> per column the four rows' samples
    const uint32_t a0 = __byte_perm(m[0], m[1], 0x5140), a1 = __byte_perm(m[2], m[3], 0x5140);
    const uint32_t a2 = __byte_perm(m[0], m[1], 0x7362), a3 = __byte_perm(m[2], m[3], 0x7362);
    const uint32_t b0 = __byte_perm(m[4], m[5], 0x5140), b1 = __byte_perm(m[6], m[7], 0x5140);
    const uint32_t b2 = __byte_perm(m[4], m[5], 0x7362), b3 = __byte_perm(m[6], m[7], 0x7362);
    const uint32_t lo[4] = { __byte_perm(a0, a1, 0x5410), __byte_perm(a0, a1, 0x7632), __byte_perm(a2, a3, 0x5410), __byte_perm(a2, a3, 0x7632) };
    const uint32_t hi[4] = { __byte_perm(b0, b1, 0x5410), __byte_perm(b0, b1, 0x7632), __byte_perm(b2, b3, 0x5410), __byte_perm(b2, b3, 0x7632) };
    int v[4];
#pragma unroll
    for (int c = 0; c < 4; c++) v[c] = (dp4a_us(lo[c], tyl, dp4a_us(hi[c], tyh, 64)) + (py ? 0 : (int)(lo[c] >> 24))) >> 7;
    out[r] = pack_sat_u8x4(v[0], v[1], v[2], v[3]);
  }
  return true;
}

// what a unit needs to know about its block's staged windows
struct McWin {
  int off[2];     // offset (samples) of the unit's luma window in the staging buffer per reference, -1: not staged
  int bx, by;     // luma position of the staging unit
  int S;          // its size
};

template <typename Px> struct McOut;
template <> struct McOut<uint8_t> { typedef uint32_t T; };     // 4 samples of a unit row
template <> struct McOut<uint16_t> { typedef uint2 T; };

// one 4 x H unit of plane `plane` at (x, y): all references, compound average (vpx_convolve_avg,
// vpx_convolve.c:58-59).  stage == nullptr: nothing staged for this unit.
template <typename Px, int H>
VP9B_DEV void mc_unit(const FrameParams &P, const Px *stage, const McWin &W, int plane, const MiEntry &e, int mi_row, int mi_col,
                      int x, int y, typename McOut<Px>::T *out) {
  typedef typename McOut<Px>::T Row;
  const int ss_x = plane ? P.ss_x : 0, ss_y = plane ? P.ss_y : 0;
  const bool sub8 = e.bsize < 3;
  const int maxv = (1 << P.bd) - 1;
#pragma unroll 1
  for (int r = 0; r < (e.ref1 ? 2 : 1); r++) {
    const RefGeom &R = P.ref[(r ? e.ref1 : e.ref0) - 1];
    int mvr, mvc;
    if (sub8) {
      // average_split_mvs (vp9_reconinter.c:90-113): the block index is the running count of the
      // plane's own 4x4 grid inside the 8x8 block (dec_build_inter_predictors_sb, dfc:702-718)
      const vp9b_submv &sm = P.submv[e.aux];
      if (ss_x && ss_y) { mvr = sm.chroma[r][0]; mvc = sm.chroma[r][1]; }
      else {
        const int k = ((y & ((8 >> ss_y) - 1)) >> 2) * (2 >> ss_x) + ((x & ((8 >> ss_x) - 1)) >> 2);
        mvr = sm.luma[k][r][0]; mvc = sm.luma[k][r][1];
        if (ss_x | ss_y) {     // 4:2:2: blocks k, k+1; 4:4:0: blocks k, k+2 (mi_mv_pred_q2)
          const int k2 = k + (ss_x ? 1 : 2);
          const int sr = mvr + sm.luma[k2][r][0], sc = mvc + sm.luma[k2][r][1];
          mvr = (sr < 0 ? sr - 1 : sr + 1) / 2;
          mvc = (sc < 0 ? sc - 1 : sc + 1) / 2;
        }
      }
    } else {
      mvr = e.mv[r][0];
      mvc = e.mv[r][1];
    }
    Row pr[H];
    if (R.x_step_q4 == 16 && R.y_step_q4 == 16) {
      const int smvr = mvr * (1 << (1 - ss_y)), smvc = mvc * (1 << (1 - ss_x));
      const int x0 = x + (smvc >> 4), y0 = y + (smvr >> 4);
      const Px *src = (const Px *)R.pl[plane].ptr;
      int stride = R.pl[plane].stride, xs = x0 - 3, ys = y0 - 3;
      bool edge;
      if (stage && W.off[r] >= 0) {
        // window of the staging unit in this plane: origin = unit position + integer MV - 3
        const int Sp = W.S >> ss_x;
        int off = W.off[r];
        if (plane) off += McStage<Px>::region(W.S) + (plane - 1) * McStage<Px>::region(Sp);
        src = stage + off;
        stride = McStage<Px>::boxw(Sp);
        // first tap of the staging unit in this plane sits (x0 - 3 of the unit's first column) mod alignment into the box
        xs = x - (W.bx >> ss_x) + (((W.bx >> ss_x) + (smvc >> 4) - 3) & (McStage<Px>::kAlign - 1));
        ys = y - (W.by >> ss_y);
        edge = false;
      } else if (sizeof(Px) == 1) {
        edge = !(xs >= 0 && ys >= 0 && x0 + 8 <= R.w[plane] && y0 + H + 4 <= R.h[plane]);
      } else {   // aligned word reads: columns x0-4 .. x0+8
        edge = !(x0 - 4 >= 0 && ys >= 0 && x0 + 9 <= R.w[plane] && y0 + H + 4 <= R.h[plane]);
      }
      if (sizeof(Px) == 1)
        mc_unit_fast<H>((const uint8_t *)src, stride, xs, ys, smvc & 15, smvr & 15, e.filter, edge, R.w[plane], R.h[plane], (uint32_t *)pr);
      else
        mc_unit_fast16<H>((const uint16_t *)src, stride, xs, ys, smvc & 15, smvr & 15, e.filter, edge, R.w[plane], R.h[plane], maxv,
                          (uint2 *)pr);
    } else {  // scaled reference
      bool done = false;
      if (sizeof(Px) == 1 && stage && !P.tmaps) {
        // the staging buffer is free in frames with scaled references: per-thread scratch column
        const int bw8 = d_bw8[e.bsize], bh8 = d_bh8[e.bsize];
        const int mi_col0 = mi_col & ~(bw8 - 1), mi_row0 = mi_row & ~(bh8 - 1);
        const int bx = x - ((mi_col0 * 8) >> ss_x), by = y - ((mi_row0 * 8) >> ss_y);
        const int xoff = sub8 ? (bx & ~3) : 0, yoff = sub8 ? (by & ~3) : 0;
        const PredPos pp = pred_position(P, R, plane, mi_row0, mi_col0, bw8, bh8, xoff, yoff, mvr, mvc);
        done = mc_unit_scaled<H>(R, plane, pp, bx - xoff, by - yoff, e.filter, (uint32_t *)const_cast<Px *>(stage) + threadIdx.x, (uint32_t *)pr);
      }
      if (!done) {   // exact per-sample path (frame edges, high bit depth, very large steps)
#pragma unroll
        for (int yy = 0; yy < H; yy++) {
          if (sizeof(Px) == 1) ((uint32_t *)pr)[yy] = mc_row_slow(P, R, plane, e, mi_row, mi_col, x, y, yy, mvr, mvc);
          else ((uint2 *)pr)[yy] = mc16_row_slow(P, R, plane, e, mi_row, mi_col, x, y, yy, mvr, mvc);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < H; i++) {
      if (r == 0) out[i] = pr[i];
      else if (sizeof(Px) == 1) ((uint32_t *)out)[i] = avg_u8x4(((uint32_t *)out)[i], ((uint32_t *)pr)[i]);
      else ((uint2 *)out)[i] = make_uint2(avg_u16x2(((uint2 *)out)[i].x, ((uint2 *)pr)[i].x), avg_u16x2(((uint2 *)out)[i].y, ((uint2 *)pr)[i].y));
    }
  }
}

// unit (x, y) of `plane` belonging to 8x8 cell `cell` of the superblock: predict and store
template <typename Px, int H>
VP9B_DEV void mc_do_unit(const FrameParams &P, const Px *stage, const MiEntry *s_mi, const int (*s_woff)[2], int plane, int cell,
                         int sbr, int sbc, int x, int y) {
  typedef typename McOut<Px>::T Row;
  const int ss_x = plane ? P.ss_x : 0, ss_y = plane ? P.ss_y : 0;
  if (x >= ((P.mi_cols * 8) >> ss_x) || y >= ((P.mi_rows * 8) >> ss_y)) return;
  const MiEntry e = s_mi[cell];
  if (!e.ref0) return;
  const int ci = cell >> 3, cj = cell & 7, mi_row = sbr * 8 + ci, mi_col = sbc * 8 + cj;
  const int dstride = P.dst[plane].stride;
  Px *dst = (Px *)P.dst[plane].ptr + (size_t)y * dstride + x;
  McWin W;
  W.off[0] = W.off[1] = -1;
  W.bx = W.by = 0;
  W.S = 0;
  if (H == 8 && e.bsize < 3) {  // sub-8x8 block in an unsubsampled plane: two 4x4 units
#pragma unroll 1
    for (int half = 0; half < 2; half++) {
      Row o[4];
      mc_unit<Px, 4>(P, nullptr, W, plane, e, mi_row, mi_col, x, y + 4 * half, o);
#pragma unroll
      for (int i = 0; i < 4; i++) *(Row *)(dst + (size_t)(i + 4 * half) * dstride) = o[i];
    }
    return;
  }
  if (e.bsize >= 6) {           // 16x16 and larger: the block's staging unit that contains this cell
    const int S8 = min((int)d_bw8[e.bsize], (int)d_bh8[e.bsize]);
    const int ci0 = ci & ~(S8 - 1), cj0 = cj & ~(S8 - 1);
    W.off[0] = s_woff[ci0 * 8 + cj0][0];
    W.off[1] = s_woff[ci0 * 8 + cj0][1];
    W.bx = (sbc * 8 + cj0) * 8;
    W.by = (sbr * 8 + ci0) * 8;
    W.S = S8 * 8;
  }
  Row o[H];
  mc_unit<Px, H>(P, stage, W, plane, e, mi_row, mi_col, x, y, o);
#pragma unroll
  for (int i = 0; i < H; i++) *(Row *)(dst + (size_t)i * dstride) = o[i];
}

template <typename Px, typename FB>
__global__ void __launch_bounds__(128, sizeof(Px) == 1 ? VP9B_MC_MINB : 3) k_inter_pred_sb(const __grid_constant__ FB B) {
  const FrameParams &P = B.f[blockIdx.z];
  if (!P.do_inter) return;
  const int sbc = blockIdx.x, sbr = blockIdx.y, tid = threadIdx.x;
  if (sbc >= P.sb_cols || sbr >= P.sb_rows) return;   // the grid covers the largest frame of the batch
  extern __shared__ __align__(128) unsigned char mc_smem[];
  Px *const stage = (Px *)mc_smem;
  __shared__ MiEntry s_mi[64];
  __shared__ int s_woff[64][2];
  __shared__ int s_alloc;
  __shared__ __align__(8) uint64_t s_bar;
  // ---- the 64 cells of the superblock
  if (tid < 64) {
    const int mi_row = sbr * 8 + (tid >> 3), mi_col = sbc * 8 + (tid & 7);
    MiEntry e;
    e.mv[0][0] = e.mv[0][1] = e.mv[1][0] = e.mv[1][1] = 0;
    e.ref0 = e.ref1 = e.filter = e.bsize = 0;
    e.aux = 0;
    if (mi_row < P.mi_rows && mi_col < P.mi_cols) e = P.mi[(size_t)mi_row * P.mi_cols + mi_col];
    s_mi[tid] = e;
    s_woff[tid][0] = s_woff[tid][1] = -1;
  }
  if (tid == 0) {
    s_alloc = 0;
    mbar_init(&s_bar, 64);
  }
  __syncthreads();
  // ---- staging: the thread of a staging unit's first cell issues its copies
  if (tid < 64) {
    const MiEntry e = s_mi[tid];
    uint32_t bytes = 0;
    const int ci = tid >> 3, cj = tid & 7;
    if (P.tmaps && e.ref0 && e.bsize >= 6) {
      const int S8 = min((int)d_bw8[e.bsize], (int)d_bh8[e.bsize]), S = S8 * 8, Sc = S >> P.ss_x;
      if (!(ci & (S8 - 1)) && !(cj & (S8 - 1))) {
        const int bx = (sbc * 8 + cj) * 8, by = (sbr * 8 + ci) * 8;
        const int need = McStage<Px>::region(S) + 2 * McStage<Px>::region(Sc);
#pragma unroll 1
        for (int r = 0; r < (e.ref1 ? 2 : 1); r++) {
          const RefGeom &R = P.ref[(r ? e.ref1 : e.ref0) - 1];
          if (R.x_step_q4 != 16 || R.y_step_q4 != 16) continue;
          // window origins (first tap) in luma and chroma; every tap must lie inside the reference
          const int wx = bx + ((e.mv[r][1] * 2) >> 4) - 3, wy = by + ((e.mv[r][0] * 2) >> 4) - 3;
          const int cmx = e.mv[r][1] * (1 << (1 - P.ss_x)), cmy = e.mv[r][0] * (1 << (1 - P.ss_y));
          const int wxc = (bx >> P.ss_x) + (cmx >> 4) - 3, wyc = (by >> P.ss_y) + (cmy >> 4) - 3;
          if (wx < 0 || wy < 0 || wx + S + 7 > R.w[0] || wy + S + 7 > R.h[0] || wxc < 0 || wyc < 0 || wxc + Sc + 7 > R.w[1] ||
              wyc + Sc + 7 > R.h[1])
            continue;
          const int off = atomicAdd(&s_alloc, need);
          if (off + need > McStage<Px>::kBudget) continue;      // buffer full: this unit reads global memory
          s_woff[tid][r] = off;
          const CUtensorMap *maps = P.tmaps + (size_t)R.slot * 12;
          constexpr int am = ~(McStage<Px>::kAlign - 1);
          tma_load_2d(stage + off, maps + McStage<Px>::cls(S), wx & am, wy, &s_bar);
          tma_load_2d(stage + off + McStage<Px>::region(S), maps + 4 + McStage<Px>::cls(Sc), wxc & am, wyc, &s_bar);
          tma_load_2d(stage + off + McStage<Px>::region(S) + McStage<Px>::region(Sc), maps + 8 + McStage<Px>::cls(Sc), wxc & am, wyc, &s_bar);
          bytes += (uint32_t)sizeof(Px) * (McStage<Px>::boxw(S) * (S + 7) + 2 * McStage<Px>::boxw(Sc) * (Sc + 7));
        }
      }
    }
    // expect_tx may follow the copies: the phase cannot complete before this thread has arrived
    if (bytes) mbar_arrive_expect_tx(&s_bar, bytes);
    else mbar_arrive(&s_bar);
  }
  mbar_wait(&s_bar, 0);     // every staged window has landed; s_woff is visible (arrive = release, wait = acquire)
  // ---- luma: 16 x 8 units of 4 x 8
  mc_do_unit<Px, 8>(P, stage, s_mi, s_woff, 0, ((tid >> 4) << 3) | ((tid & 15) >> 1), sbr, sbc, sbc * 64 + (tid & 15) * 4, sbr * 64 + (tid >> 4) * 8);
  // ---- chroma: units of 4 x 4 when the plane is vertically subsampled, else 4 x 8 (a unit never
  // straddles two 8x8 luma cells).  Horizontally subsampled planes are 32 samples wide per superblock:
  // 64 units per plane, the plane comes from the thread index; full-width planes have 128 units and
  // a thread takes one of each plane.
  const int per_thread = P.ss_x ? 1 : 2;
  if (P.ss_y) {     // 4:2:0, 4:4:0
#pragma unroll 1
    for (int it = 0; it < per_thread; it++) {
      const int pl = P.ss_x ? 1 + (tid >> 6) : 1 + it, i = P.ss_x ? (tid & 63) : tid;
      const int ux = P.ss_x ? (i & 7) : (i & 15), uy = P.ss_x ? (i >> 3) : (i >> 4);     // unit column / row in the superblock
      mc_do_unit<Px, 4>(P, stage, s_mi, s_woff, pl, uy * 8 + (P.ss_x ? ux : (ux >> 1)), sbr, sbc, sbc * (64 >> P.ss_x) + ux * 4, sbr * 32 + uy * 4);
    }
  } else {          // 4:4:4, 4:2:2
#pragma unroll 1
    for (int it = 0; it < per_thread; it++) {
      const int pl = P.ss_x ? 1 + (tid >> 6) : 1 + it, i = P.ss_x ? (tid & 63) : tid;
      const int ux = P.ss_x ? (i & 7) : (i & 15), uy = P.ss_x ? (i >> 3) : (i >> 4);
      mc_do_unit<Px, 8>(P, stage, s_mi, s_woff, pl, uy * 8 + (P.ss_x ? ux : (ux >> 1)), sbr, sbc, sbc * (64 >> P.ss_x) + ux * 4, sbr * 64 + uy * 8);
    }
  }
}

// ------------------------------------------------------------------------------------------
// stage 1, flat variant (no staging): thread = one 4 x H unit anywhere in the plane, warp = 8 x 4
// units, reference rows straight from global memory through L1.  All three planes in one launch:
// luma tiles (32 x 128 samples) first, then the two chroma planes.  Kept next to k_inter_pred_sb as
// the A/B baseline (VP9B_MC_FLAT=1).
template <typename Px, int H>
VP9B_DEV void inter_pred_flat_body(const FrameParams &P, int plane, int bx_, int by_) {
  typedef typename McOut<Px>::T Row;
  const int ss_x = plane ? P.ss_x : 0, ss_y = plane ? P.ss_y : 0;
  const int pw = (P.mi_cols * 8) >> ss_x, ph = (P.mi_rows * 8) >> ss_y;
  const int x = (bx_ * 8 + (threadIdx.x & 7)) * 4;
  const int y = (by_ * 16 + (threadIdx.x >> 3)) * H;
  if (x >= pw || y >= ph) return;
  const int dstride = P.dst[plane].stride;
  Px *dst = (Px *)P.dst[plane].ptr + (size_t)y * dstride + x;
  const int mi_col = (x << ss_x) >> 3, mi_row = (y << ss_y) >> 3;
  const MiEntry e = P.mi[(size_t)mi_row * P.mi_cols + mi_col];
  if (!e.ref0) return;
  McWin W;
  W.off[0] = W.off[1] = -1;
  W.bx = W.by = 0;
  W.S = 0;
  if (H == 8 && e.bsize < 3) {  // sub-8x8 block in an unsubsampled plane: two 4x4 units
#pragma unroll 1
    for (int half = 0; half < 2; half++) {
      Row o[4];
      mc_unit<Px, 4>(P, nullptr, W, plane, e, mi_row, mi_col, x, y + 4 * half, o);
#pragma unroll
      for (int i = 0; i < 4; i++) *(Row *)(dst + (size_t)(i + 4 * half) * dstride) = o[i];
    }
    return;
  }
  Row o[H];
  mc_unit<Px, H>(P, nullptr, W, plane, e, mi_row, mi_col, x, y, o);
#pragma unroll
  for (int i = 0; i < H; i++) *(Row *)(dst + (size_t)i * dstride) = o[i];
}

template <typename Px, typename FB>
__global__ void __launch_bounds__(128, sizeof(Px) == 1 ? 8 : 4) k_inter_pred_flat(const __grid_constant__ FB B) {
  const FrameParams &P = B.f[blockIdx.z];
  if (!P.do_inter) return;
  const int pwy = P.mi_cols * 8, phy = P.mi_rows * 8, pwc = pwy >> P.ss_x, phc = phy >> P.ss_y;
  const int gy_y = (phy + 127) / 128, gx_y = (pwy + 31) / 32, gx_c = (pwc + 31) / 32, gy_c = P.ss_x ? (phc + 63) / 64 : (phc + 127) / 128;
  int by = blockIdx.y, plane = 0;
  if (by >= gy_y) {
    by -= gy_y;
    plane = 1;
    if (by >= gy_c) { by -= gy_c; plane = 2; }
    if (by >= gy_c || (int)blockIdx.x >= gx_c) return;
  } else if ((int)blockIdx.x >= gx_y) {
    return;
  }
  if (plane && P.ss_x) inter_pred_flat_body<Px, 4>(P, plane, blockIdx.x, by);
  else inter_pred_flat_body<Px, 8>(P, plane, blockIdx.x, by);
}

// ------------------------------------------------------------------------------------------
// stages 2/3: inverse transform + add, one warp per transform block.  Rows then columns
// (vpx_idctNxN_add_c), final shift 4/5/6/6, clip to the pixel range.  tbuf: per-warp shared
// scratch of N*(N+1) ints.
// Four consecutive samples of a row: residuals added to the prediction in place (one 32-/64-bit
// read-modify-write), or stored as a residual plane for the intra wavefront (which then only
// predicts and adds).
template <typename Px> struct Quad;                         // four samples of a row as one load
template <> struct Quad<uint8_t> { typedef uint32_t T; };
template <> struct Quad<uint16_t> { typedef uint2 T; };
VP9B_DEV void add4_clip(uint8_t *p, uint32_t w, int v0, int v1, int v2, int v3, int) {
  const int r0 = clampi((int)(w & 0xff) + v0, 0, 255), r1 = clampi((int)((w >> 8) & 0xff) + v1, 0, 255);
  const int r2 = clampi((int)((w >> 16) & 0xff) + v2, 0, 255), r3 = clampi((int)(w >> 24) + v3, 0, 255);
  *(uint32_t *)p = (uint32_t)r0 | ((uint32_t)r1 << 8) | ((uint32_t)r2 << 16) | ((uint32_t)r3 << 24);
}
VP9B_DEV void add4_clip(uint16_t *p, uint2 w, int v0, int v1, int v2, int v3, int maxv) {
  const int r0 = clampi((int)(w.x & 0xffff) + v0, 0, maxv), r1 = clampi((int)(w.x >> 16) + v1, 0, maxv);
  const int r2 = clampi((int)(w.y & 0xffff) + v2, 0, maxv), r3 = clampi((int)(w.y >> 16) + v3, 0, maxv);
  *(uint2 *)p = make_uint2((uint32_t)r0 | ((uint32_t)r1 << 16), (uint32_t)r2 | ((uint32_t)r3 << 16));
}
VP9B_DEV void put4_raw(int16_t *p, int v0, int v1, int v2, int v3) {
  *(uint2 *)p = make_uint2(((uint32_t)v0 & 0xffff) | ((uint32_t)v1 << 16), ((uint32_t)v2 & 0xffff) | ((uint32_t)v3 << 16));
}
VP9B_DEV void put4_raw(int32_t *p, int v0, int v1, int v2, int v3) { *(int4 *)p = make_int4(v0, v1, v2, v3); }
template <typename Px> struct ResidT { typedef int16_t T; };
template <> struct ResidT<uint16_t> { typedef int32_t T; };

// Rows then columns of the blocks a warp holds in shared memory (tb = this lane's block, N + 1 ints
// per row, lane l = row / column l), in place: afterwards tb[r * S + c] is the final residual
// (vpx_idctNxN_add_c: final shift 4/5/6/6).  xf = this lane takes part; lead = warp-uniform eob class.
template <typename PT, int N>
VP9B_DEV void itx_2d_inplace(int *tb, int l, bool xf, int lead, bool row_adst, bool col_adst) {
  constexpr int S = N + 1;
  constexpr int shift = (N == 4) ? 4 : (N == 8 ? 5 : 6);
#pragma unroll 1
  for (int pass = 0; pass < 2; pass++) {   // one copy of the butterflies for both passes
    if (xf && (pass || l < lead)) {        // rows lead.. are zero and stay zero
      int out[N];
      itx_line_lead<PT, N>(tb + l * (pass ? 1 : S), pass ? S : 1, pass ? col_adst : row_adst, lead, out);
      if (pass == 0) {
#pragma unroll
        for (int j = 0; j < N; j++) tb[l * S + j] = out[j];
      } else {
#pragma unroll
        for (int j = 0; j < N; j++) tb[j * S + l] = (out[j] + (1 << (shift - 1))) >> shift;
      }
    }
    __syncwarp();
  }
}
// vpx_iwht4x4_16_add_c (inv_txfm.c:18-66), same convention
VP9B_DEV void iwht_2d_inplace(int *tb, int l, bool xf) {
  constexpr int S = 5;
  int o[4];
  if (xf) {
    iwht4(tb[l * S + 0] >> 2, tb[l * S + 1] >> 2, tb[l * S + 2] >> 2, tb[l * S + 3] >> 2, o);
    for (int j = 0; j < 4; j++) tb[l * S + j] = o[j];
  }
  __syncwarp();
  if (xf) {
    iwht4(tb[0 * S + l], tb[1 * S + l], tb[2 * S + l], tb[3 * S + l], o);
    for (int j = 0; j < 4; j++) tb[j * S + l] = o[j];
  }
  __syncwarp();
}
// Coefficients of a block -> their positions in the shared-memory block (scan order in the packed
// stream).  Two phases so that the global loads are in flight while the block is being cleared:
// coef_preload() issues the first two rounds of loads, coef_scatter() stores them and walks the rest.
struct CoefPre { int pos[2], val[2]; };
VP9B_DEV int coef_at(const void *coef_base, int coef_bytes, uint32_t i) {
  return coef_bytes == 2 ? (int)((const int16_t *)coef_base)[i] : ((const int32_t *)coef_base)[i];
}
VP9B_DEV CoefPre coef_preload(const void *coef_base, int coef_bytes, uint32_t coef_off, int eob, const int16_t *scan, int l, int step,
                              bool on) {
  CoefPre c;
#pragma unroll
  for (int k = 0; k < 2; k++) {
    const int i = l + k * step;
    c.pos[k] = -1;
    c.val[k] = 0;
    if (on && i < eob) {
      c.pos[k] = scan[i];
      c.val[k] = coef_at(coef_base, coef_bytes, coef_off + i);
    }
  }
  return c;
}
template <int N>
VP9B_DEV void coef_scatter(int *tb, const CoefPre &c, const void *coef_base, int coef_bytes, uint32_t coef_off, int eob,
                           const int16_t *scan, int l, int step) {
  constexpr int S = N + 1;
#pragma unroll
  for (int k = 0; k < 2; k++)
    if (c.pos[k] >= 0) tb[(c.pos[k] / N) * S + (c.pos[k] % N)] = c.val[k];
  for (int i = l + 2 * step; i < eob; i += step) {
    const int pos = scan[i];
    tb[(pos / N) * S + (pos % N)] = coef_at(coef_base, coef_bytes, coef_off + i);
  }
}

// inverse transform of one intra block into the residual plane (one warp per block); rows of the
// plane are written as 4-sample vectors
template <typename PT, int N, typename R>
VP9B_DEV void warp_itx(int *tbuf, const void *coef_base, int coef_bytes, const vp9b_txb &t, R *dst, int stride,
                       int lossless, int lane) {
  constexpr int S = N + 1, Q = N / 4;     // quads per row
  const int tx_size = t.tx & 3, tx_type = lossless ? 0 : ((t.tx >> 2) & 3);
  const int eob = t.eob;
  if (eob == 1 && tx_type == 0 && !lossless) {  // DC only: vpx_idctNxN_1_add_c
    constexpr int sh = (N == 4) ? 4 : (N == 8 ? 5 : 6);
    const int dc = coef_bytes == 2 ? (int)((const int16_t *)coef_base)[t.coef_off] : ((const int32_t *)coef_base)[t.coef_off];
    const int v1 = rs14<PT>((PT)dc * (PT)c16);
    const int a1 = (rs14<PT>((PT)v1 * (PT)c16) + (1 << (sh - 1))) >> sh;
    for (int q = lane; q < N * Q; q += 32) put4_raw(dst + (size_t)(q / Q) * stride + (q % Q) * 4, a1, a1, a1, a1);
    return;
  }
  const int16_t *scan = d_scan + scan_offset(tx_size, tx_type);
  const CoefPre cp = coef_preload(coef_base, coef_bytes, t.coef_off, eob, scan, lane, 32, true);
  for (int i = lane; i < N * S; i += 32) tbuf[i] = 0;
  __syncwarp();
  coef_scatter<N>(tbuf, cp, coef_base, coef_bytes, t.coef_off, eob, scan, lane, 32);
  __syncwarp();
  if (N == 4 && lossless) {
    iwht_2d_inplace(tbuf, lane, lane < 4);
  } else {
    const bool row_adst = (N < 32) && (tx_type & 2), col_adst = (N < 32) && (tx_type & 1);
    itx_2d_inplace<PT, N>(tbuf, lane, lane < N, eob_lead<N>(tx_type, eob), row_adst, col_adst);
  }
  for (int q = lane; q < N * Q; q += 32) {
    const int r = q / Q, c4 = (q % Q) * 4;
    const int *tv = tbuf + r * S + c4;
    put4_raw(dst + (size_t)r * stride + c4, tv[0], tv[1], tv[2], tv[3]);
  }
  __syncwarp();
}

template <typename Px, typename PT>
VP9B_DEV void warp_intra_residual(const FrameParams &P, int *tbuf, const vp9b_txb &t, int plane, int lane) {
  typedef typename ResidT<Px>::T R;
  const int stride = P.dst[plane].stride;
  R *dst = (R *)P.resid[plane] + (size_t)t.y4 * 4 * stride + t.x4 * 4;
  const int cb = sizeof(Px) == 1 ? 2 : 4;
  switch (t.tx & 3) {
    case 0: warp_itx<PT, 4>(tbuf, P.coef, cb, t, dst, stride, P.lossless, lane); break;
    case 1: warp_itx<PT, 8>(tbuf, P.coef, cb, t, dst, stride, 0, lane); break;
    case 2: warp_itx<PT, 16>(tbuf, P.coef, cb, t, dst, stride, 0, lane); break;
    default: warp_itx<PT, 32>(tbuf, P.coef, cb, t, dst, stride, 0, lane); break;
  }
}

// Inter residuals: a warp takes 32 / N blocks of size N (lane = (block, row / column)), transforms
// them in shared memory and adds them to the prediction row-wise, four samples per access.
template <typename Px, typename PT, int N>
VP9B_DEV void group_itx_add(int *tbuf, const FrameParams &P, int first, int count, int lane) {
  constexpr int S = N + 1, Q = N / 4, QB = N * Q, NSUB = 32 / N;
  const int sub = lane / N, l = lane % N;
  int *tb = tbuf + sub * (N * S);
  const bool active = sub < count;
  vp9b_txb t;
  t.coef_off = 0; t.x4 = t.y4 = 0; t.eob = 0; t.tx = 0; t.aux = 0;
  if (active) t = P.txb_inter[first + sub];
  const int plane = (t.tx >> 4) & 3;
  const int stride = P.dst[plane].stride;
  Px *dst = (Px *)P.dst[plane].ptr + (size_t)t.y4 * 4 * stride + t.x4 * 4;
  const int lossless = (N == 4) ? P.lossless : 0;
  const int tx_type = lossless ? 0 : ((t.tx >> 2) & 3);
  const int eob = t.eob;
  constexpr int shift = (N == 4) ? 4 : (N == 8 ? 5 : 6);
  const int maxv = (1 << P.bd) - 1;
  // DC-only DCT blocks (more than half of all inter blocks): both passes reduce to one value
  // (vpx_idctNxN_1_add_c, inv_txfm.c:159-173 and the 8/16/32 twins)
  const bool dc_only = active && eob == 1 && tx_type == 0 && !lossless;
  int fill = 0;
  if (dc_only) {
    const int dc = sizeof(Px) == 1 ? (int)((const int16_t *)P.coef)[t.coef_off] : ((const int32_t *)P.coef)[t.coef_off];
    const int v1 = rs14<PT>((PT)dc * (PT)c16);
    fill = (rs14<PT>((PT)v1 * (PT)c16) + (1 << (shift - 1))) >> shift;
  }
  const bool xf = active && !dc_only;
  // the prediction under the blocks: quad q of the warp = (block, row, 4 columns), the block's
  // destination comes from its first lane.  Loaded now, used after the transforms (one L2 / HBM
  // round trip hidden behind them).
  typedef typename Quad<Px>::T QT;
  constexpr int NQ = NSUB * QB / 32;
  QT pq[NQ];
#pragma unroll
  for (int k = 0; k < NQ; k++) {
    const int q = lane + 32 * k, sb = q / QB, r = (q % QB) / Q, c4 = (q % Q) * 4;
    const unsigned long long dp = __shfl_sync(0xffffffffu, active ? (unsigned long long)dst : 0ull, sb * N);
    const int st = __shfl_sync(0xffffffffu, stride, sb * N);
    if (dp) pq[k] = *(const QT *)((Px *)dp + (size_t)r * st + c4);
  }
  const int16_t *scan = d_scan + scan_offset(t.tx & 3, tx_type);
  const CoefPre cp = coef_preload(P.coef, sizeof(Px) == 1 ? 2 : 4, t.coef_off, eob, scan, l, N, xf);
  if (active)
    for (int i = l; i < N * S; i += N) tb[i] = fill;
  __syncwarp();
  if (xf) coef_scatter<N>(tb, cp, P.coef, sizeof(Px) == 1 ? 2 : 4, t.coef_off, eob, scan, l, N);
  __syncwarp();
  if (N == 4 && lossless) {
    iwht_2d_inplace(tb, l, xf);
  } else {
    // eob class: the blocks of a warp take the widest one among them so that the warp stays converged
    const int lead = __reduce_max_sync(0xffffffffu, xf ? eob_lead<N>(tx_type, eob) : 0);
    if (lead) itx_2d_inplace<PT, N>(tb, l, xf, lead, (N < 32) && (tx_type & 2), (N < 32) && (tx_type & 1));
  }
  // add the residuals to the prediction loaded above
#pragma unroll
  for (int k = 0; k < NQ; k++) {
    const int q = lane + 32 * k, sb = q / QB, r = (q % QB) / Q, c4 = (q % Q) * 4;
    const unsigned long long dp = __shfl_sync(0xffffffffu, active ? (unsigned long long)dst : 0ull, sb * N);
    const int st = __shfl_sync(0xffffffffu, stride, sb * N);
    if (dp) {
      const int *tv = tbuf + sb * (N * S) + r * S + c4;
      add4_clip((Px *)dp + (size_t)r * st + c4, pq[k], tv[0], tv[1], tv[2], tv[3], maxv);
    }
  }
  __syncwarp();
}

constexpr int kResWarps = 4;
template <typename Px, typename PT, typename FB>
__global__ void __launch_bounds__(kResWarps * 32, sizeof(Px) == 1 ? 8 : 6) k_residual(const __grid_constant__ FB B) {
  const FrameParams &P = B.f[blockIdx.z];
  if (!P.do_res) return;
  __shared__ int tbuf[kResWarps][32 * 33];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nwarps = gridDim.x * kResWarps;
  // warp tasks: ceil(n/8) groups of 4x4, ceil(n/4) of 8x8, ceil(n/2) of 16x16, n of 32x32
  const int n4 = P.n_txb_size[0], n8 = P.n_txb_size[1], n16 = P.n_txb_size[2], n32 = P.n_txb_size[3];
  const int t4 = (n4 + 7) >> 3, t8 = (n8 + 3) >> 2, t16 = (n16 + 1) >> 1;
  const int total = t4 + t8 + t16 + n32;
  // big transforms first: they are the long tasks
  for (int i = blockIdx.x * kResWarps + warp; i < total; i += nwarps) {
    int k = i;
    if (k < n32) { group_itx_add<Px, PT, 32>(tbuf[warp], P, n4 + n8 + n16 + k, 1, lane); continue; }
    k -= n32;
    if (k < t16) { group_itx_add<Px, PT, 16>(tbuf[warp], P, n4 + n8 + 2 * k, min(2, n16 - 2 * k), lane); continue; }
    k -= t16;
    if (k < t8) { group_itx_add<Px, PT, 8>(tbuf[warp], P, n4 + 4 * k, min(4, n8 - 4 * k), lane); continue; }
    k -= t8;
    group_itx_add<Px, PT, 4>(tbuf[warp], P, 8 * k, min(8, n4 - 8 * k), lane);
  }
}

// Residuals of the intra blocks: every block's inverse transform is independent of the prediction,
// so it is taken off the intra wavefront's dependency chain and computed here, one warp per block,
// into a residual plane; the wavefront (k_intra) then only predicts and adds.
template <typename Px, typename PT, typename FB>
__global__ void __launch_bounds__(kResWarps * 32, VP9B_IRES_MINB) k_intra_residual(const __grid_constant__ FB B) {
  const FrameParams &P = B.f[blockIdx.z];
  if (!P.do_intra) return;
  __shared__ int tbuf[kResWarps][32 * 33];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nwarps = gridDim.x * kResWarps;
  const int n0 = P.n_txb_intra[0], n1 = P.n_txb_intra[1], total = n0 + n1 + P.n_txb_intra[2];
  for (int i = blockIdx.x * kResWarps + warp; i < total; i += nwarps) {
    int plane = 0, k = i;
    if (k >= n0) { k -= n0; plane = 1; if (k >= n1) { k -= n1; plane = 2; } }
    const vp9b_txb t = P.txb_intra[plane][k];
    if (t.eob) warp_intra_residual<Px, PT>(P, tbuf[warp], t, plane, lane);
  }
}

// ------------------------------------------------------------------------------------------
// stage 3: intra prediction.  Edge rules of build_intra_predictors (vp9_reconintra.c:113-402):
// unavailable left -> base+1, unavailable above -> base-1, above-right only read for 4x4 blocks
// with have_right, everything beyond the aligned frame size replicates the last sample.
// Predictors: vpx_dsp/intrapred.c (closed-form per sample).
template <typename Px> struct LfVec;       // vector type moving 8 samples
template <> struct LfVec<uint8_t> { typedef uint2 T; };
template <> struct LfVec<uint16_t> { typedef uint4 T; };
// shared-memory rows are only 4-byte aligned (odd word stride): move vectors as 32-bit words
VP9B_DEV void smem_put(void *d, uint2 v) { uint32_t *w = (uint32_t *)d; w[0] = v.x; w[1] = v.y; }
VP9B_DEV void smem_put(void *d, uint4 v) { uint32_t *w = (uint32_t *)d; w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w; }
VP9B_DEV void smem_get(const void *s, uint2 &v) { const uint32_t *w = (const uint32_t *)s; v.x = w[0]; v.y = w[1]; }
VP9B_DEV void smem_get(const void *s, uint4 &v) { const uint32_t *w = (const uint32_t *)s; v.x = w[0]; v.y = w[1]; v.z = w[2]; v.w = w[3]; }

#define AVG2(a, b) (((a) + (b) + 1) >> 1)
#define AVG3(a, b, c) (((a) + 2 * (b) + (c) + 2) >> 2)

VP9B_DEV uint32_t ld_acquire(const uint32_t *p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
VP9B_DEV void st_release(uint32_t *p, uint32_t v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
template <typename Px>
VP9B_DEV int ld_px(const Px *p) {  // L2-coherent load (neighbour samples are written by other SMs)
  return (int)__ldcg(p);
}

// edge buffer per warp: e[0..63] = above[0..63], e[64] unused pad; left l[0..31]; tl
struct IntraEdge {
  int above[64];
  int left[32];
  int tl;
};

template <typename Px>
VP9B_DEV void warp_intra_predict(const FrameParams &P, IntraEdge &E, const vp9b_txb &t, Px *dst, int stride, int lane) {
  const int tx_size = t.tx & 3, plane = (t.tx >> 4) & 3;
  const int mode = t.aux & 15, have_top = (t.aux >> 4) & 1, have_left = (t.aux >> 5) & 1, have_right = (t.aux >> 6) & 1;
  const int bs = 4 << tx_size, bd = P.bd;
  const int ss_x = plane ? P.ss_x : 0, ss_y = plane ? P.ss_y : 0;
  const int frame_w = (P.mi_cols * 8) >> ss_x, frame_h = (P.mi_rows * 8) >> ss_y;
  const int x0 = t.x4 * 4, y0 = t.y4 * 4;
  const int base = 128 << (bd - 8);
  // --- edges
  if (lane < bs) {
    int v = base + 1;
    if (have_left) {
      const int yy = (y0 + lane < frame_h) ? lane : frame_h - y0 - 1;
      v = dst[yy * stride - 1];
    }
    E.left[lane] = v;
  }
  for (int i = lane; i < 2 * bs; i += 32) {
    int v = base - 1;
    if (have_top) {
      const bool real = (i < bs) || (have_right && bs == 4);
      int xx = real ? i : bs - 1;
      if (x0 + xx >= frame_w) xx = frame_w - x0 - 1;
      v = dst[xx - stride];
    }
    E.above[i] = v;
  }
  if (lane == 0) E.tl = have_top ? (have_left ? (int)dst[-stride - 1] : base + 1) : base - 1;
  __syncwarp();
  const int *a = E.above, *l = E.left;
  const int tl = E.tl;
  const int maxv = (1 << bd) - 1;
  int dc = 0;
  if (mode == 0) {
    // dc_pred[left][top]: warp reduction of the available edges
    int s = 0, cnt = 0;
    if (have_top) { for (int i = lane; i < bs; i += 32) s += a[i]; cnt += bs; }
    if (have_left) { for (int i = lane; i < bs; i += 32) s += l[i]; cnt += bs; }
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    // cnt is bs or 2 * bs: the division of dc_predictor (intrapred.c) is a shift
    dc = cnt ? (s + (cnt >> 1)) >> (2 + tx_size + (have_top & have_left)) : base;
  }
#define AT(i) ((i) < 0 ? tl : a[i])   /* above with above[-1] = top-left */
#define LT(i) ((i) < 0 ? tl : l[i])   /* left with left[-1] = top-left  */
  const int lgb = 2 + tx_size, npx = bs * bs;
  // one sample loop per mode: the dispatch is outside the loop, row / column come from shifts
#define INTRA_LOOP(...)                                    \
  for (int p = lane; p < npx; p += 32) {                   \
    const int r = p >> lgb, c = p & (bs - 1);              \
    int v;                                                 \
    __VA_ARGS__                                            \
    dst[r * stride + c] = (Px)v;                           \
  }
  switch (mode) {
    case 0: INTRA_LOOP(v = dc;) break;
    case 1: INTRA_LOOP(v = a[c];) break;
    case 2: INTRA_LOOP(v = l[r];) break;
    case 9: INTRA_LOOP(v = clampi(l[r] + a[c] - tl, 0, maxv);) break;
    case 3:  // D45
      INTRA_LOOP({
        const int i = r + c;
        if (bs == 4) v = (i == 6) ? a[7] : AVG3(a[i], a[i + 1], a[i + 2]);
        else v = (i < bs - 1) ? AVG3(a[i], a[i + 1], a[i + 2]) : a[bs - 1];
      })
      break;
    case 8:  // D63
      INTRA_LOOP({
        const int k = r >> 1, i = k + c;
        if (bs != 4 && r >= 2 && c >= bs - 1 - k) v = a[bs - 1];
        else v = (r & 1) ? AVG3(a[i], a[i + 1], a[i + 2]) : AVG2(a[i], a[i + 1]);
      })
      break;
    case 7:  // D207
      INTRA_LOOP({
        const int j = r + (c >> 1);
        if (j >= bs - 1) v = l[bs - 1];
        else if (!(c & 1)) v = AVG2(l[j], l[j + 1]);
        else v = AVG3(l[j], l[j + 1], l[min(j + 2, bs - 1)]);
      })
      break;
    case 5:  // D117
      INTRA_LOOP({
        const int k = min(r >> 1, c), rr = r - 2 * k, cc = c - k;
        if (rr == 0) v = AVG2(AT(cc - 1), a[cc]);
        else if (rr == 1) v = cc == 0 ? AVG3(l[0], tl, a[0]) : AVG3(AT(cc - 2), a[cc - 1], a[cc]);
        else v = AVG3(LT(rr - 3), l[rr - 2], l[rr - 1]);
      })
      break;
    case 4:  // D135
      INTRA_LOOP({
        const int d = c - r;
        if (d == 0) v = AVG3(l[0], tl, a[0]);
        else if (d > 0) v = AVG3(AT(d - 2), a[d - 1], a[d]);
        else v = AVG3(LT(-d - 2), l[-d - 1], l[-d]);
      })
      break;
    default:  // 6: D153
      INTRA_LOOP({
        const int k = min(r, c >> 1), rr = r - k, cc = c - 2 * k;
        if (cc >= 2) v = AVG3(AT(cc - 3), a[cc - 2], a[cc - 1]);
        else if (cc == 0) v = AVG2(LT(rr - 1), l[rr]);
        else if (rr == 0) v = AVG3(l[0], tl, a[0]);
        else v = AVG3(LT(rr - 2), l[rr - 1], l[rr]);
      })
      break;
  }
#undef INTRA_LOOP
#undef AT
#undef LT
  __syncwarp();
}

// residual of one block: 2^LG x 2^LG samples spread over the warp, lane + 32k -> (row, column)
template <typename R, int LG>
VP9B_DEV void resid_fetch(const R *rp, int stride, int lane, R *rv) {
  constexpr int bs = 1 << LG, npx = bs * bs, it = (npx + 31) / 32;
#pragma unroll
  for (int k = 0; k < it; k++) {
    const int p = lane + 32 * k;
    if (npx >= 32 || p < npx) rv[k] = __ldg(rp + (size_t)(p >> LG) * stride + (p & (bs - 1)));
  }
}
template <typename Px, typename R, int LG>
VP9B_DEV void resid_add(Px *d, int ts, int lane, const R *rv, int maxv) {
  constexpr int bs = 1 << LG, npx = bs * bs, it = (npx + 31) / 32;
#pragma unroll
  for (int k = 0; k < it; k++) {
    const int p = lane + 32 * k;
    if (npx >= 32 || p < npx) {
      Px *q = d + (p >> LG) * ts + (p & (bs - 1));
      *q = (Px)clampi((int)*q + (int)rv[k], 0, maxv);
    }
  }
}

constexpr int kWaveWarps = 4;

// progress counters hold prog_base + count; stale values of earlier frames compare as negative
VP9B_DEV uint32_t ld_relaxed(const uint32_t *p) {
  uint32_t v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
VP9B_DEV void wait_progress(const uint32_t *p, uint32_t base, int need) {
  // usually satisfied on arrival: one acquire load (a single L2 round trip) decides; otherwise poll
  // with relaxed loads (no L1 invalidation per poll), then one acquire to order the data reads
  if ((int)(ld_acquire(p) - base) >= need) return;
  while ((int)(ld_relaxed(p) - base) < need) __nanosleep(20);
  (void)ld_acquire(p);
}

// One persistent warp per (superblock row, plane) walks its row left to right.  Intra edges
// only reach into the left, above-left and above superblocks (above-right samples are taken
// inside the prediction block only, vp9_reconintra.c:356-386,412), so row r may process
// superblock c once row r-1 has finished superblock c: a wavefront with lag 1, tracked by one
// progress counter per (plane, row).  The superblock's samples (plus the row above and the
// column to the left) are staged in shared memory, every transform block of the superblock is
// predicted and reconstructed there in decode order, and the tile is written back once.
template <typename Px> struct IntraTileGeom {
  static constexpr int kMargin = 16 / sizeof(Px);            // left margin: interior rows 16-byte aligned
  static constexpr int kTS = 64 + kMargin + 16 / sizeof(Px); // row stride in samples (16-byte multiple)
  static constexpr int kRows = 65;
  static constexpr size_t kTileBytes = (size_t)kRows * kTS * sizeof(Px);
  static constexpr size_t kWarpBytes = (kTileBytes + 32 * 33 * sizeof(int) + sizeof(IntraEdge) + 15) & ~(size_t)15;
};

// MINB = CTAs per SM the kernel is compiled for: 1 (255 registers, fastest single warp) or 4 (128
// registers, no spills): the second variant is launched when a batch brings more CTAs than fit at
// 255 registers (2 per SM), i.e. for large batches of large 8-bit frames (4K, 32 per launch: +4 %).
template <typename Px, typename PT, typename FB, int MINB>
__global__ void __launch_bounds__(kWaveWarps * 32, MINB) k_intra(const __grid_constant__ FB B) {
  const FrameParams &P = B.f[blockIdx.z];
  if (!P.do_intra) return;
  extern __shared__ __align__(16) unsigned char intra_smem[];
  typedef IntraTileGeom<Px> G;
  typedef typename LfVec<Px>::T Vec;  // 8 samples
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned char *wbase = intra_smem + (size_t)warp * G::kWarpBytes;
  Px *tile = (Px *)wbase;
  int *tbuf = (int *)(wbase + G::kTileBytes);
  IntraEdge &edge = *(IntraEdge *)(wbase + G::kTileBytes + 32 * 33 * sizeof(int));
  // one ticket per CTA: its warps take four consecutive units, as with a blockIdx mapping
  __shared__ int s_cta;
  if (threadIdx.x == 0) s_cta = (int)(atomicAdd(P.ticket_intra, 1u) - P.ticket_intra_base);
  __syncthreads();
  // Programmatic dependent launch: the loop filter of this frame may start as soon as every CTA of
  // this grid is running (or done).  It follows the reconstruction one superblock row behind through
  // prog_done and never blocks a CTA of this grid, so starting it early cannot stall the wavefront.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const int unit = s_cta * kWaveWarps + warp;
  if (unit >= 3 * P.sb_rows) return;
  // units are ordered row-major so that a row's dependency (row-1, same plane) has a lower index
  const int sbr = unit / 3, plane = unit % 3;
  const int nsb = P.sb_cols * P.sb_rows;
  const int n = (plane && P.ss_x) ? 32 : 64;    // superblock width in this plane
  const int nh = (plane && P.ss_y) ? 32 : 64;   // and its height (4:2:2 / 4:4:0: not square)
  const int stride = P.dst[plane].stride;
  Px *gplane = (Px *)P.dst[plane].ptr;
  uint32_t *prog = P.prog_intra + plane * P.sb_rows;
  uint32_t *done = P.prog_done + plane * P.sb_rows;
  const vp9b_sbrange *rgs = P.sb_intra + plane * nsb + sbr * P.sb_cols;
  constexpr int TS = G::kTS, M = G::kMargin;
  Px *T0 = tile + TS + M;  // sample (0,0) of the superblock; row -1 / column -1 hold the neighbours
  bool left_in_smem = false;
  int published = 0;   // superblocks of this row already released to the row below
  constexpr int kPer = 64 * 8 / 32;
  Vec tmp[kPer];       // staged tile interior (also the prefetch of the next non-empty superblock)
  int pf_sbc = -1;
  // Superblocks without intra blocks are final after the inter stages: walk only the non-empty
  // ones (32 ranges at a time, one per lane) and release everything skipped in one store.
  for (int c0 = 0; c0 < P.sb_cols; c0 += 32) {
    vp9b_sbrange mine;
    mine.start = 0; mine.count = 0;
    if (c0 + lane < P.sb_cols) mine = rgs[c0 + lane];
    uint32_t todo = __ballot_sync(0xffffffffu, mine.count != 0);
    while (todo) {
    const int src = __ffs(todo) - 1;
    todo &= todo - 1;
    const int sbc = c0 + src;
    vp9b_sbrange rg;
    rg.start = __shfl_sync(0xffffffffu, mine.start, src);
    rg.count = __shfl_sync(0xffffffffu, mine.count, src);
    if (sbc > published) {
      // everything to the left is done (program order); nothing of it is in flight
      if (lane == 0) {
        st_release(prog + sbr, P.prog_base + (uint32_t)sbc);
        st_release(done + sbr, P.prog_base + (uint32_t)sbc);   // earlier tiles were stored in program order
      }
      left_in_smem = false;
    }
    Px *g = gplane + (size_t)(sbr * nh) * stride + sbc * n;
    // interior rows can be fetched before the dependency wait (nobody else writes them)
    const int vpr = n >> 3, vsh = n == 64 ? 3 : 2, nvec = nh * vpr;
    // all loads of the tile are issued before the first one is consumed (one L2 round trip); they
    // were already issued at the end of the previous superblock when that one ran in this loop
    if (pf_sbc != sbc) {
#pragma unroll
      for (int k = 0; k < kPer; k++) {
        const int i = lane + 32 * k;
        if (i < nvec) tmp[k] = __ldcg((const Vec *)(g + (size_t)(i >> vsh) * stride + (i & (vpr - 1)) * 8));
      }
    }
#pragma unroll
    for (int k = 0; k < kPer; k++) {
      const int i = lane + 32 * k;
      if (i < nvec) *(Vec *)(T0 + (i >> vsh) * TS + (i & (vpr - 1)) * 8) = tmp[k];
    }
    if (sbc > 0 && !left_in_smem)
      for (int r = lane; r < nh; r += 32) T0[r * TS - 1] = __ldcg((const Px *)(g + (size_t)r * stride - 1));
    if (sbr > 0) {
      if (lane == 0) wait_progress(prog + sbr - 1, P.prog_base, min(sbc + 1, P.sb_cols));
      __syncwarp();
      // row above (incl. top-left)
      for (int v = lane; v < vpr; v += 32) *(Vec *)(T0 - TS + v * 8) = __ldcg((const Vec *)(g - stride + v * 8));
      if (lane == 0 && sbc > 0) T0[-TS - 1] = __ldcg((const Px *)(g - stride - 1));
    }
    __syncwarp();
    const vp9b_txb *list = P.txb_intra[plane] + rg.start;
    const int px0 = sbc * n, py0 = sbr * nh;
    typedef typename ResidT<Px>::T R;
    const R *rplane = (const R *)P.resid[plane];
    const int maxv = (1 << P.bd) - 1;
    vp9b_txb t = list[0];
    for (uint32_t i = 0; i < rg.count; i++) {
      const vp9b_txb tn = list[min(i + 1, rg.count - 1)];   // next record: off the dependency chain
      Px *d = T0 + (t.y4 * 4 - py0) * TS + (t.x4 * 4 - px0);
      // the block's residual (k_intra_residual) is fetched before the prediction and added after it
      const int lg = 2 + (t.tx & 3);
      const R *rp = rplane + (size_t)t.y4 * 4 * stride + t.x4 * 4;
      R rv[32];
      if (t.eob) {
        switch (lg) {
          case 2: resid_fetch<R, 2>(rp, stride, lane, rv); break;
          case 3: resid_fetch<R, 3>(rp, stride, lane, rv); break;
          case 4: resid_fetch<R, 4>(rp, stride, lane, rv); break;
          default: resid_fetch<R, 5>(rp, stride, lane, rv); break;
        }
      }
      warp_intra_predict<Px>(P, edge, t, d, TS, lane);
      if (t.eob) {
        switch (lg) {
          case 2: resid_add<Px, R, 2>(d, TS, lane, rv, maxv); break;
          case 3: resid_add<Px, R, 3>(d, TS, lane, rv, maxv); break;
          case 4: resid_add<Px, R, 4>(d, TS, lane, rv, maxv); break;
          default: resid_add<Px, R, 5>(d, TS, lane, rv, maxv); break;
        }
        __syncwarp();
      }
      t = tn;
    }
    __syncwarp();
    // The row below only ever reads this superblock's bottom row (its "above" samples): store
    // that row first and publish -- the release then waits for 64 samples, not for the whole tile
    for (int v = lane; v < vpr; v += 32)
      __stcg((Vec *)(g + (size_t)(nh - 1) * stride + v * 8), *(const Vec *)(T0 + (nh - 1) * TS + v * 8));
    __syncwarp();
    if (lane == 0) st_release(prog + sbr, P.prog_base + (uint32_t)(sbc + 1));
    published = sbc + 1;
    // prefetch the interior of the next non-empty superblock of this chunk: its loads fly while the
    // rest of this tile is written back
    if (todo) {
      pf_sbc = c0 + __ffs(todo) - 1;
      const Px *gn = gplane + (size_t)(sbr * nh) * stride + pf_sbc * n;
#pragma unroll
      for (int k = 0; k < kPer; k++) {
        const int i = lane + 32 * k;
        if (i < nvec) tmp[k] = __ldcg((const Vec *)(gn + (size_t)(i >> vsh) * stride + (i & (vpr - 1)) * 8));
      }
    }
    // the rest of the tile (read again only by the next launch) and the next left neighbour column
#pragma unroll 4
    for (int i = lane; i < nvec - vpr; i += 32)
      __stcg((Vec *)(g + (size_t)(i >> vsh) * stride + (i & (vpr - 1)) * 8), *(const Vec *)(T0 + (i >> vsh) * TS + (i & (vpr - 1)) * 8));
    __syncwarp();
    for (int r = lane; r < nh; r += 32) T0[r * TS - 1] = T0[r * TS + n - 1];
    left_in_smem = true;
    __syncwarp();
    if (lane == 0) st_release(done + sbr, P.prog_base + (uint32_t)(sbc + 1));   // the whole tile is out
    }
  }
  if (lane == 0) {
    if (published < P.sb_cols) st_release(prog + sbr, P.prog_base + (uint32_t)P.sb_cols);
    st_release(done + sbr, P.prog_base + (uint32_t)P.sb_cols);
  }
}

// optional phase timing of two loop-filter workers (compile with -DVP9B_LF_TIMING; tools/lf_phases.py
// reads the stamps back through vp9gpu_debug_read)
#ifdef VP9B_LF_TIMING
__device__ unsigned long long d_lf_dbg[2 * 32 * 16];
VP9B_DEV unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define LFT(k) if (tid == 0 && (sbc == 4 || sbc == 5) && r < 32) d_lf_dbg[((sbc - 4) * 32 + r) * 16 + (k)] = gtime();
#else
#define LFT(k)
#endif

// ------------------------------------------------------------------------------------------
// stage 4: loop filter -- one persistent CTA per superblock column (vp9_lf_sb.cuh has the design
// and the per-thread phase functions; this is the CTA-level schedule and the flag protocol).
// F_edge(r, c) = flags[r * sb_cols + c], value = the frame's epoch once set: the last block column
// of (r, c) and the look-ahead columns (left edge of c+1 applied) are in global memory.
// Producer: all threads store (st.cg), block barrier, one thread fence + st.release.gpu.
// Consumer (worker c+1): one thread polls (relaxed, then one ld.acquire.gpu), block barrier, ld.cg.
template <typename Px, bool SUB, typename FB>
__global__ void __launch_bounds__(LfGeom<Px, SUB>::NT, SUB ? VP9B_LF_MINB : 3) k_lf_sb(const __grid_constant__ FB B) {
  const FrameParams &P = B.f[blockIdx.z];
  if (!P.do_lf) return;
  typedef LfGeom<Px, SUB> G;
  typedef typename LfVecT<Px>::T Vec;
  extern __shared__ __align__(16) unsigned char lf_smem[];   // two tile buffers
  __shared__ vp9b_lfm s_M[2][2];                             // [row parity][this superblock, next in the row]
  __shared__ int s_t;
  const int tid = threadIdx.x;
  if (tid == 0) s_t = (int)(atomicAdd(P.ticket_lf, 1u) - P.ticket_lf_base);
  __syncthreads();
  const int sbc = s_t, sb_cols = P.sb_cols, sb_rows = P.sb_rows;
  if (sbc >= sb_cols) return;   // the launch is as wide as the widest frame of the batch
  LfCtx<Px, SUB> C;
  C.sbc = sbc;
  C.sb_cols = sb_cols; C.mi_rows = P.mi_rows; C.bd = P.bd;
  C.has_next = sbc + 1 < sb_cols;
  C.lf_lim = P.lf_lim;
  C.lf_mblim = P.lf_mblim;
  Px *base[3];
#pragma unroll
  for (int p = 0; p < 3; p++) {
    C.stride[p] = P.dst[p].stride;
    base[p] = (Px *)P.dst[p].ptr + sbc * (p ? G::nC : 64);
  }
  uint32_t *const flags = P.lf_flags;
  const uint32_t epoch = P.lf_epoch;
  auto wait_left = [&](int r) {     // F_edge(r, sbc - 1)
    const uint32_t *f = flags + r * sb_cols + sbc - 1;
    while (ld_acquire(f) != epoch) __nanosleep(32);
  };
  // masks of (r, sbc) and (r, sbc + 1), 40 words each: fetched into a register one superblock ahead
  // (like the samples) and written to shared memory at the end of the iteration
  uint32_t mask_word = 0;
  auto fetch_masks = [&](int r) {
    if (tid < 80) {
      const int which = tid >= 40;
      const uint32_t *src = (const uint32_t *)&P.lfm[r * sb_cols + sbc + ((which && C.has_next) ? 1 : 0)];
      mask_word = __ldg(src + (which ? tid - 40 : tid));
    }
  };
  auto commit_masks = [&](int r) {
    if (tid < 80) ((uint32_t *)&s_M[r & 1][0])[tid] = mask_word;
  };
  auto row_ptrs = [&](int r, Px *g[3]) {
#pragma unroll
    for (int p = 0; p < 3; p++) g[p] = base[p] + (size_t)(r * (p ? G::nC : 64)) * C.stride[p];
  };
  // This kernel may have been launched while the intra wavefront of the same frame is still running
  // (programmatic dependent launch): before the samples of superblock row r are fetched, this
  // superblock and the next one in the row (look-ahead columns) must be completely reconstructed.
  const bool follow = P.do_intra != 0;
  const int need_recon = min(sbc + 2, sb_cols);
  auto wait_recon = [&](int r) {
    if (follow) {
      if (tid < 3) wait_progress(P.prog_done + tid * sb_rows + r, P.prog_base, need_recon);
      __syncthreads();
    }
  };
  // Plan of the whole column, computed once by all threads while the first tile is on its way:
  // per superblock the dependent sets (24 x u16) and the chain flag.  (One superblock's plan is
  // ~400 dependent instructions on one warp: far too long for the per-superblock critical path.)
  uint16_t *const plan = (uint16_t *)(lf_smem + 2 * G::kTileBytes);   // [sb_rows][24]
  uint8_t *const plan_chain = (uint8_t *)(plan + (size_t)sb_rows * 24);  // [sb_rows]
  for (int i = tid; i < sb_rows; i += G::NT) plan_chain[i] = 0;
  __syncthreads();
  {
    constexpr int U = LfPlanUnits<Px, SUB>::value;
    for (int i = tid; i < sb_rows * U; i += G::NT) {
      const int r = i / U, u = i - r * U;
      if (lf_plan_unit<Px, SUB>(P.lfm[r * sb_cols + sbc], sbc > 0, u, plan + r * 24)) plan_chain[r] = 1;
    }
  }
  Vec pre[G::kPre];
  // prologue: first superblock of the column
  fetch_masks(0);
  commit_masks(0);
  row_ptrs(0, C.g);
  C.tile = (Px *)lf_smem;
  wait_recon(0);
  lf_phase_prefetch<Px, SUB>(C.g, C.stride, C.has_next, tid, pre);
  lf_phase_commit<Px, SUB>(C, tid, pre);
  for (int r = 0; r < sb_rows; r++) {
    const bool next_row = r + 1 < sb_rows;
    const int keep = next_row ? 8 : 0;
    Px *const cur = (Px *)(lf_smem + (size_t)(r & 1) * G::kTileBytes), *const nxt = (Px *)(lf_smem + (size_t)((r + 1) & 1) * G::kTileBytes);
    C.tile = cur;
    C.sbr = r;
    C.has_top = r > 0;
    C.M = &s_M[r & 1][0];
    C.Mn = &s_M[r & 1][1];
    row_ptrs(r, C.g);
    __syncthreads();                 // tile + masks of this row are in shared memory
    LFT(0)
    if (next_row) {                  // the next superblock's samples and masks fly while this one is filtered
      Px *g[3];
      row_ptrs(r + 1, g);
      wait_recon(r + 1);             // one superblock row behind the reconstruction
      lf_phase_prefetch<Px, SUB>(g, C.stride, C.has_next, tid, pre);
      fetch_masks(r + 1);
    }
    C.dset = plan + r * 24;
    const bool chain = plan_chain[r] != 0;
    LFT(1)
    lf_phase_vertical(C, tid, false);                    // A
    LFT(2)
    if (chain && sbc > 0) {                              // D before H7: this row serialises through F_edge
      if (tid == 0) wait_left(r);
      __syncthreads();
      lf_phase_merge_left(C, tid);
      lf_phase_vertical(C, tid, true);
    }
    __syncthreads();
    LFT(3)
    lf_phase_horizontal(C, tid, 0);                      // H7
    __syncthreads();
    LFT(4)
    if (C.has_next) {
      lf_phase_next_left_edge(C, tid);                   // E
      __syncthreads();
    }
    LFT(5)
    lf_phase_store(C, tid, 0, keep);
    LFT(6)
    if (C.has_next) {                                    // publish F_edge(r, sbc)
      __syncthreads();
      if (tid == 0) st_release(flags + r * sb_cols + sbc, epoch);
    }
    LFT(7)
    if (!chain && sbc > 0) {                             // D
      if (tid == 0) wait_left(r);
      __syncthreads();
      LFT(8)
      lf_phase_merge_left(C, tid);
      lf_phase_vertical(C, tid, true);
    }
    __syncthreads();
    LFT(9)
    lf_phase_horizontal(C, tid, 1);                      // H
    __syncthreads();
    LFT(10)
    lf_phase_store(C, tid, 1, keep);
    LFT(11)
    if (next_row) {
      lf_phase_carry(C, nxt, tid);
      LfCtx<Px, SUB> N = C;
      N.tile = nxt;
      lf_phase_commit<Px, SUB>(N, tid, pre);
      commit_masks(r + 1);
    }
    LFT(12)
  }
}

// ------------------------------------------------------------------------------------------
// stage 4, 8-bit 4:2:0: the same column workers as k_lf_sb, but a worker is ONE WARP whose lanes
// own four sample rows / columns each and filter them with packed arithmetic (vp9_lf_warp.cuh).
// A quarter of the threads, a third of the instructions, no block barriers: the phases of a
// superblock are separated by __syncwarp().  The next superblock's samples are fetched with
// cp.async into the second tile buffer while the current one is filtered.
template <typename FB>
__global__ void __launch_bounds__(32, 16) k_lf_warp(const __grid_constant__ FB B) {
  const FrameParams &P = B.f[blockIdx.z];
  if (!P.do_lf) return;
  extern __shared__ __align__(16) unsigned char lfw_smem[];   // the tile buffer, then the plan
  __shared__ vp9b_lfm s_M[2][2];                              // [row parity][this superblock, next in the row]
  const int lane = threadIdx.x;
  int sbc = 0;
  if (lane == 0) sbc = (int)(atomicAdd(P.ticket_lf, 1u) - P.ticket_lf_base);
  sbc = __shfl_sync(0xffffffffu, sbc, 0);
  const int sb_cols = P.sb_cols, sb_rows = P.sb_rows;
  if (sbc >= sb_cols) return;   // the launch is as wide as the widest frame of the batch
  LfwCtx C;
  C.sbc = sbc;
  C.sb_cols = sb_cols; C.mi_rows = P.mi_rows;
  C.has_next = sbc + 1 < sb_cols;
  C.lf_lim = P.lf_lim;
  C.lf_mblim = P.lf_mblim;
  uint8_t *base[3];
#pragma unroll
  for (int p = 0; p < 3; p++) {
    C.stride[p] = P.dst[p].stride;
    base[p] = (uint8_t *)P.dst[p].ptr + sbc * (p ? 32 : 64);
  }
  uint32_t *const flags = P.lf_flags;
  const uint32_t epoch = P.lf_epoch;
  auto wait_left = [&](int r) {     // F_edge(r, sbc - 1)
    if (lane == 0) {
      const uint32_t *f = flags + r * sb_cols + sbc - 1;
      while (ld_acquire(f) != epoch) __nanosleep(32);
    }
    __syncwarp();
  };
  // masks of (r, sbc) and (r, sbc + 1), 40 words each, fetched one superblock ahead
  uint32_t mw[3] = { 0, 0, 0 };
  auto fetch_masks = [&](int r) {
#pragma unroll
    for (int k = 0; k < 3; k++) {
      const int i = lane + 32 * k;
      if (i < 80) {
        const int which = i >= 40;
        const uint32_t *src = (const uint32_t *)&P.lfm[r * sb_cols + sbc + ((which && C.has_next) ? 1 : 0)];
        mw[k] = __ldg(src + (which ? i - 40 : i));
      }
    }
  };
  auto commit_masks = [&](int r) {
#pragma unroll
    for (int k = 0; k < 3; k++)
      if (lane + 32 * k < 80) ((uint32_t *)&s_M[r & 1][0])[lane + 32 * k] = mw[k];
  };
  auto row_ptrs = [&](int r, uint8_t *g[3]) {
#pragma unroll
    for (int p = 0; p < 3; p++) g[p] = base[p] + (size_t)(r * (p ? 32 : 64)) * C.stride[p];
  };
  // plan of the whole column (dependent sets + chain flag per superblock), computed once
  uint16_t *const plan = (uint16_t *)(lfw_smem + LfwGeom::kTileBytes);   // [sb_rows][24]
  uint8_t *const plan_chain = (uint8_t *)(plan + (size_t)sb_rows * 24);       // [sb_rows]
  for (int i = lane; i < sb_rows; i += 32) plan_chain[i] = 0;
  for (int i = lane; i < sb_rows * 12; i += 32) ((uint32_t *)plan)[i] = 0;
  __syncwarp();
  for (int i = lane; i < sb_rows * kLfwPlanUnits; i += 32) {
    const int r = i / kLfwPlanUnits, u = i - r * kLfwPlanUnits;
    if (lfw_plan_unit(P.lfm[r * sb_cols + sbc], sbc > 0, u, plan + r * 24)) plan_chain[r] = 1;
  }
  __syncwarp();
  for (int r = lane; r < sb_rows; r += 32)   // bit 1: a dependent set reaches beyond block column 0
    if (lfw_plan_deep(plan + r * 24)) plan_chain[r] |= 2;
  // prologue: first superblock of the column
  fetch_masks(0);
  commit_masks(0);
  row_ptrs(0, C.g);
  C.tile = lfw_smem;
  lfw_phase_fetch(C.tile, C.g, C.stride, C.has_next, lane);
  for (int r = 0; r < sb_rows; r++) {
    const bool next_row = r + 1 < sb_rows;
    const int keep = next_row ? 8 : 0;
    C.sbr = r;
    C.has_top = r > 0;
    C.M = &s_M[r & 1][0];
    C.Mn = &s_M[r & 1][1];
    row_ptrs(r, C.g);
    lfw_fetch_wait();
    __syncwarp();                    // tile + masks of this row are in shared memory
    if (next_row) fetch_masks(r + 1);
    C.dset = plan + r * 24;
    const bool chain = (plan_chain[r] & 1) != 0;
    lfw_phase_vertical(C, lane, false);                  // A
    if (chain && sbc > 0) {                              // D before H7: this row serialises through F_edge
      wait_left(r);
      lfw_phase_merge_left(C, lane);
      __syncwarp();
      lfw_phase_vertical(C, lane, true);
    }
    __syncwarp();
    lfw_phase_horizontal(C, lane, 0);                    // H7
    __syncwarp();
    if (C.has_next) {
      lfw_phase_next_left_edge(C, lane);                 // E
      __syncwarp();
    }
    lfw_phase_store(C, lane, 0, keep);
    if (C.has_next) {                                    // publish F_edge(r, sbc)
      __syncwarp();
      if (lane == 0) st_release(flags + r * sb_cols + sbc, epoch);
    }
    // D + H.  The left edge and what depends on it normally stay inside block column 0: the other
    // block columns are filtered first, which hides the wait for the left neighbour.
    const bool early = !chain && sbc > 0 && !(plan_chain[r] & 2);
    if (early) lfw_phase_horizontal(C, lane, 2);
    if (!chain && sbc > 0) {                             // D
      wait_left(r);
      lfw_phase_merge_left(C, lane);
      __syncwarp();
      lfw_phase_vertical(C, lane, true);
    }
    __syncwarp();
    lfw_phase_horizontal(C, lane, early ? 3 : 1);        // H
    __syncwarp();
    lfw_phase_store(C, lane, 1, keep);
    if (next_row) {
      // one tile buffer per worker (twice as many workers fit on an SM): the bottom rows move to the
      // top, then the next superblock is fetched over the interior; other workers cover the latency
      lfw_phase_carry(C, C.tile, lane);
      commit_masks(r + 1);
      __syncwarp();
      uint8_t *g[3];
      row_ptrs(r + 1, g);
      lfw_phase_fetch(C.tile, g, C.stride, C.has_next, lane);
    }
  }
}

// ------------------------------------------------------------------------------------------
// stage 4b, 4:2:2 / 4:4:0 only: loop filter of the two chroma planes (LF_PATH_SLOW:
// vp9_filter_block_plane_non420, vp9_loopfilter.c:1180-1239).  These samplings are rare; the kernel
// is the reference order itself -- superblocks in raster order, per superblock all vertical edges
// (thread = sample row, filter_selectively_vert :652-714) then all horizontal edges (thread = sample
// column, filter_selectively_horiz :453-650) -- one CTA per plane, working in global memory.  The luma
// plane of such frames goes through k_lf_sb with empty chroma masks.
template <typename Px, typename FB>
__global__ void __launch_bounds__(64) k_lf_uv_non420(const __grid_constant__ FB B) {
  const FrameParams &P = B.f[blockIdx.z];
  if (!P.do_lf || !P.lfm_uv) return;
  const int plane = 1 + blockIdx.x, tid = threadIdx.x;
  const int ss_x = P.ss_x, ss_y = P.ss_y, nw = 64 >> ss_x, nh = 64 >> ss_y, stride = P.dst[plane].stride;
  Px *const base = (Px *)P.dst[plane].ptr;
  __shared__ vp9b_lfm s_lvl;      // only lfl_y is used: levels compacted to [band][block column] for lf_h_unit
  __shared__ vp9b_lfm_uv s_m;
  for (int sbr = 0; sbr < P.sb_rows; sbr++)
    for (int sbc = 0; sbc < P.sb_cols; sbc++) {
      __syncthreads();             // previous superblock finished (its samples and the staged masks)
      if (tid < 32) ((uint32_t *)&s_m)[tid] = __ldg((const uint32_t *)&P.lfm_uv[sbr * P.sb_cols + sbc] + tid);
      __syncthreads();
      {
        const int b = tid >> 3, c = tid & 7;     // band b = mask row b << ss_y
        s_lvl.lfl_y[tid] = (b << ss_y) < 8 ? s_m.lfl[((b << ss_y) << 3) + c] : 0;
      }
      Px *const sb = base + (size_t)(sbr * nh) * stride + sbc * nw;
      // vertical edges: thread = sample row of the superblock
      if (tid < nh) {
        const int r = (tid >> 3) << ss_y;        // mask row of this band
        if (sbr * 8 + r < P.mi_rows) {
          uint32_t m16 = s_m.v16[r], m8 = s_m.v8[r], m4 = s_m.v4[r], m4i = s_m.v4i[r];
          Px *s = sb + (size_t)tid * stride;
          const uint8_t *lfl = &s_m.lfl[r << 3];
          for (uint32_t mask = m16 | m8 | m4 | m4i; mask; mask >>= 1) {
            const int lvl = *lfl;
            if (m16 & 1) lf_px<Px>(s, 1, 16, P.lf_mblim[lvl], P.lf_lim[lvl], lvl >> 4, P.bd);
            else if (m8 & 1) lf_px<Px>(s, 1, 8, P.lf_mblim[lvl], P.lf_lim[lvl], lvl >> 4, P.bd);
            else if (m4 & 1) lf_px<Px>(s, 1, 4, P.lf_mblim[lvl], P.lf_lim[lvl], lvl >> 4, P.bd);
            if (m4i & 1) lf_px<Px>(s + 4, 1, 4, P.lf_mblim[lvl], P.lf_lim[lvl], lvl >> 4, P.bd);
            s += 8;
            lfl += 1;
            m16 >>= 1; m8 >>= 1; m4 >>= 1; m4i >>= 1;
          }
        }
      }
      __syncthreads();
      // horizontal edges: thread = sample column; masks compacted to one row of 8 bits per band
      if (tid < nw) {
        LfHMasks H;
        H.a16 = H.a8 = H.a4 = H.ai = 0;
        H.rsh = 3; H.lstep = 1; H.nrow = 8;
        for (int b = 0; b < (8 >> ss_y); b++) {
          const int r = b << ss_y;
          if (sbr * 8 + r >= P.mi_rows) break;
          H.a16 |= (uint64_t)s_m.h16[r] << (8 * b);
          H.a8 |= (uint64_t)s_m.h8[r] << (8 * b);
          H.a4 |= (uint64_t)s_m.h4[r] << (8 * b);
          H.ai |= (uint64_t)s_m.h4i[r] << (8 * b);
        }
        const int c = tid >> 3;
        uint64_t todo = ((H.a16 | H.a8 | H.a4 | H.ai) >> c) & 0x0101010101010101ull;
        while (todo) {
          const int bit0 = (todo & 0xffffffffu) ? __ffs((uint32_t)todo) - 1 : 32 + __ffs((uint32_t)(todo >> 32)) - 1;
          todo &= todo - 1;
          lf_h_unit<Px>(H, s_lvl, tid, bit0 >> 3, sb + tid, stride, P.lf_lim, P.lf_mblim, P.bd);
        }
      }
    }
}

// ------------------------------------------------------------------------------------------
// output path: scaling of a shown 4:2:0 frame to the display size (vpxdec --scale: vpx_image_scale ->
// libyuv I420Scale / I420Scale_16 with kFilterBox, vpxdec.c:910-949).  For every size change within
// a factor of two libyuv reduces the box filter to its bilinear paths (ScaleFilterReduce,
// scale_common.cc:1162-1203); those two are reproduced bit-exactly, one thread per output sample:
//   up   (ScalePlaneBilinearUp, scale.cc:1128-1262): rows yi and yi+1 are filtered horizontally
//        (ScaleFilterCols_C: a + ((f >> 9) * (b - a) + 64 >> 7), 16-bit samples: full 16-bit fraction),
//        then blended vertically (InterpolateRow_C: (a * (256 - yf) + b * yf + 128) >> 8; the 16-bit
//        row function does not round, except for yf == 128);
//   down (ScalePlaneBilinearDown, :928-1034): vertical blend first, then the horizontal filter.
// Positions are libyuv's 16.16 steps (ScaleSlope, scale_common.cc:1218-1288), computed on the host.
struct ScalePlan {
  int x0, dx, y0, dy;   // 16.16
  int up;               // ScalePlaneBilinearUp (dst_height > src_height) or ...Down
  int linear;           // kFilterLinear: no vertical blend
};

template <typename Px>
VP9B_DEV int scale_hblend(int a, int b, int f) {
  if (sizeof(Px) == 1) return a + ((((f >> 9) * (b - a)) + 0x40) >> 7);
  return a + (int)((((long long)f * (b - a)) + 0x8000) >> 16);
}
template <typename Px>
VP9B_DEV int scale_vblend(int a, int b, int yf) {
  if (yf == 0) return a;
  if (yf == 128) return (a + b + 1) >> 1;
  if (sizeof(Px) == 1) return (a * (256 - yf) + b * yf + 128) >> 8;
  return (a * (256 - yf) + b * yf) >> 8;
}

template <typename Px>
__global__ void k_scale_plane(const Px *__restrict__ src, int sstride, int sh, Px *__restrict__ dst, int dstride, int dw, int dh,
                              ScalePlan pl) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
  if (k >= dw || j >= dh) return;
  const long long max_y = (long long)(sh - 1) << 16;
  long long y = (long long)pl.y0 + (long long)j * pl.dy;
  if (y > max_y) y = max_y;
  const long long x = (long long)pl.x0 + (long long)k * pl.dx;
  const int yi = (int)(y >> 16), yf = pl.linear ? 0 : (int)(y >> 8) & 255;
  const int xi = (int)(x >> 16), f = (int)(x & 0xffff);
  const Px *r0 = src + (size_t)yi * sstride + xi, *r1 = r0 + (yf ? sstride : 0);
  const int a0 = r0[0], b0 = r0[1], a1 = r1[0], b1 = r1[1];
  int v;
  if (pl.up) v = scale_vblend<Px>(scale_hblend<Px>(a0, b0, f), scale_hblend<Px>(a1, b1, f), yf);
  else v = scale_hblend<Px>(scale_vblend<Px>(a0, a1, yf), scale_vblend<Px>(b0, b1, yf), f);
  dst[(size_t)j * dstride + k] = (Px)v;
}

// output path: 16-bit samples -> 8 bits (truncating shift, highbd_img_downshift in tools_common.c)
__global__ void k_downshift(const uint16_t *__restrict__ src, int sstride, uint8_t *__restrict__ dst, int dstride, int w,
                            int h, int shift) {
  const int x = (blockIdx.x * blockDim.x + threadIdx.x) * 4, y = blockIdx.y;
  if (x >= w || y >= h) return;
  const uint16_t *s = src + (size_t)y * sstride + x;
  uint8_t *d = dst + (size_t)y * dstride + x;
  if (x + 4 <= w) {
    const uint2 v = *(const uint2 *)s;   // plane rows are 128-byte aligned, x is a multiple of 4
    const uint32_t a = (v.x & 0xffffu) >> shift, b = (v.x >> 16) >> shift, c = (v.y & 0xffffu) >> shift, e = (v.y >> 16) >> shift;
    *(uint32_t *)d = a | (b << 8) | (c << 16) | (e << 24);
  } else {
    for (int i = 0; x + i < w; i++) d[i] = (uint8_t)(s[i] >> shift);
  }
}

}  // namespace vp9b

// ==========================================================================================
// host side: context, frame pool, launch
using namespace vp9b;

struct DevFrame {
  void *base = nullptr;
  size_t bytes = 0;
  PlaneGeom pl[3];
  uint32_t width = 0, height = 0;
  int bd = 0, ss_x = 0, ss_y = 0, bps = 1;
  bool valid = false;
  // geometry the slot's TMA descriptors were encoded for
  void *tm_base = nullptr;
  int tm_stride[3] = { 0, 0, 0 }, tm_rows[3] = { 0, 0, 0 }, tm_bps = 0;
};

struct vp9gpu_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  std::string err;
  DevFrame slot[VP9B_NUM_SLOTS];
  // upload ring
  static const int kRing = 3;
  void *h_pinned[kRing] = { nullptr, nullptr, nullptr };
  void *d_packed[kRing] = { nullptr, nullptr, nullptr };
  size_t ring_bytes[kRing] = { 0, 0, 0 };
  cudaEvent_t ring_ev[kRing];
  int ring_pos = 0;
  // per-geometry workspace
  size_t cap_mi = 0, cap_sb = 0;  // capacities of d_mi (8x8 cells) and d_lf_flags (superblocks)
  MiEntry *d_mi = nullptr;
  uint32_t *d_tickets = nullptr;  // wavefront progress counters
  uint32_t *d_lf_flags = nullptr; // loop-filter data-flow flags, 1 per superblock
  CUtensorMap *d_tmaps = nullptr; // [slot][plane][class] TMA descriptors of the frame pool (motion compensation)
  bool tma_ok = true;             // false once the driver refused a descriptor: kernels then read global memory only
  uint32_t epoch = 0;
  cudaEvent_t join_ev = nullptr;  // joins this context's stream with a batch leader's (submit_device_multi)
  unsigned attr_done = 0;         // kernels whose dynamic shared-memory limit has been raised on this device
  uint8_t *d_out8 = nullptr;      // staging for 8-bit output of high-bit-depth frames
  uint8_t *d_resid = nullptr;     // residual planes of intra blocks
  size_t resid_bytes = 0;
  size_t out8_bytes = 0;
  uint32_t ticket_lf_base = 0, ticket_intra_base = 0;   // running totals of tickets handed out
  // tickets (vp9gpu_mark / vp9gpu_wait)
  static const int kTickets = 64;
  cudaEvent_t ticket_ev[kTickets];
  bool tickets_ready = false;
  uint64_t next_ticket = 1;
  int stop_stage = 0;
  bool lf_overlap = getenv("VP9B_LF_NO_OVERLAP") == nullptr;   // loop filter launched programmatically behind k_intra
  bool mc_flat = getenv("VP9B_MC_FLAT") != nullptr;   // A/B: motion compensation without staging (k_inter_pred_flat)
  bool lf_warp = getenv("VP9B_LF_WARP") != nullptr;   // A/B: 8-bit 4:2:0 through k_lf_warp instead of k_lf_sb
  bool validate = true;           // record-level validation of host packed buffers (vp9gpu_set_validation)
  uint64_t launches = 0;
  int num_sms = 148;
  bool profiling = false;
  cudaEvent_t prof_ev[7];
  float stage_ms[8] = { 0 };
  bool tables_ready = false;
};

// NVTX range for the lifetime of a scope (no-op unless a profiler is attached)
struct NvtxRange {
  explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};

#define CK(call)                                                                       \
  do {                                                                                 \
    cudaError_t e_ = (call);                                                           \
    if (e_ != cudaSuccess) {                                                           \
      ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_);                   \
      return VP9B_ERR_CUDA;                                                            \
    }                                                                                  \
  } while (0)

#define CK2(c, call)                                                                   \
  do {                                                                                 \
    cudaError_t e_ = (call);                                                           \
    if (e_ != cudaSuccess) {                                                           \
      (c)->err = std::string(#call) + ": " + cudaGetErrorString(e_);                  \
      return VP9B_ERR_CUDA;                                                            \
    }                                                                                  \
  } while (0)

static int init_tables(vp9gpu_ctx *ctx) {
  std::vector<int16_t> scan;
  auto app = [&](const int16_t *p, int n) { scan.insert(scan.end(), p, p + n); };
  app(kDefaultScan_4x4, 16); app(kRowScan_4x4, 16); app(kColScan_4x4, 16);
  app(kDefaultScan_8x8, 64); app(kRowScan_8x8, 64); app(kColScan_8x8, 64);
  app(kDefaultScan_16x16, 256); app(kRowScan_16x16, 256); app(kColScan_16x16, 256);
  app(kDefaultScan_32x32, 1024);
  CK(cudaMemcpyToSymbol(d_scan, scan.data(), scan.size() * sizeof(int16_t)));
  CK(cudaMemcpyToSymbol(d_filters16, kSubpelFilters, sizeof(int16_t) * 4 * 16 * 8));
  int8_t f8[4 * 16 * 8];
  for (int i = 0; i < 4 * 16 * 8; i++) f8[i] = (int8_t)(kSubpelFilters[i] > 127 ? 127 : kSubpelFilters[i]);
  uint32_t t4[4 * 16 * 2];
  for (int i = 0; i < 4 * 16; i++)
    for (int hhalf = 0; hhalf < 2; hhalf++) {
      uint32_t w = 0;
      for (int k = 0; k < 4; k++) w |= (uint32_t)(uint8_t)(int8_t)f8[i * 8 + hhalf * 4 + k] << (8 * k);
      t4[i * 2 + hhalf] = w;
    }
  CK(cudaMemcpyToSymbol(d_taps4, t4, sizeof(t4)));
  ctx->tables_ready = true;
  return 0;
}

extern "C" vp9gpu_ctx *vp9gpu_create(int device) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) return nullptr;
  vp9gpu_ctx *ctx = new vp9gpu_ctx();
  if (device < 0) cudaGetDevice(&device);
  ctx->device = device;
  if (cudaSetDevice(device) != cudaSuccess) { delete ctx; return nullptr; }
  if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return nullptr; }
  for (int i = 0; i < vp9gpu_ctx::kRing; i++) cudaEventCreateWithFlags(&ctx->ring_ev[i], cudaEventDisableTiming);
  for (int i = 0; i < 7; i++) cudaEventCreate(&ctx->prof_ev[i]);
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) ctx->num_sms = prop.multiProcessorCount;
  if (init_tables(ctx)) { fprintf(stderr, "vp9gpu_create: %s\n", ctx->err.c_str()); delete ctx; return nullptr; }
  return ctx;
}

extern "C" void vp9gpu_destroy(vp9gpu_ctx *ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  for (auto &s : ctx->slot) cudaFree(s.base);
  for (int i = 0; i < vp9gpu_ctx::kRing; i++) {
    cudaFreeHost(ctx->h_pinned[i]);
    cudaFree(ctx->d_packed[i]);
    cudaEventDestroy(ctx->ring_ev[i]);
  }
  for (int i = 0; i < 7; i++) cudaEventDestroy(ctx->prof_ev[i]);
  if (ctx->tickets_ready)
    for (int i = 0; i < vp9gpu_ctx::kTickets; i++) cudaEventDestroy(ctx->ticket_ev[i]);
  if (ctx->join_ev) cudaEventDestroy(ctx->join_ev);
  cudaFree(ctx->d_mi);
  cudaFree(ctx->d_out8);
  cudaFree(ctx->d_resid);
  cudaFree(ctx->d_tickets);
  cudaFree(ctx->d_lf_flags);
  cudaFree(ctx->d_tmaps);
  cudaStreamDestroy(ctx->stream);
  delete ctx;
}

extern "C" const char *vp9gpu_last_error(vp9gpu_ctx *ctx) { return ctx ? ctx->err.c_str() : "null ctx"; }
extern "C" void *vp9gpu_stream(vp9gpu_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }
extern "C" uint64_t vp9gpu_launch_count(vp9gpu_ctx *ctx) { return ctx ? ctx->launches : 0; }
extern "C" int vp9gpu_set_stop_stage(vp9gpu_ctx *ctx, int stage) {
  if (!ctx) return VP9B_ERR_INVALID;
  ctx->stop_stage = stage;
  return 0;
}
extern "C" int vp9gpu_set_validation(vp9gpu_ctx *ctx, int enable) {
  if (!ctx) return VP9B_ERR_INVALID;
  ctx->validate = enable != 0;
  return 0;
}
extern "C" int vp9gpu_set_profiling(vp9gpu_ctx *ctx, int enable) {
  if (!ctx) return VP9B_ERR_INVALID;
  ctx->profiling = enable != 0;
  return 0;
}
extern "C" int vp9gpu_get_stage_ms(vp9gpu_ctx *ctx, float ms[8], int reset) {
  if (!ctx) return VP9B_ERR_INVALID;
  memcpy(ms, ctx->stage_ms, sizeof(ctx->stage_ms));
  if (reset) memset(ctx->stage_ms, 0, sizeof(ctx->stage_ms));
  return 0;
}
#ifdef VP9B_LF_TIMING
extern "C" int vp9gpu_debug_read(unsigned long long *out) { return (int)cudaMemcpyFromSymbol(out, vp9b::d_lf_dbg, sizeof(unsigned long long) * 1024); }
#endif
extern "C" int vp9gpu_get_frame_8bit_async(vp9gpu_ctx *ctx, int slot, uint8_t *y, uint8_t *u, uint8_t *v, int y_stride,
                                           int uv_stride) {
  if (!ctx || slot < 0 || slot >= VP9B_NUM_SLOTS || !ctx->slot[slot].valid) return VP9B_ERR_INVALID;
  const DevFrame &f = ctx->slot[slot];
  if (f.bd == 8) return vp9gpu_get_frame_async(ctx, slot, y, u, v, y_stride, uv_stride);
  CK(cudaSetDevice(ctx->device));
  int pw[3], ph[3], ps[3];
  size_t off[3], total = 0;
  for (int p = 0; p < 3; p++) {
    pw[p] = p ? (int)((f.width + f.ss_x) >> f.ss_x) : (int)f.width;
    ph[p] = p ? (int)((f.height + f.ss_y) >> f.ss_y) : (int)f.height;
    ps[p] = (pw[p] + 127) & ~127;
    off[p] = total;
    total += (size_t)ps[p] * ph[p];
  }
  if (ctx->out8_bytes < total) {
    CK(cudaStreamSynchronize(ctx->stream));
    cudaFree(ctx->d_out8);
    ctx->d_out8 = nullptr;
    ctx->out8_bytes = 0;
    CK(cudaMalloc(&ctx->d_out8, total));
    ctx->out8_bytes = total;
  }
  uint8_t *out[3] = { y, u, v };
  for (int p = 0; p < 3; p++) {
    dim3 blk(64), grd((pw[p] / 4 + 64) / 64, ph[p]);
    k_downshift<<<grd, blk, 0, ctx->stream>>>((const uint16_t *)f.pl[p].ptr, f.pl[p].stride, ctx->d_out8 + off[p], ps[p], pw[p],
                                              ph[p], f.bd - 8);
    ctx->launches++;
    CK(cudaMemcpy2DAsync(out[p], p ? uv_stride : y_stride, ctx->d_out8 + off[p], ps[p], pw[p], ph[p], cudaMemcpyDeviceToHost,
                         ctx->stream));
  }
  CK(cudaGetLastError());
  return 0;
}

// libyuv's choice of algorithm and 16.16 steps for one plane (ScaleFilterReduce + ScalePlane's dispatch
// + ScaleSlope); false: a path that is not reproduced here (box kernels of the 3/4, 1/2, 3/8, 1/4 and
// < 1/2 ratios, height-only changes, one-sample planes)
static bool scale_plan(int sw, int sh, int dw, int dh, ScalePlan &p) {
  memset(&p, 0, sizeof(p));
  if (sw < 2 || sh < 1 || dw < 1 || dh < 1) return false;
  if (!(dw * 2 >= sw && dh * 2 >= sh)) return false;             // stays kFilterBox: ScalePlaneBox
  bool linear = sh == 1 || dh == sh || dh * 3 == sh;
  if (linear && (dw == sw || dw * 3 == sw)) return false;        // reduced to kFilterNone
  if (dw == sw) return false;                                     // ScalePlaneVertical
  if (dw <= sw && dh <= sh) {
    if ((4 * dw == 3 * sw && 4 * dh == 3 * sh) || (2 * dw == sw && 2 * dh == sh) || (8 * dw == 3 * sw && 8 * dh == 3 * sh)) return false;
  }
  auto fixed_div = [](int num, int div) { return (int)(((int64_t)num << 16) / div); };
  auto fixed_div1 = [](int num, int div) { return (int)((((int64_t)num << 16) - 0x00010001) / (div - 1)); };
  if (dw <= sw) { p.dx = fixed_div(sw, dw); p.x0 = (p.dx >> 1) - 32768; }
  else if (dw > 1) { p.dx = fixed_div1(sw, dw); p.x0 = 0; }
  if (linear) { p.dy = fixed_div(sh, dh); p.y0 = p.dy >> 1; }
  else if (dh <= sh) { p.dy = fixed_div(sh, dh); p.y0 = (p.dy >> 1) - 32768; }
  else if (dh > 1) { p.dy = fixed_div1(sh, dh); p.y0 = 0; }
  p.linear = linear;
  p.up = dh > sh;
  // both paths clamp y to the last row only from above; a negative start cannot happen (dy >= 65536 when centred)
  return true;
}

extern "C" int vp9gpu_get_frame_scaled_async(vp9gpu_ctx *ctx, int slot, uint32_t dst_w, uint32_t dst_h, uint8_t *y, uint8_t *u,
                                             uint8_t *v, int y_stride, int uv_stride) {
  if (!ctx || slot < 0 || slot >= VP9B_NUM_SLOTS || !ctx->slot[slot].valid || !dst_w || !dst_h) return VP9B_ERR_INVALID;
  const DevFrame &f = ctx->slot[slot];
  if (f.width == dst_w && f.height == dst_h) return vp9gpu_get_frame_async(ctx, slot, y, u, v, y_stride, uv_stride);
  if (!(f.ss_x == 1 && f.ss_y == 1)) { ctx->err = "output scaling: 4:2:0 only (as vpxdec's libyuv_scale)"; return VP9B_ERR_UNSUP; }
  CK(cudaSetDevice(ctx->device));
  ScalePlan plan[3];
  int sw[3], sh[3], dw[3], dh[3], ps[3];
  size_t off[3], total = 0;
  for (int p = 0; p < 3; p++) {
    sw[p] = p ? (int)((f.width + 1) >> 1) : (int)f.width;
    sh[p] = p ? (int)((f.height + 1) >> 1) : (int)f.height;
    dw[p] = p ? (int)((dst_w + 1) >> 1) : (int)dst_w;
    dh[p] = p ? (int)((dst_h + 1) >> 1) : (int)dst_h;
    if (!scale_plan(sw[p], sh[p], dw[p], dh[p], plan[p])) {
      ctx->err = "output scaling: this size ratio takes one of libyuv's box kernels, which are not implemented";
      return VP9B_ERR_UNSUP;
    }
    ps[p] = (dw[p] * f.bps + 127) & ~127;
    off[p] = total;
    total += (size_t)ps[p] * dh[p];
  }
  if (ctx->out8_bytes < total) {     // shares the staging buffer of the 8-bit output path
    CK(cudaStreamSynchronize(ctx->stream));
    cudaFree(ctx->d_out8);
    ctx->d_out8 = nullptr;
    ctx->out8_bytes = 0;
    CK(cudaMalloc(&ctx->d_out8, total));
    ctx->out8_bytes = total;
  }
  uint8_t *out[3] = { y, u, v };
  for (int p = 0; p < 3; p++) {
    const dim3 blk(128), grd((dw[p] + 127) / 128, dh[p]);
    if (f.bps == 1)
      k_scale_plane<uint8_t><<<grd, blk, 0, ctx->stream>>>((const uint8_t *)f.pl[p].ptr, f.pl[p].stride, sh[p], ctx->d_out8 + off[p], ps[p],
                                                             dw[p], dh[p], plan[p]);
    else
      k_scale_plane<uint16_t><<<grd, blk, 0, ctx->stream>>>((const uint16_t *)f.pl[p].ptr, f.pl[p].stride, sh[p],
                                                              (uint16_t *)(ctx->d_out8 + off[p]), ps[p] / 2, dw[p], dh[p], plan[p]);
    ctx->launches++;
    CK(cudaMemcpy2DAsync(out[p], p ? uv_stride : y_stride, ctx->d_out8 + off[p], ps[p], (size_t)dw[p] * f.bps, dh[p],
                         cudaMemcpyDeviceToHost, ctx->stream));
  }
  CK(cudaGetLastError());
  return 0;
}

extern "C" int vp9gpu_mark(vp9gpu_ctx *ctx, uint64_t *ticket) {
  if (!ctx || !ticket) return VP9B_ERR_INVALID;
  CK(cudaSetDevice(ctx->device));
  if (!ctx->tickets_ready) {
    for (int i = 0; i < vp9gpu_ctx::kTickets; i++) CK(cudaEventCreateWithFlags(&ctx->ticket_ev[i], cudaEventDisableTiming));
    ctx->tickets_ready = true;
  }
  const uint64_t t = ctx->next_ticket++;
  CK(cudaEventRecord(ctx->ticket_ev[t % vp9gpu_ctx::kTickets], ctx->stream));
  *ticket = t;
  return 0;
}
extern "C" int vp9gpu_wait(vp9gpu_ctx *ctx, uint64_t ticket) {
  if (!ctx || ticket == 0 || ticket >= ctx->next_ticket) return VP9B_ERR_INVALID;
  CK(cudaSetDevice(ctx->device));
  // a ticket whose event slot has been re-recorded is older than that record: stream order makes
  // waiting for the newer record sufficient
  CK(cudaEventSynchronize(ctx->ticket_ev[ticket % vp9gpu_ctx::kTickets]));
  return 0;
}
extern "C" int vp9gpu_sync(vp9gpu_ctx *ctx) {
  if (!ctx) return VP9B_ERR_INVALID;
  CK(cudaStreamSynchronize(ctx->stream));
  return 0;
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                    const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
// cuTensorMapEncodeTiled through the runtime (the library does not link libcuda)
static PFN_encodeTiled get_encode_tiled() {
  void *fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
    return nullptr;
  return (PFN_encodeTiled)fn;
}

// device frame geometry: planes padded by 64 samples right/bottom (transform blocks may hang
// over the aligned frame edge, as they do into libvpx's frame border), stride multiple of 128 B.
static int setup_slot(vp9gpu_ctx *ctx, int s, uint32_t w, uint32_t h, int bd, int ss_x, int ss_y) {
  DevFrame &f = ctx->slot[s];
  const int bps = bd > 8 ? 2 : 1;
  const int aw = (w + 7) & ~7, ah = (h + 7) & ~7;
  size_t off[3], total = 0;
  int stride[3];
  for (int p = 0; p < 3; p++) {
    const int sx = p ? ss_x : 0, sy = p ? ss_y : 0;
    const int st = (((aw >> sx) + 64) * bps + 127) & ~127;
    const int rows = (ah >> sy) + 64;
    stride[p] = st / bps;
    off[p] = total;
    total += (size_t)st * rows + 256;
  }
  if (f.bytes < total) {
    CK(cudaStreamSynchronize(ctx->stream));
    cudaFree(f.base);
    f.base = nullptr;
    f.bytes = 0;
    CK(cudaMalloc(&f.base, total));
    CK(cudaMemsetAsync(f.base, 0, total, ctx->stream));
    f.bytes = total;
  }
  for (int p = 0; p < 3; p++) {
    f.pl[p].ptr = (uint8_t *)f.base + off[p];
    f.pl[p].stride = stride[p];
  }
  f.width = w; f.height = h; f.bd = bd; f.ss_x = ss_x; f.ss_y = ss_y; f.bps = bps;
  f.valid = true;
  // TMA descriptors: per plane one box per staging-unit size (8, 16, 32, 64 samples + 7 taps).  The
  // tensor is the whole padded plane allocation, so a box that hangs over the picture reads padding,
  // never another allocation; the kernel only stages windows that lie inside the picture anyway.
  int rows[3];
  for (int p = 0; p < 3; p++) rows[p] = (ah >> (p ? ss_y : 0)) + 64;
  const bool same = f.tm_base == f.base && f.tm_bps == bps && !memcmp(f.tm_stride, stride, sizeof(stride)) &&
                    !memcmp(f.tm_rows, rows, sizeof(rows));
  if (!same && ctx->tma_ok) {
    static PFN_encodeTiled encode = get_encode_tiled();
    if (!encode) ctx->tma_ok = false;
    if (ctx->tma_ok && !ctx->d_tmaps) {
      CK(cudaMalloc(&ctx->d_tmaps, sizeof(CUtensorMap) * VP9B_NUM_SLOTS * 12));
      CK(cudaMemsetAsync(ctx->d_tmaps, 0, sizeof(CUtensorMap) * VP9B_NUM_SLOTS * 12, ctx->stream));
    }
    alignas(64) CUtensorMap maps[12];
    for (int p = 0; p < 3 && ctx->tma_ok; p++)
      for (int k = 0; k < 4; k++) {
        const int S = 8 << k;
        const int boxw = bps == 1 ? McStage<uint8_t>::boxw(S) : McStage<uint16_t>::boxw(S);
        const cuuint64_t dims[2] = { (cuuint64_t)stride[p], (cuuint64_t)rows[p] };
        const cuuint64_t strides[1] = { (cuuint64_t)stride[p] * bps };
        const cuuint32_t box[2] = { (cuuint32_t)boxw, (cuuint32_t)(S + 7) }, estr[2] = { 1, 1 };
        const CUresult r = encode(&maps[p * 4 + k], bps == 1 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, f.pl[p].ptr, dims,
                                  strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { ctx->tma_ok = false; break; }
      }
    if (ctx->tma_ok) {
      // stream ordered: kernels already queued keep seeing the descriptors they were launched with
      CK(cudaMemcpyAsync(ctx->d_tmaps + (size_t)s * 12, maps, sizeof(maps), cudaMemcpyHostToDevice, ctx->stream));
      f.tm_base = f.base; f.tm_bps = bps;
      memcpy(f.tm_stride, stride, sizeof(stride));
      memcpy(f.tm_rows, rows, sizeof(rows));
    }
  }
  return 0;
}

// Per-geometry workspace.  Capacities only grow: a stream whose coded size changes (scaled reference
// frames) must not pay a device-wide synchronisation (cudaFree) at every change.  Stale content is
// harmless: k_expand_blocks rewrites every cell of the motion map of an inter frame, and the
// loop-filter flags compare against the frame's epoch.
static int setup_workspace(vp9gpu_ctx *ctx, const vp9b_frame_hdr *h) {
  const size_t need_mi = (size_t)h->mi_cols * h->mi_rows, need_sb = (size_t)h->sb_cols * h->sb_rows;
  if (need_mi > ctx->cap_mi || need_sb > ctx->cap_sb || !ctx->d_tickets) {
    CK(cudaStreamSynchronize(ctx->stream));
    if (need_mi > ctx->cap_mi) {
      cudaFree(ctx->d_mi);
      ctx->d_mi = nullptr;
      ctx->cap_mi = 0;
      CK(cudaMalloc(&ctx->d_mi, sizeof(MiEntry) * need_mi));
      CK(cudaMemsetAsync(ctx->d_mi, 0, sizeof(MiEntry) * need_mi, ctx->stream));
      ctx->cap_mi = need_mi;
    }
    if (need_sb > ctx->cap_sb) {
      cudaFree(ctx->d_lf_flags);
      ctx->d_lf_flags = nullptr;
      ctx->cap_sb = 0;
      CK(cudaMalloc(&ctx->d_lf_flags, sizeof(uint32_t) * 2 * need_sb));
      CK(cudaMemsetAsync(ctx->d_lf_flags, 0, sizeof(uint32_t) * 2 * need_sb, ctx->stream));
      ctx->cap_sb = need_sb;
    }
    if (!ctx->d_tickets) {
      CK(cudaMalloc(&ctx->d_tickets, sizeof(uint32_t) * (4096 + 64)));
      CK(cudaMemsetAsync(ctx->d_tickets, 0, sizeof(uint32_t) * (4096 + 64), ctx->stream));
    }
    CK(cudaStreamSynchronize(ctx->stream));
  }
  return 0;
}

// One launch per stage for a batch of n prepared frames (one per context, all of the same sample
// type and chroma format) on the stream of the first context.  Grids are the maximum over the
// batch; a frame that does not take part in a stage has its flag cleared and its CTAs return.
template <typename Px, typename PT, typename FB>
static int launch_batch(vp9gpu_ctx *const *ctxs, FB &B, int n) {
  constexpr unsigned abase = FB::kN == kMaxBatch ? 0u : 16u;   // attribute bits of the small-block instantiations
  vp9gpu_ctx *ctx = ctxs[0];
  cudaStream_t st = ctx->stream;
  NvtxRange nvtx_range(n == 1 ? "vp9b: reconstruct frame" : "vp9b: reconstruct frame batch");   // visible in nsys / ncu timelines
  const int stop = ctx->stop_stage;
  const bool prof = ctx->profiling;
  int g_expand = 0, g_mcx = 0, g_mcy = 0, g_res = 0, g_ires = 0, g_intra = 0, g_lf = 0;
  for (int i = 0; i < n; i++) {
    FrameParams &P = B.f[i];
    if (P.do_inter) {
      g_expand = std::max(g_expand, (P.n_blocks * 8 + 127) / 128);
      g_mcx = std::max(g_mcx, P.sb_cols);
      g_mcy = std::max(g_mcy, P.sb_rows);
    }
    P.do_res = P.do_inter && stop != 1 && P.n_txb_inter > 0;
    if (P.do_res) {
      const int tasks = ((P.n_txb_size[0] + 7) >> 3) + ((P.n_txb_size[1] + 3) >> 2) + ((P.n_txb_size[2] + 1) >> 1) + P.n_txb_size[3];
      g_res = std::max(g_res, std::min((tasks + kResWarps - 1) / kResWarps, ctx->num_sms * 8));
    }
    const int n_intra = P.n_txb_intra[0] + P.n_txb_intra[1] + P.n_txb_intra[2];
    P.do_intra = stop != 1 && stop != 2 && n_intra > 0;
    if (P.do_intra) {
      g_ires = std::max(g_ires, std::min((n_intra + kResWarps - 1) / kResWarps, ctx->num_sms * 8));
      g_intra = std::max(g_intra, (3 * P.sb_rows + kWaveWarps - 1) / kWaveWarps);
    }
    P.do_lf = P.do_lf && stop == 0;
    if (P.do_lf) g_lf = std::max(g_lf, P.sb_cols);
  }
  // 8-bit 4:2:0 can take the warp-per-column loop filter (VP9B_LF_WARP=1)
  const bool lf_warp = sizeof(Px) == 1 && B.f[0].ss_x != 0 && B.f[0].ss_y != 0 && ctx->lf_warp;
  // tickets: every participating frame's counter advances by the number of CTAs launched for it
  for (int i = 0; i < n; i++) {
    FrameParams &P = B.f[i];
    P.ticket_intra_base = ctxs[i]->ticket_intra_base;
    P.ticket_lf_base = ctxs[i]->ticket_lf_base;
    if (P.do_intra) ctxs[i]->ticket_intra_base += (uint32_t)g_intra;
    if (P.do_lf) ctxs[i]->ticket_lf_base += (uint32_t)g_lf;
  }
  if (prof) cudaEventRecord(ctx->prof_ev[0], st);
  if (g_expand) {
    k_expand_blocks<FB><<<dim3(g_expand, 1, n), 128, 0, st>>>(B);
    ctx->launches++;
  }
  if (prof) cudaEventRecord(ctx->prof_ev[1], st);
  if (g_mcx && ctx->mc_flat && B.f[0].ss_x == B.f[0].ss_y) {
    int gx = 0, gy = 0;
    for (int i = 0; i < n; i++) {
      const FrameParams &P = B.f[i];
      if (!P.do_inter) continue;
      const int pwy = P.mi_cols * 8, phy = P.mi_rows * 8, phc = phy >> P.ss_y;
      gx = std::max(gx, (pwy + 31) / 32);
      gy = std::max(gy, (phy + 127) / 128 + 2 * (P.ss_x ? (phc + 63) / 64 : (phc + 127) / 128));
    }
    k_inter_pred_flat<Px, FB><<<dim3(gx, gy, n), 128, 0, st>>>(B);
    ctx->launches++;
  } else if (g_mcx) {
    const dim3 grid(g_mcx, g_mcy, n);   // one CTA per superblock
    const size_t smem = McStage<Px>::kBudget * sizeof(Px);
    const unsigned abit = (sizeof(Px) == 1 ? 64u : 128u) << abase;
    if (!(ctx->attr_done & abit)) {
      CK(cudaFuncSetAttribute(k_inter_pred_sb<Px, FB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      ctx->attr_done |= abit;
    }
    k_inter_pred_sb<Px, FB><<<grid, 128, smem, st>>>(B);
    ctx->launches++;
  }
  if (prof) cudaEventRecord(ctx->prof_ev[2], st);
  if (g_res) {
    k_residual<Px, PT, FB><<<dim3(g_res, 1, n), kResWarps * 32, 0, st>>>(B);
    ctx->launches++;
  }
  if (prof) cudaEventRecord(ctx->prof_ev[3], st);
  if (g_intra) {
    const size_t ismem = IntraTileGeom<Px>::kWarpBytes * kWaveWarps;
    // opt-in shared memory size: per device, so remembered per context (one context = one device)
    const bool dense = sizeof(Px) == 1 && (g_intra * n > 2 * ctx->num_sms || VP9B_INTRA_MINB > 1);
    const unsigned abit = (sizeof(Px) == 1 ? 1u : 2u) << abase;
    if (!(ctx->attr_done & abit)) {
      CK(cudaFuncSetAttribute((k_intra<Px, PT, FB, 1>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ismem));
      CK(cudaFuncSetAttribute((k_intra<Px, PT, FB, VP9B_INTRA_DENSE>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ismem));
      ctx->attr_done |= abit;
    }
    k_intra_residual<Px, PT, FB><<<dim3(g_ires, 1, n), kResWarps * 32, 0, st>>>(B);
    ctx->launches++;
    if (prof) cudaEventRecord(ctx->prof_ev[4], st);
    if (dense) k_intra<Px, PT, FB, VP9B_INTRA_DENSE><<<dim3(g_intra, 1, n), kWaveWarps * 32, ismem, st>>>(B);
    else k_intra<Px, PT, FB, 1><<<dim3(g_intra, 1, n), kWaveWarps * 32, ismem, st>>>(B);
    ctx->launches++;
  } else if (prof) {
    cudaEventRecord(ctx->prof_ev[4], st);
  }
  if (prof) cudaEventRecord(ctx->prof_ev[5], st);
  if (g_lf) {
    // one persistent CTA per superblock column of the widest frame in the batch; threads = sample rows of the three planes
    int max_rows = 0;
    for (int i = 0; i < n; i++) max_rows = std::max(max_rows, B.f[i].sb_rows);
    const size_t plan_bytes = ((size_t)max_rows * 49 + 15) & ~(size_t)15;   // per superblock: 24 dependent sets + chain flag
    // opt-in shared-memory limit, once per device and kernel variant: two tiles + the plan of the tallest frame (16384 / 64 rows)
    // 4:2:2 / 4:4:0: the luma plane goes through the 4:2:0 kernel (its chroma masks are empty and its
    // 32x32 chroma tiles stay inside the planes); the chroma planes follow in k_lf_uv_non420
    const bool sub = B.f[0].ss_x != 0 || B.f[0].ss_y != 0;
    const size_t tiles = 2 * (sub ? LfGeom<Px, true>::kTileBytes : LfGeom<Px, false>::kTileBytes);
    const unsigned abit = ((sizeof(Px) == 1 ? 4u : 8u) << (sub ? 2 : 0)) << abase;
    if (!(ctx->attr_done & abit)) {
      const int lim = (int)(tiles + 256 * 49 + 16);
      if (sub) CK(cudaFuncSetAttribute((k_lf_sb<Px, true, FB>), cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
      else CK(cudaFuncSetAttribute((k_lf_sb<Px, false, FB>), cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
      ctx->attr_done |= abit;
    }
    if (lf_warp) {
      k_lf_warp<FB><<<dim3(g_lf, 1, n), 32, LfwGeom::kTileBytes + plan_bytes, st>>>(B);
    } else {
      // programmatic dependent launch behind k_intra (see there): the column workers start while the
      // intra wavefront is still running and follow it through prog_done.  Not while profiling (the
      // stage events would sit between the two kernels) and only directly behind k_intra.
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(g_lf, 1, n);
      cfg.blockDim = dim3(sub ? LfGeom<Px, true>::NT : LfGeom<Px, false>::NT, 1, 1);
      cfg.dynamicSmemBytes = tiles + plan_bytes;
      cfg.stream = st;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      at[0].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = at;
      cfg.numAttrs = (g_intra && !prof && ctx->lf_overlap) ? 1 : 0;
      if (sub) CK(cudaLaunchKernelEx(&cfg, k_lf_sb<Px, true, FB>, B));
      else CK(cudaLaunchKernelEx(&cfg, k_lf_sb<Px, false, FB>, B));
    }
    ctx->launches++;
    if (B.f[0].ss_x != B.f[0].ss_y) {   // 4:2:2 / 4:4:0: the chroma planes, after the luma pass has finished
      k_lf_uv_non420<Px, FB><<<dim3(2, 1, n), 64, 0, st>>>(B);
      ctx->launches++;
    }
  }
  if (prof) {
    cudaEventRecord(ctx->prof_ev[6], st);
    cudaEventSynchronize(ctx->prof_ev[6]);
    for (int i = 0; i < 6; i++) {
      float ms = 0;
      cudaEventElapsedTime(&ms, ctx->prof_ev[i], ctx->prof_ev[i + 1]);
      ctx->stage_ms[i] += ms;
    }
  }
  CK(cudaGetLastError());
  return 0;
}

// Header of a packed buffer: magic, slots, sampling, section bounds, geometry.
static int check_header(vp9gpu_ctx *ctx, const vp9b_frame_hdr *h) {
  if (h->magic != VP9B_MAGIC) { ctx->err = "bad packed buffer magic"; return VP9B_ERR_INVALID; }
  if (h->dst_slot < 0 || h->dst_slot >= VP9B_NUM_SLOTS) { ctx->err = "bad dst slot"; return VP9B_ERR_INVALID; }
  if (h->ss_x > 1 || h->ss_y > 1) { ctx->err = "bad chroma subsampling"; return VP9B_ERR_INVALID; }
  {
    // the header is the only part of a packed buffer the host can check cheaply: every section
    // must lie inside the buffer and the per-superblock tables must have the frame's geometry
    static const uint32_t elem[VP9B_SEC_COUNT] = { sizeof(vp9b_block), sizeof(vp9b_submv), sizeof(vp9b_txb), sizeof(vp9b_txb),
                                                   sizeof(vp9b_txb), sizeof(vp9b_txb), sizeof(vp9b_sbrange), 0,
                                                   sizeof(vp9b_lfm), sizeof(vp9b_lfm_uv), 0, 0 };
    const uint32_t coef_elem = h->bit_depth > 8 ? 4 : 2;
    for (int i = 0; i < VP9B_SEC_COUNT; i++) {
      const uint64_t bytes = (uint64_t)h->sec[i].count * (i == VP9B_SEC_COEF ? coef_elem : elem[i]);
      if ((h->sec[i].offset & 15) || (uint64_t)h->sec[i].offset + bytes > h->total_bytes) {
        ctx->err = "packed buffer: section out of bounds";
        return VP9B_ERR_INVALID;
      }
    }
    const uint32_t nsb = h->sb_cols * h->sb_rows;
    if (h->sb_cols != (h->mi_cols + 7) / 8 || h->sb_rows != (h->mi_rows + 7) / 8 || h->mi_cols != (h->width + 7) / 8 ||
        h->mi_rows != (h->height + 7) / 8 || h->sec[VP9B_SEC_SB_INTRA].count != 3 * nsb ||
        (h->lf_level && h->sec[VP9B_SEC_LFM].count != nsb) ||
        (h->lf_level && h->ss_x != h->ss_y && h->sec[VP9B_SEC_LFM_UV].count != nsb) || (h->bit_depth != 8 && h->bit_depth != 10 && h->bit_depth != 12)) {
      ctx->err = "packed buffer: inconsistent frame geometry";
      return VP9B_ERR_INVALID;
    }
  }
  return 0;
}

// Record-level validation of a packed buffer that lives in host memory.  The header checks are in
// prepare_frame(); this pass makes sure that no record can make a kernel index outside the frame,
// the coefficient stream, the sub-8x8 MV table or a shared-memory tile.  It runs for every buffer
// that enters through vp9gpu_frame_submit / vp9gpu_upload_packed (a host-side packer is a foreign
// producer); the whole-decoder layer, whose buffers come from vp9b_parser_* in this library, turns
// it off (vp9gpu_set_validation).
static int validate_records(const uint8_t *p, const vp9b_frame_hdr *h, std::string &err) {
  auto bad = [&](const char *m) { err = m; return VP9B_ERR_INVALID; };
  const uint32_t n_submv = h->sec[VP9B_SEC_SUBMV].count, n_coef = h->sec[VP9B_SEC_COEF].count;
  const vp9b_block *blk = (const vp9b_block *)(p + h->sec[VP9B_SEC_BLOCKS].offset);
  for (uint32_t i = 0; i < h->sec[VP9B_SEC_BLOCKS].count; i++) {
    const vp9b_block &b = blk[i];
    if (b.bsize > 12 || b.mi_col >= h->mi_cols || b.mi_row >= h->mi_rows || b.ref0 > 3 || b.ref1 > 3 || b.filter > 3)
      return bad("packed buffer: block record out of range");
    if (b.ref0 && h->intra_only) return bad("packed buffer: inter block in an intra-only frame");
    if (b.ref1 && !b.ref0) return bad("packed buffer: second reference without a first");
    if (b.ref0 && b.bsize < 3) {
      uint32_t idx;
      memcpy(&idx, &b.mv[0][0], 4);
      if (idx >= n_submv) return bad("packed buffer: sub-8x8 motion vector index out of range");
    }
  }
  const uint32_t nsb = h->sb_cols * h->sb_rows;
  auto check_txb = [&](const vp9b_txb &t, int want_plane, bool intra) -> bool {
    const int tx = t.tx & 3, plane = (t.tx >> 4) & 3, bs4 = 1 << tx;
    if (plane > 2 || (want_plane >= 0 && plane != want_plane)) return false;
    if (t.eob > (16u << (2 * tx)) || (uint64_t)t.coef_off + t.eob > n_coef) return false;
    if (h->lossless && tx != 0) return false;
    // a transform block may hang over the aligned frame edge into the 64-sample padding
    const uint32_t pw4 = ((h->mi_cols * 8) >> (plane ? h->ss_x : 0)) / 4 + 16, ph4 = ((h->mi_rows * 8) >> (plane ? h->ss_y : 0)) / 4 + 16;
    if ((uint32_t)t.x4 + bs4 > pw4 || (uint32_t)t.y4 + bs4 > ph4) return false;
    if (intra && (t.aux & 15) > 9) return false;
    return true;
  };
  const vp9b_txb *ti = (const vp9b_txb *)(p + h->sec[VP9B_SEC_TXB_INTER].offset);
  uint32_t first = 0;
  for (int sz = 0; sz < 4; sz++) {     // the list is grouped by transform size: a warp relies on it
    for (uint32_t i = first; i < first + h->n_txb_inter_size[sz] && i < h->sec[VP9B_SEC_TXB_INTER].count; i++)
      if ((ti[i].tx & 3) != sz || !check_txb(ti[i], -1, false) || !ti[i].eob) return bad("packed buffer: inter transform record out of range");
    first += h->n_txb_inter_size[sz];
  }
  const vp9b_sbrange *rg = (const vp9b_sbrange *)(p + h->sec[VP9B_SEC_SB_INTRA].offset);
  for (int pl = 0; pl < 3; pl++) {
    const vp9b_txb *t = (const vp9b_txb *)(p + h->sec[VP9B_SEC_TXB_INTRA_Y + pl].offset);
    const uint32_t cnt = h->sec[VP9B_SEC_TXB_INTRA_Y + pl].count;
    const int n4 = (pl && h->ss_x) ? 8 : 16, n4y = (pl && h->ss_y) ? 8 : 16;       // superblock size in 4-sample units
    for (uint32_t sb = 0; sb < nsb; sb++) {
      const vp9b_sbrange &r = rg[(size_t)pl * nsb + sb];
      if ((uint64_t)r.start + r.count > cnt) return bad("packed buffer: intra range out of bounds");
      const uint32_t sbr = sb / h->sb_cols, sbc = sb % h->sb_cols;
      for (uint32_t i = r.start; i < r.start + r.count; i++) {
        const int bs4 = 1 << (t[i].tx & 3);
        // the wavefront kernel reconstructs a superblock inside a shared-memory tile: every record of
        // the range has to lie inside that superblock
        if (!check_txb(t[i], pl, true) || t[i].x4 < sbc * n4 || t[i].x4 + bs4 > (sbc + 1) * n4 || t[i].y4 < sbr * n4y ||
            t[i].y4 + bs4 > (sbr + 1) * n4y)
          return bad("packed buffer: intra transform record out of range");
      }
    }
  }
  if (h->lf_level) {
    const vp9b_lfm *m = (const vp9b_lfm *)(p + h->sec[VP9B_SEC_LFM].offset);
    for (uint32_t sb = 0; sb < nsb; sb++)
      for (int i = 0; i < 64; i++)
        if (m[sb].lfl_y[i] > 63) return bad("packed buffer: loop-filter level out of range");
    if (h->ss_x != h->ss_y) {
      const vp9b_lfm_uv *mu = (const vp9b_lfm_uv *)(p + h->sec[VP9B_SEC_LFM_UV].offset);
      for (uint32_t sb = 0; sb < nsb; sb++) {
        for (int i = 0; i < 64; i++)
          if (mu[sb].lfl[i] > 63) return bad("packed buffer: loop-filter level out of range");
        // 4:2:2 / 4:4:0 chroma is filtered by its own kernel: the 4:2:0 chroma masks must be empty
        for (int i = 0; i < 4; i++)
          if (m[sb].left_uv[i] | m[sb].above_uv[i] | m[sb].int_4x4_uv) return bad("packed buffer: 4:2:0 chroma masks in a non-4:2:0 frame");
      }
    }
  }
  return 0;
}

// Validate one packed frame, make sure its context has the memory it needs and fill its FrameParams.
static int prepare_frame(vp9gpu_ctx *ctx, const uint8_t *d_packed, const vp9b_frame_hdr *h, FrameParams &P) {
  {
    const int he = check_header(ctx, h);
    if (he) return he;
  }
  CK(cudaSetDevice(ctx->device));
  int e = 0;
  if (!ctx->slot[h->dst_slot].base) {
    // first use: size the whole pool now so that no allocation happens in the steady state
    for (int s2 = 0; s2 < VP9B_NUM_SLOTS && !e; s2++)
      if (!ctx->slot[s2].base) {
        e = setup_slot(ctx, s2, h->width, h->height, h->bit_depth, h->ss_x, h->ss_y);
        ctx->slot[s2].valid = false;
      }
    if (e) return e;
  }
  e = setup_slot(ctx, h->dst_slot, h->width, h->height, h->bit_depth, h->ss_x, h->ss_y);
  if (e) return e;
  e = setup_workspace(ctx, h);
  if (e) return e;
  memset(&P, 0, sizeof(P));
  const DevFrame &D = ctx->slot[h->dst_slot];
  for (int p = 0; p < 3; p++) P.dst[p] = D.pl[p];
  if (!h->intra_only) {
    for (int r = 0; r < 3; r++) {
      const vp9b_ref_desc &rd = h->ref[r];
      if (rd.slot < 0 || rd.slot >= VP9B_NUM_SLOTS || !ctx->slot[rd.slot].valid) {
        ctx->err = "reference slot not populated";
        return VP9B_ERR_INVALID;
      }
      const DevFrame &F = ctx->slot[rd.slot];
      if (F.bd != h->bit_depth || F.ss_x != h->ss_x || F.ss_y != h->ss_y) {
        ctx->err = "reference format mismatch";
        return VP9B_ERR_INVALID;
      }
      // the clamp extent of motion compensation must be the slot's real size, and a frame cannot
      // predict from the buffer it is being written to
      if (rd.width != F.width || rd.height != F.height || rd.slot == h->dst_slot) {
        ctx->err = "reference descriptor does not match its frame slot";
        return VP9B_ERR_INVALID;
      }
      if (rd.x_scale_fp >= 0 && (rd.x_step_q4 < 1 || rd.x_step_q4 > 80 || rd.y_step_q4 < 1 || rd.y_step_q4 > 80)) {
        ctx->err = "reference scale out of range";
        return VP9B_ERR_INVALID;
      }
      P.ref[r].slot = rd.slot;
      for (int p = 0; p < 3; p++) {
        P.ref[r].pl[p] = F.pl[p];
        P.ref[r].w[p] = p ? (int)((rd.width + h->ss_x) >> h->ss_x) : (int)rd.width;
        P.ref[r].h[p] = p ? (int)((rd.height + h->ss_y) >> h->ss_y) : (int)rd.height;
      }
      P.ref[r].x_scale_fp = rd.x_scale_fp; P.ref[r].y_scale_fp = rd.y_scale_fp;
      P.ref[r].x_step_q4 = rd.x_step_q4; P.ref[r].y_step_q4 = rd.y_step_q4;
    }
  }
  P.mi_cols = h->mi_cols; P.mi_rows = h->mi_rows; P.sb_cols = h->sb_cols; P.sb_rows = h->sb_rows;
  P.ss_x = h->ss_x; P.ss_y = h->ss_y; P.bd = h->bit_depth; P.lossless = h->lossless;
  P.blocks = (const vp9b_block *)(d_packed + h->sec[VP9B_SEC_BLOCKS].offset);
  P.n_blocks = h->sec[VP9B_SEC_BLOCKS].count;
  P.submv = (const vp9b_submv *)(d_packed + h->sec[VP9B_SEC_SUBMV].offset);
  P.txb_inter = (const vp9b_txb *)(d_packed + h->sec[VP9B_SEC_TXB_INTER].offset);
  P.n_txb_inter = h->sec[VP9B_SEC_TXB_INTER].count;
  for (int i = 0; i < 4; i++) P.n_txb_size[i] = (int)h->n_txb_inter_size[i];
  if (h->n_txb_inter_size[0] + h->n_txb_inter_size[1] + h->n_txb_inter_size[2] + h->n_txb_inter_size[3] != (uint32_t)P.n_txb_inter) {
    ctx->err = "inter transform block counts do not add up";
    return VP9B_ERR_INVALID;
  }
  for (int p = 0; p < 3; p++) {
    P.txb_intra[p] = (const vp9b_txb *)(d_packed + h->sec[VP9B_SEC_TXB_INTRA_Y + p].offset);
    P.n_txb_intra[p] = (int)h->sec[VP9B_SEC_TXB_INTRA_Y + p].count;
  }
  if (P.n_txb_intra[0] + P.n_txb_intra[1] + P.n_txb_intra[2]) {
    // residual planes: same sample stride and rows as the destination planes
    const size_t rsz = h->bit_depth > 8 ? 4 : 2;
    size_t off[3], total = 0;
    for (int p = 0; p < 3; p++) {
      const int rows = (int)((((h->height + 7) & ~7u) >> (p ? h->ss_y : 0)) + 64);
      off[p] = total;
      total += ((size_t)D.pl[p].stride * rows * rsz + 255) & ~(size_t)255;
    }
    if (ctx->resid_bytes < total) {
      CK(cudaStreamSynchronize(ctx->stream));
      cudaFree(ctx->d_resid);
      ctx->d_resid = nullptr;
      ctx->resid_bytes = 0;
      CK(cudaMalloc(&ctx->d_resid, total));
      ctx->resid_bytes = total;
    }
    for (int p = 0; p < 3; p++) P.resid[p] = ctx->d_resid + off[p];
  }
  P.sb_intra = (const vp9b_sbrange *)(d_packed + h->sec[VP9B_SEC_SB_INTRA].offset);
  P.coef = d_packed + h->sec[VP9B_SEC_COEF].offset;
  P.lfm = (const vp9b_lfm *)(d_packed + h->sec[VP9B_SEC_LFM].offset);
  P.lfm_uv = (h->ss_x != h->ss_y) ? (const vp9b_lfm_uv *)(d_packed + h->sec[VP9B_SEC_LFM_UV].offset) : nullptr;
  P.mi = ctx->d_mi;
  P.tmaps = ctx->tma_ok ? ctx->d_tmaps : nullptr;
  // frames with a scaled reference do not stage windows: their units use the staging buffer as scratch
  for (int r = 0; r < 3 && !h->intra_only; r++)
    if (P.ref[r].x_step_q4 != 16 || P.ref[r].y_step_q4 != 16) P.tmaps = nullptr;
  if (h->ss_x != h->ss_y) P.tmaps = nullptr;   // 4:2:2 / 4:4:0: chroma windows are not square, no staging
  // progress counters: [0, 3*1024) intra (plane, row), [3072, 4096) loop filter.  Values are
  // prog_base + count with prog_base advancing by 2048 per frame; cleared only when it wraps.
  if (h->width > 16384 || h->height > 16384) { ctx->err = "frame too large"; return VP9B_ERR_UNSUP; }
  if (++ctx->epoch == 0) ctx->epoch = 1;
  if ((ctx->epoch & 0xfffff) == 0) CK(cudaMemsetAsync(ctx->d_tickets, 0, sizeof(uint32_t) * 4096, ctx->stream));
  P.prog_intra = ctx->d_tickets;
  P.prog_done = ctx->d_tickets + 3072;
  P.prog_base = (ctx->epoch & 0xfffff) << 11;
  P.ticket_lf = ctx->d_tickets + 4096;
  P.lf_flags = ctx->d_lf_flags;
  P.lf_epoch = ctx->epoch;   // never 0: the flags start out as 0
  P.ticket_intra = ctx->d_tickets + 4097;
  memcpy(P.lf_lim, h->lf_lim, 64);
  memcpy(P.lf_mblim, h->lf_mblim, 64);
  P.do_inter = !h->intra_only;
  P.do_lf = h->lf_level != 0;
  return 0;
}

static int submit_device(vp9gpu_ctx *ctx, const uint8_t *d_packed, const vp9b_frame_hdr *h) {
  FrameBatchSmall B;
  int e = prepare_frame(ctx, d_packed, h, B.f[0]);
  if (e) return e;
  vp9gpu_ctx *one[1] = { ctx };
  if (h->bit_depth == 8) return launch_batch<uint8_t, int>(one, B, 1);
  return launch_batch<uint16_t, long long>(one, B, 1);
}

template <typename FB>
static int submit_multi_impl(vp9gpu_ctx *const *ctxs, const void *const *d_packed, const vp9b_frame_hdr *const *hdr, int n) {
  vp9gpu_ctx *lead = ctxs[0];
  FB B;
  for (int i = 0; i < n; i++) {
    const int e = prepare_frame(ctxs[i], (const uint8_t *)d_packed[i], hdr[i], B.f[i]);
    if (e) { lead->err = ctxs[i]->err; return e; }
  }
  CK2(lead, cudaSetDevice(lead->device));
  for (int i = 1; i < n; i++) {   // join: the lead stream continues after everything queued on the others
    if (!ctxs[i]->join_ev) CK2(lead, cudaEventCreateWithFlags(&ctxs[i]->join_ev, cudaEventDisableTiming));
    CK2(lead, cudaEventRecord(ctxs[i]->join_ev, ctxs[i]->stream));
    CK2(lead, cudaStreamWaitEvent(lead->stream, ctxs[i]->join_ev, 0));
  }
  int e;
  if (hdr[0]->bit_depth == 8) e = launch_batch<uint8_t, int>(ctxs, B, n);
  else e = launch_batch<uint16_t, long long>(ctxs, B, n);
  if (e) return e;
  for (int i = 1; i < n; i++) {   // fork: later work on the other streams waits for this batch
    CK2(lead, cudaEventRecord(ctxs[i]->join_ev, lead->stream));
    CK2(lead, cudaStreamWaitEvent(ctxs[i]->stream, ctxs[i]->join_ev, 0));
  }
  return 0;
}

// Frames of n different contexts through one launch per stage.  They run on the stream of the first
// context; the other contexts' streams are joined before and after, so their earlier and later
// work stays ordered.  Frames that cannot share a launch (other sample type / chroma format) are
// submitted one by one.
extern "C" int vp9gpu_frame_submit_device_multi(vp9gpu_ctx *const *ctxs, const void *const *d_packed,
                                                const vp9b_frame_hdr *const *hdr, int n) {
  if (!ctxs || !d_packed || !hdr || n <= 0) return VP9B_ERR_INVALID;
  for (int i = 0; i < n; i++)
    if (!ctxs[i] || !d_packed[i] || !hdr[i]) return VP9B_ERR_INVALID;
  bool same = n <= kMaxBatch;
  for (int i = 1; i < n && same; i++) {
    same = (hdr[i]->bit_depth > 8) == (hdr[0]->bit_depth > 8) && hdr[i]->ss_x == hdr[0]->ss_x && hdr[i]->ss_y == hdr[0]->ss_y &&
           ctxs[i]->device == ctxs[0]->device;
    for (int j = 0; j < i; j++) same = same && ctxs[i] != ctxs[j];
  }
  if (!same || n == 1) {
    for (int i = 0; i < n; i++) {
      const int e = submit_device(ctxs[i], (const uint8_t *)d_packed[i], hdr[i]);
      if (e) return e;
    }
    return 0;
  }
  if (n <= FrameBatchSmall::kN) return submit_multi_impl<FrameBatchSmall>(ctxs, d_packed, hdr, n);
  return submit_multi_impl<FrameBatch>(ctxs, d_packed, hdr, n);
}

extern "C" int vp9gpu_frame_submit_device(vp9gpu_ctx *ctx, const void *d_packed, const vp9b_frame_hdr *hdr) {
  if (!ctx || !d_packed || !hdr) return VP9B_ERR_INVALID;
  return submit_device(ctx, (const uint8_t *)d_packed, hdr);
}

extern "C" int vp9gpu_frame_submit_device_n(vp9gpu_ctx *ctx, const void *const *d_packed,
                                            const vp9b_frame_hdr *const *hdr, int n) {
  if (!ctx || !d_packed || !hdr || n < 0) return VP9B_ERR_INVALID;
  for (int i = 0; i < n; i++) {
    if (!d_packed[i] || !hdr[i]) return VP9B_ERR_INVALID;
    const int e = submit_device(ctx, (const uint8_t *)d_packed[i], hdr[i]);
    if (e) return e;
  }
  return 0;
}

// next entry of the upload ring, at least `bytes` large and no longer in use by the GPU
static int acquire_upload(vp9gpu_ctx *ctx, size_t bytes) {
  CK(cudaSetDevice(ctx->device));
  const int k = ctx->ring_pos;
  CK(cudaEventSynchronize(ctx->ring_ev[k]));  // previous user of this ring entry is done
  if (ctx->ring_bytes[k] < bytes) {
    // grow every ring entry together (allocation synchronises the device: do it rarely)
    const size_t nb = std::max((bytes * 3 / 2 + (1 << 20)) & ~(size_t)4095, (size_t)4 << 20);
    CK(cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < vp9gpu_ctx::kRing; i++) {
      cudaFreeHost(ctx->h_pinned[i]);
      cudaFree(ctx->d_packed[i]);
      ctx->h_pinned[i] = ctx->d_packed[i] = nullptr;
      ctx->ring_bytes[i] = 0;
      CK(cudaMallocHost(&ctx->h_pinned[i], nb));
      CK(cudaMalloc(&ctx->d_packed[i], nb));
      ctx->ring_bytes[i] = nb;
    }
  }
  return 0;
}

// upload + reconstruct the frame that sits in the current ring entry
static int submit_ring_entry(vp9gpu_ctx *ctx, size_t bytes) {
  NvtxRange nvtx_range("vp9b: upload packed frame + submit");
  const int k = ctx->ring_pos;
  ctx->ring_pos = (k + 1) % vp9gpu_ctx::kRing;
  CK(cudaMemcpyAsync(ctx->d_packed[k], ctx->h_pinned[k], bytes, cudaMemcpyHostToDevice, ctx->stream));
  const int e = submit_device(ctx, (const uint8_t *)ctx->d_packed[k], (const vp9b_frame_hdr *)ctx->h_pinned[k]);
  cudaEventRecord(ctx->ring_ev[k], ctx->stream);
  return e;
}

static int check_host_packed(vp9gpu_ctx *ctx, const void *packed, size_t bytes) {
  const vp9b_frame_hdr *h = (const vp9b_frame_hdr *)packed;
  if (h->magic != VP9B_MAGIC || h->total_bytes > bytes) { ctx->err = "corrupt packed buffer"; return VP9B_ERR_INVALID; }
  if (ctx->validate) {
    int e = check_header(ctx, h);
    if (!e) e = validate_records((const uint8_t *)packed, h, ctx->err);
    if (e) return e;
  }
  return 0;
}

extern "C" int vp9gpu_frame_submit(vp9gpu_ctx *ctx, const void *packed, size_t bytes) {
  if (!ctx || !packed || bytes < sizeof(vp9b_frame_hdr)) return VP9B_ERR_INVALID;
  int e = check_host_packed(ctx, packed, bytes);
  if (e) return e;
  const vp9b_frame_hdr *h = (const vp9b_frame_hdr *)packed;
  if (packed == ctx->h_pinned[ctx->ring_pos]) return submit_ring_entry(ctx, h->total_bytes);   // built in place
  e = acquire_upload(ctx, bytes);
  if (e) return e;
  memcpy(ctx->h_pinned[ctx->ring_pos], packed, h->total_bytes);
  return submit_ring_entry(ctx, h->total_bytes);
}

extern "C" void *vp9gpu_upload_buffer(vp9gpu_ctx *ctx, size_t bytes) {
  if (!ctx || acquire_upload(ctx, bytes)) return nullptr;
  return ctx->h_pinned[ctx->ring_pos];
}

extern "C" int vp9gpu_frame_submit_staged(vp9gpu_ctx *ctx, const void *packed, size_t bytes) {
  if (!ctx || !packed || bytes < sizeof(vp9b_frame_hdr)) return VP9B_ERR_INVALID;
  if (packed != ctx->h_pinned[ctx->ring_pos] || bytes > ctx->ring_bytes[ctx->ring_pos]) {
    ctx->err = "vp9gpu_frame_submit_staged: not the buffer handed out by vp9gpu_upload_buffer";
    return VP9B_ERR_INVALID;
  }
  const int e = check_host_packed(ctx, packed, bytes);
  if (e) return e;
  CK(cudaSetDevice(ctx->device));
  return submit_ring_entry(ctx, ((const vp9b_frame_hdr *)packed)->total_bytes);
}

extern "C" int vp9gpu_upload_packed(vp9gpu_ctx *ctx, const void *packed, size_t bytes, void **d_packed) {
  if (!ctx || !packed || !d_packed || bytes < sizeof(vp9b_frame_hdr)) return VP9B_ERR_INVALID;
  if (ctx->validate) {
    const vp9b_frame_hdr *h = (const vp9b_frame_hdr *)packed;
    if (h->magic != VP9B_MAGIC || h->total_bytes > bytes) { ctx->err = "corrupt packed buffer"; return VP9B_ERR_INVALID; }
    int e = check_header(ctx, h);
    if (!e) e = validate_records((const uint8_t *)packed, h, ctx->err);
    if (e) return e;
  }
  CK(cudaSetDevice(ctx->device));
  CK(cudaMalloc(d_packed, bytes));
  CK(cudaMemcpy(*d_packed, packed, bytes, cudaMemcpyHostToDevice));
  return 0;
}

extern "C" int vp9gpu_free_packed(vp9gpu_ctx *ctx, void *d_packed) {
  if (!ctx) return VP9B_ERR_INVALID;
  CK(cudaSetDevice(ctx->device));
  CK(cudaStreamSynchronize(ctx->stream));
  CK(cudaFree(d_packed));
  return 0;
}

extern "C" void *vp9gpu_alloc_host(size_t bytes) {
  void *p = nullptr;
  if (cudaMallocHost(&p, bytes) != cudaSuccess) return nullptr;
  return p;
}
extern "C" void vp9gpu_free_host(void *p) { cudaFreeHost(p); }

static int get_frame_impl(vp9gpu_ctx *ctx, int slot, uint8_t *y, uint8_t *u, uint8_t *v, int y_stride, int uv_stride,
                          bool sync) {
  if (!ctx || slot < 0 || slot >= VP9B_NUM_SLOTS || !ctx->slot[slot].valid) return VP9B_ERR_INVALID;
  CK(cudaSetDevice(ctx->device));
  const DevFrame &f = ctx->slot[slot];
  uint8_t *out[3] = { y, u, v };
  for (int p = 0; p < 3; p++) {
    const int w = p ? (int)((f.width + f.ss_x) >> f.ss_x) : (int)f.width;
    const int hh = p ? (int)((f.height + f.ss_y) >> f.ss_y) : (int)f.height;
    CK(cudaMemcpy2DAsync(out[p], p ? uv_stride : y_stride, f.pl[p].ptr, (size_t)f.pl[p].stride * f.bps,
                         (size_t)w * f.bps, hh, cudaMemcpyDeviceToHost, ctx->stream));
  }
  if (sync) CK(cudaStreamSynchronize(ctx->stream));
  return 0;
}

extern "C" int vp9gpu_get_frame(vp9gpu_ctx *ctx, int slot, uint8_t *y, uint8_t *u, uint8_t *v, int y_stride,
                                int uv_stride) {
  return get_frame_impl(ctx, slot, y, u, v, y_stride, uv_stride, true);
}
extern "C" int vp9gpu_get_frame_async(vp9gpu_ctx *ctx, int slot, uint8_t *y, uint8_t *u, uint8_t *v, int y_stride,
                                      int uv_stride) {
  return get_frame_impl(ctx, slot, y, u, v, y_stride, uv_stride, false);
}

extern "C" int vp9gpu_set_frame(vp9gpu_ctx *ctx, int slot, uint32_t width, uint32_t height, int bit_depth, int ss_x,
                                int ss_y, const uint8_t *y, const uint8_t *u, const uint8_t *v, int y_stride,
                                int uv_stride) {
  if (!ctx || slot < 0 || slot >= VP9B_NUM_SLOTS) return VP9B_ERR_INVALID;
  CK(cudaSetDevice(ctx->device));
  int e = setup_slot(ctx, slot, width, height, bit_depth, ss_x, ss_y);
  if (e) return e;
  const DevFrame &f = ctx->slot[slot];
  const uint8_t *in[3] = { y, u, v };
  for (int p = 0; p < 3; p++) {
    const int w = p ? (int)((width + ss_x) >> ss_x) : (int)width;
    const int hh = p ? (int)((height + ss_y) >> ss_y) : (int)height;
    CK(cudaMemcpy2DAsync(f.pl[p].ptr, (size_t)f.pl[p].stride * f.bps, in[p], p ? uv_stride : y_stride,
                         (size_t)w * f.bps, hh, cudaMemcpyHostToDevice, ctx->stream));
  }
  CK(cudaStreamSynchronize(ctx->stream));
  return 0;
}
