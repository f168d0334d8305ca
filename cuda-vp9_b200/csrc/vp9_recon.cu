// VP9 reconstruction kernels for sm_100a (B200) + the vp9gpu_* C ABI (include/vp9_b200.h).
//
// One packed buffer per frame is uploaded; reference frames stay resident in HBM.  Stages per
// frame, all enqueued on the context stream:
//   0  k_expand_blocks   block records -> dense per-8x8 motion map (L2 resident)
//   1  k_inter_pred      8-tap sub-pel MC incl. compound + scaled refs   (a1-a3)
//   2  k_residual        inverse transform + add for inter blocks          (a6-a7)
//   3  k_intra           intra prediction + residual, superblock wavefront (a4-a7)
//   4  k_loopfilter      lpf 4/8/16 in libvpx's normative superblock order (a9-a10)
// Integer byte work, HBM/L2 bound -- no tensor cores by design (bit exactness).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <mutex>
#include <string>
#include <vector>

#include "vp9_b200.h"
#include "vp9_tables.h"
#include "vp9_txfm.cuh"

namespace vp9b {

// ------------------------------------------------------------------------------------------
// device-side constant tables
__device__ int16_t d_scan[(16 + 64 + 256) * 3 + 1024];  // [4x4: def,row,col][8x8..][16x16..][32x32 def]
__device__ int8_t d_filters[4 * 16 * 8];                 // taps fit int8 except 128 (stored as int16 below)
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
  const vp9b_txb *txb_intra[3];
  const vp9b_sbrange *sb_intra;  // [3][nsb]
  const void *coef;
  const vp9b_lfm *lfm;
  MiEntry *mi;
  const uint32_t *order_intra;  // ticket -> sb index (anti-diagonal order r+c)
  const uint32_t *order_lf;     // ticket -> sb index (order 2r+c)
  uint32_t *flags_intra;        // [3][nsb] epoch when done
  uint32_t *flags_lf;           // [3][nsb]
  uint32_t *tickets;            // [2]
  uint32_t epoch;
  uint8_t lf_lim[64], lf_mblim[64];
};

__device__ __constant__ uint8_t d_bw8[13] = { 1, 1, 1, 1, 1, 2, 2, 2, 4, 4, 4, 8, 8 };
__device__ __constant__ uint8_t d_bh8[13] = { 1, 1, 1, 1, 2, 1, 2, 4, 2, 4, 8, 4, 8 };

VP9B_DEV int clampi(int v, int lo, int hi) { return min(max(v, lo), hi); }

// ------------------------------------------------------------------------------------------
// stage 0: block list -> dense mi map
__global__ void k_expand_blocks(const vp9b_block *__restrict__ blocks, int n, MiEntry *__restrict__ mi, int mi_cols,
                                int mi_rows) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
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
  for (int y = 0; y < bh; y++)
    for (int x = 0; x < bw; x++) mi[(size_t)(b.mi_row + y) * mi_cols + b.mi_col + x] = e;
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

// Generic (any block size / scaling / bit depth) per-sample predictor.  Used for scaled
// references, high bit depth and frame-edge tiles; the common 8-bit unscaled case goes through
// k_inter_pred_tile below.
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

template <typename Px>
__global__ void k_inter_pred_generic(const __grid_constant__ FrameParams P, int plane) {
  const int ss_x = plane ? P.ss_x : 0, ss_y = plane ? P.ss_y : 0;
  const int pw = (P.mi_cols * 8) >> ss_x, ph = (P.mi_rows * 8) >> ss_y;
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= pw || y >= ph) return;
  const int mi_col = (x << ss_x) >> 3, mi_row = (y << ss_y) >> 3;
  const MiEntry e = P.mi[(size_t)mi_row * P.mi_cols + mi_col];
  if (!e.ref0) return;
  const int bw8 = d_bw8[e.bsize], bh8 = d_bh8[e.bsize];
  const int mi_col0 = mi_col & ~(bw8 - 1), mi_row0 = mi_row & ~(bh8 - 1);
  // prediction unit inside the block: whole block, or one 4x4 for sub-8x8 blocks
  int xoff = 0, yoff = 0, k = 0;
  const int bx = x - ((mi_col0 * 8) >> ss_x), by = y - ((mi_row0 * 8) >> ss_y);
  const bool sub8 = e.bsize < 3;
  if (sub8) {
    xoff = bx & ~3;
    yoff = by & ~3;
    k = (yoff >> 2) * 2 + (xoff >> 2);  // only meaningful for unsubsampled planes
  }
  int val = 0;
#pragma unroll 1
  for (int r = 0; r < (e.ref1 ? 2 : 1); r++) {
    const RefGeom &R = P.ref[(r ? e.ref1 : e.ref0) - 1];
    int mvr, mvc;
    if (sub8) {
      const vp9b_submv &s = P.submv[e.aux];
      if (ss_x && ss_y) { mvr = s.chroma[r][0]; mvc = s.chroma[r][1]; }
      else { mvr = s.luma[k][r][0]; mvc = s.luma[k][r][1]; }
    } else {
      mvr = e.mv[r][0];
      mvc = e.mv[r][1];
    }
    const PredPos pp = pred_position(P, R, plane, mi_row0, mi_col0, bw8, bh8, xoff, yoff, mvr, mvc);
    const int v = predict_sample<Px>(P, R, plane, pp, bx - xoff, by - yoff, e.filter);
    val = r ? ((val + v + 1) >> 1) : v;
  }
  ((Px *)P.dst[plane].ptr)[(size_t)y * P.dst[plane].stride + x] = (Px)val;
}

// ------------------------------------------------------------------------------------------
// stages 2/3: inverse transform + add, one warp per transform block.  Rows then columns
// (vpx_idctNxN_add_c), final shift 4/5/6/6, clip to the pixel range.  tbuf: per-warp shared
// scratch of N*(N+1) ints.
template <typename Px, typename PT, int N>
VP9B_DEV void warp_itx_add(int *tbuf, const void *coef_base, int coef_bytes, const vp9b_txb &t, Px *dst, int stride,
                           int bd, int lossless, int lane) {
  constexpr int S = N + 1;
  const int tx_size = t.tx & 3, tx_type = lossless ? 0 : ((t.tx >> 2) & 3);
  const int eob = t.eob;
  for (int i = lane; i < N * S; i += 32) tbuf[i] = 0;
  __syncwarp();
  const int16_t *scan = d_scan + scan_offset(tx_size, tx_type);
  if (coef_bytes == 2) {
    const int16_t *c = (const int16_t *)coef_base + t.coef_off;
    for (int i = lane; i < eob; i += 32) {
      const int pos = scan[i];
      tbuf[(pos / N) * S + (pos % N)] = c[i];
    }
  } else {
    const int32_t *c = (const int32_t *)coef_base + t.coef_off;
    for (int i = lane; i < eob; i += 32) {
      const int pos = scan[i];
      tbuf[(pos / N) * S + (pos % N)] = c[i];
    }
  }
  __syncwarp();
  const int maxv = (1 << bd) - 1;
  if (lossless) {  // vpx_iwht4x4_16_add_c (inv_txfm.c:18-66)
    if (N == 4) {
      int o[4];
      if (lane < 4) {
        iwht4(tbuf[lane * S + 0] >> 2, tbuf[lane * S + 1] >> 2, tbuf[lane * S + 2] >> 2, tbuf[lane * S + 3] >> 2, o);
        for (int j = 0; j < 4; j++) tbuf[lane * S + j] = o[j];
      }
      __syncwarp();
      if (lane < 4) {
        iwht4(tbuf[0 * S + lane], tbuf[1 * S + lane], tbuf[2 * S + lane], tbuf[3 * S + lane], o);
        for (int j = 0; j < 4; j++) dst[j * stride + lane] = (Px)clampi((int)dst[j * stride + lane] + o[j], 0, maxv);
      }
    }
    __syncwarp();
    return;
  }
  const bool row_adst = (N < 32) && (tx_type & 2), col_adst = (N < 32) && (tx_type & 1);
  // rows: lane r transforms row r (all-zero rows stay zero)
  if (lane < N) {
    int in[N], out[N];
    int any = 0;
#pragma unroll
    for (int j = 0; j < N; j++) { in[j] = tbuf[lane * S + j]; any |= in[j]; }
    if (any) {
      txfm1d<PT, N>(in, out, row_adst);
#pragma unroll
      for (int j = 0; j < N; j++) tbuf[lane * S + j] = out[j];
    }
  }
  __syncwarp();
  // columns
  constexpr int shift = (N == 4) ? 4 : (N == 8 ? 5 : 6);
  if (lane < N) {
    int in[N], out[N];
#pragma unroll
    for (int j = 0; j < N; j++) in[j] = tbuf[j * S + lane];
    txfm1d<PT, N>(in, out, col_adst);
#pragma unroll
    for (int j = 0; j < N; j++) {
      Px *p = dst + (size_t)j * stride + lane;
      *p = (Px)clampi((int)*p + ((out[j] + (1 << (shift - 1))) >> shift), 0, maxv);
    }
  }
  __syncwarp();
}

template <typename Px, typename PT>
VP9B_DEV void warp_residual(const FrameParams &P, int *tbuf, const vp9b_txb &t, int lane) {
  if (!t.eob) return;
  const int plane = (t.tx >> 4) & 3, tx_size = t.tx & 3;
  const int stride = P.dst[plane].stride;
  Px *dst = (Px *)P.dst[plane].ptr + (size_t)t.y4 * 4 * stride + t.x4 * 4;
  const int cb = sizeof(Px) == 1 ? 2 : 4;
  switch (tx_size) {
    case 0: warp_itx_add<Px, PT, 4>(tbuf, P.coef, cb, t, dst, stride, P.bd, P.lossless, lane); break;
    case 1: warp_itx_add<Px, PT, 8>(tbuf, P.coef, cb, t, dst, stride, P.bd, 0, lane); break;
    case 2: warp_itx_add<Px, PT, 16>(tbuf, P.coef, cb, t, dst, stride, P.bd, 0, lane); break;
    default: warp_itx_add<Px, PT, 32>(tbuf, P.coef, cb, t, dst, stride, P.bd, 0, lane); break;
  }
}

constexpr int kResWarps = 4;
template <typename Px, typename PT>
__global__ void __launch_bounds__(kResWarps * 32) k_residual(const __grid_constant__ FrameParams P) {
  __shared__ int tbuf[kResWarps][32 * 33];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nwarps = gridDim.x * kResWarps;
  for (int i = blockIdx.x * kResWarps + warp; i < P.n_txb_inter; i += nwarps) {
    const vp9b_txb t = P.txb_inter[i];
    warp_residual<Px, PT>(P, tbuf[warp], t, lane);
  }
}

// ------------------------------------------------------------------------------------------
// stage 3: intra prediction.  Edge rules of build_intra_predictors (vp9_reconintra.c:113-402):
// unavailable left -> base+1, unavailable above -> base-1, above-right only read for 4x4 blocks
// with have_right, everything beyond the aligned frame size replicates the last sample.
// Predictors: vpx_dsp/intrapred.c (closed-form per sample).
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
VP9B_DEV void warp_intra_predict(const FrameParams &P, IntraEdge &E, const vp9b_txb &t, int lane) {
  const int tx_size = t.tx & 3, plane = (t.tx >> 4) & 3;
  const int mode = t.aux & 15, have_top = (t.aux >> 4) & 1, have_left = (t.aux >> 5) & 1, have_right = (t.aux >> 6) & 1;
  const int bs = 4 << tx_size, bd = P.bd;
  const int ss_x = plane ? P.ss_x : 0, ss_y = plane ? P.ss_y : 0;
  const int frame_w = (P.mi_cols * 8) >> ss_x, frame_h = (P.mi_rows * 8) >> ss_y;
  const int x0 = t.x4 * 4, y0 = t.y4 * 4;
  const int stride = P.dst[plane].stride;
  Px *dst = (Px *)P.dst[plane].ptr + (size_t)y0 * stride + x0;
  const int base = 128 << (bd - 8);
  // --- edges
  if (lane < bs) {
    int v = base + 1;
    if (have_left) {
      const int yy = (y0 + lane < frame_h) ? lane : frame_h - y0 - 1;
      v = ld_px(dst + (size_t)yy * stride - 1);
    }
    E.left[lane] = v;
  }
  for (int i = lane; i < 2 * bs; i += 32) {
    int v = base - 1;
    if (have_top) {
      const bool real = (i < bs) || (have_right && bs == 4);
      int xx = real ? i : bs - 1;
      if (x0 + xx >= frame_w) xx = frame_w - x0 - 1;
      v = ld_px(dst - stride + xx);
    }
    E.above[i] = v;
  }
  if (lane == 0) E.tl = have_top ? (have_left ? ld_px(dst - stride - 1) : base + 1) : base - 1;
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
    dc = cnt ? (s + (cnt >> 1)) / cnt : base;
  }
#define AT(i) ((i) < 0 ? tl : a[i])   /* above with above[-1] = top-left */
#define LT(i) ((i) < 0 ? tl : l[i])   /* left with left[-1] = top-left  */
  for (int p = lane; p < bs * bs; p += 32) {
    const int r = p / bs, c = p % bs;
    int v;
    switch (mode) {
      case 0: v = dc; break;
      case 1: v = a[c]; break;
      case 2: v = l[r]; break;
      case 9: v = clampi(l[r] + a[c] - tl, 0, maxv); break;
      case 3: {  // D45
        const int i = r + c;
        if (bs == 4) v = (i == 6) ? a[7] : AVG3(a[i], a[i + 1], a[i + 2]);
        else v = (i < bs - 1) ? AVG3(a[i], a[i + 1], a[i + 2]) : a[bs - 1];
        break;
      }
      case 8: {  // D63
        const int k = r >> 1, i = k + c;
        if (bs != 4 && r >= 2 && c >= bs - 1 - k) v = a[bs - 1];
        else v = (r & 1) ? AVG3(a[i], a[i + 1], a[i + 2]) : AVG2(a[i], a[i + 1]);
        break;
      }
      case 7: {  // D207
        const int j = r + (c >> 1);
        if (j >= bs - 1) v = l[bs - 1];
        else if (!(c & 1)) v = AVG2(l[j], l[j + 1]);
        else v = AVG3(l[j], l[j + 1], l[min(j + 2, bs - 1)]);
        break;
      }
      case 5: {  // D117
        const int k = min(r >> 1, c), rr = r - 2 * k, cc = c - k;
        if (rr == 0) v = AVG2(AT(cc - 1), a[cc]);
        else if (rr == 1) v = cc == 0 ? AVG3(l[0], tl, a[0]) : AVG3(AT(cc - 2), a[cc - 1], a[cc]);
        else v = AVG3(LT(rr - 3), l[rr - 2], l[rr - 1]);
        break;
      }
      case 4: {  // D135
        const int d = c - r;
        if (d == 0) v = AVG3(l[0], tl, a[0]);
        else if (d > 0) v = AVG3(AT(d - 2), a[d - 1], a[d]);
        else v = AVG3(LT(-d - 2), l[-d - 1], l[-d]);
        break;
      }
      default: {  // 6: D153
        const int k = min(r, c >> 1), rr = r - k, cc = c - 2 * k;
        if (cc >= 2) v = AVG3(AT(cc - 3), a[cc - 2], a[cc - 1]);
        else if (cc == 0) v = AVG2(LT(rr - 1), l[rr]);
        else if (rr == 0) v = AVG3(l[0], tl, a[0]);
        else v = AVG3(LT(rr - 2), l[rr - 1], l[rr]);
        break;
      }
    }
    dst[(size_t)r * stride + c] = (Px)v;
  }
#undef AT
#undef LT
  __syncwarp();
}

constexpr int kWaveWarps = 4;

// Persistent wavefront over superblocks.  Work units are handed out by an atomic ticket in an
// order in which every dependency has a smaller ticket, so a waiting warp always waits on a
// unit that is already running or done (forward progress without co-residency assumptions).
template <typename Px, typename PT>
__global__ void __launch_bounds__(kWaveWarps * 32) k_intra(const __grid_constant__ FrameParams P) {
  __shared__ int tbuf[kWaveWarps][32 * 33];
  __shared__ IntraEdge edges[kWaveWarps];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nsb = P.sb_cols * P.sb_rows;
  const uint32_t total = 3u * nsb;
  while (true) {
    uint32_t tk = 0;
    if (lane == 0) tk = atomicAdd(&P.tickets[0], 1u);
    tk = __shfl_sync(0xffffffffu, tk, 0);
    if (tk >= total) break;
    const int plane = tk % 3, sb = P.order_intra[tk / 3];
    const int sbr = sb / P.sb_cols, sbc = sb % P.sb_cols;
    const vp9b_sbrange rg = P.sb_intra[plane * nsb + sb];
    uint32_t *flags = P.flags_intra + plane * nsb;
    if (rg.count) {
      // wait for left, above-left, above (same plane)
      if (lane < 3) {
        const int dr = lane == 0 ? 0 : -1, dc = lane == 2 ? 0 : -1;
        const int r = sbr + dr, c = sbc + dc;
        if (r >= 0 && c >= 0) {
          const uint32_t *f = flags + r * P.sb_cols + c;
          while (ld_acquire(f) != P.epoch) __nanosleep(40);
        }
      }
      __syncwarp();
      const vp9b_txb *list = P.txb_intra[plane] + rg.start;
      for (uint32_t i = 0; i < rg.count; i++) {
        const vp9b_txb t = list[i];
        warp_intra_predict<Px>(P, edges[warp], t, lane);
        warp_residual<Px, PT>(P, tbuf[warp], t, lane);
      }
      __syncwarp();
      __threadfence();
    }
    if (lane == 0) st_release(flags + sb, P.epoch);
  }
}

// ------------------------------------------------------------------------------------------
// stage 4: loop filter.  Per (superblock, plane) unit, wavefront order 2*row+col (a unit
// depends on its left and above-right neighbours: the lag libvpx's own row-MT filter uses,
// vp9_thread_common.c:38-87).  The unit stages its 64x64 (+8 halo left/top) tile in shared
// memory, filters all vertical edges (lane = sample row) then all horizontal edges
// (lane = sample column) exactly in vp9_filter_block_plane_ss00/ss11 order
// (vp9_loopfilter.c:1241-1422) and writes the tile back.  Filter arithmetic:
// vpx_dsp/loopfilter.c:360-743.
struct LfTile {
  uint16_t px[72 * 73];  // stride 73 to spread banks for column walks
};

VP9B_DEV int sc_clamp(int t, int bd) {
  const int lim = 128 << (bd - 8);
  return clampi(t, -lim, lim - 1);
}

// s: pointer to q0 in the tile; step: distance between taps
VP9B_DEV void lf_px(uint16_t *s, int step, int size, int blimit, int limit, int thresh, int bd) {
  const int sh = bd - 8;
  const int p3 = s[-4 * step], p2 = s[-3 * step], p1 = s[-2 * step], p0 = s[-step];
  const int q0 = s[0], q1 = s[step], q2 = s[2 * step], q3 = s[3 * step];
  const int lim16 = limit << sh, blim16 = blimit << sh, one = 1 << sh;
  const bool fm = !(abs(p3 - p2) > lim16 || abs(p2 - p1) > lim16 || abs(p1 - p0) > lim16 || abs(q1 - q0) > lim16 ||
                    abs(q2 - q1) > lim16 || abs(q3 - q2) > lim16 || abs(p0 - q0) * 2 + abs(p1 - q1) / 2 > blim16);
  bool flat = false;
  if (size >= 8)
    flat = !(abs(p1 - p0) > one || abs(q1 - q0) > one || abs(p2 - p0) > one || abs(q2 - q0) > one ||
             abs(p3 - p0) > one || abs(q3 - q0) > one);
  if (size == 16 && flat && fm) {
    const int p7 = s[-8 * step], p6 = s[-7 * step], p5 = s[-6 * step], p4 = s[-5 * step];
    const int q4 = s[4 * step], q5 = s[5 * step], q6 = s[6 * step], q7 = s[7 * step];
    const bool flat2 = !(abs(p4 - p0) > one || abs(q4 - q0) > one || abs(p5 - p0) > one || abs(q5 - q0) > one ||
                         abs(p6 - p0) > one || abs(q6 - q0) > one || abs(p7 - p0) > one || abs(q7 - q0) > one);
    if (flat2) {
      s[-7 * step] = (uint16_t)((p7 * 7 + p6 * 2 + p5 + p4 + p3 + p2 + p1 + p0 + q0 + 8) >> 4);
      s[-6 * step] = (uint16_t)((p7 * 6 + p6 + p5 * 2 + p4 + p3 + p2 + p1 + p0 + q0 + q1 + 8) >> 4);
      s[-5 * step] = (uint16_t)((p7 * 5 + p6 + p5 + p4 * 2 + p3 + p2 + p1 + p0 + q0 + q1 + q2 + 8) >> 4);
      s[-4 * step] = (uint16_t)((p7 * 4 + p6 + p5 + p4 + p3 * 2 + p2 + p1 + p0 + q0 + q1 + q2 + q3 + 8) >> 4);
      s[-3 * step] = (uint16_t)((p7 * 3 + p6 + p5 + p4 + p3 + p2 * 2 + p1 + p0 + q0 + q1 + q2 + q3 + q4 + 8) >> 4);
      s[-2 * step] = (uint16_t)((p7 * 2 + p6 + p5 + p4 + p3 + p2 + p1 * 2 + p0 + q0 + q1 + q2 + q3 + q4 + q5 + 8) >> 4);
      s[-1 * step] = (uint16_t)((p7 + p6 + p5 + p4 + p3 + p2 + p1 + p0 * 2 + q0 + q1 + q2 + q3 + q4 + q5 + q6 + 8) >> 4);
      s[0] = (uint16_t)((p6 + p5 + p4 + p3 + p2 + p1 + p0 + q0 * 2 + q1 + q2 + q3 + q4 + q5 + q6 + q7 + 8) >> 4);
      s[1 * step] = (uint16_t)((p5 + p4 + p3 + p2 + p1 + p0 + q0 + q1 * 2 + q2 + q3 + q4 + q5 + q6 + q7 * 2 + 8) >> 4);
      s[2 * step] = (uint16_t)((p4 + p3 + p2 + p1 + p0 + q0 + q1 + q2 * 2 + q3 + q4 + q5 + q6 + q7 * 3 + 8) >> 4);
      s[3 * step] = (uint16_t)((p3 + p2 + p1 + p0 + q0 + q1 + q2 + q3 * 2 + q4 + q5 + q6 + q7 * 4 + 8) >> 4);
      s[4 * step] = (uint16_t)((p2 + p1 + p0 + q0 + q1 + q2 + q3 + q4 * 2 + q5 + q6 + q7 * 5 + 8) >> 4);
      s[5 * step] = (uint16_t)((p1 + p0 + q0 + q1 + q2 + q3 + q4 + q5 * 2 + q6 + q7 * 6 + 8) >> 4);
      s[6 * step] = (uint16_t)((p0 + q0 + q1 + q2 + q3 + q4 + q5 + q6 * 2 + q7 * 7 + 8) >> 4);
      return;
    }
  }
  if (flat && fm) {
    s[-3 * step] = (uint16_t)((p3 + p3 + p3 + 2 * p2 + p1 + p0 + q0 + 4) >> 3);
    s[-2 * step] = (uint16_t)((p3 + p3 + p2 + 2 * p1 + p0 + q0 + q1 + 4) >> 3);
    s[-1 * step] = (uint16_t)((p3 + p2 + p1 + 2 * p0 + q0 + q1 + q2 + 4) >> 3);
    s[0] = (uint16_t)((p2 + p1 + p0 + 2 * q0 + q1 + q2 + q3 + 4) >> 3);
    s[1 * step] = (uint16_t)((p1 + p0 + q0 + 2 * q1 + q2 + q3 + q3 + 4) >> 3);
    s[2 * step] = (uint16_t)((p0 + q0 + q1 + 2 * q2 + q3 + q3 + q3 + 4) >> 3);
    return;
  }
  // filter4
  const int off = 0x80 << sh;
  const int ps1 = p1 - off, ps0 = p0 - off, qs0 = q0 - off, qs1 = q1 - off;
  const int t16 = thresh << sh;
  const int hev = (abs(p1 - p0) > t16 || abs(q1 - q0) > t16) ? -1 : 0;
  const int mask = fm ? -1 : 0;
  int f = sc_clamp(ps1 - qs1, bd) & hev;
  f = sc_clamp(f + 3 * (qs0 - ps0), bd) & mask;
  const int f1 = sc_clamp(f + 4, bd) >> 3, f2 = sc_clamp(f + 3, bd) >> 3;
  s[0] = (uint16_t)(sc_clamp(qs0 - f1, bd) + off);
  s[-step] = (uint16_t)(sc_clamp(ps0 + f2, bd) + off);
  f = ((f1 + 1) >> 1) & ~hev;
  s[step] = (uint16_t)(sc_clamp(qs1 - f, bd) + off);
  s[-2 * step] = (uint16_t)(sc_clamp(ps1 + f, bd) + off);
}

template <typename Px>
__global__ void __launch_bounds__(kWaveWarps * 32) k_loopfilter(const __grid_constant__ FrameParams P) {
  __shared__ LfTile tiles[kWaveWarps];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nsb = P.sb_cols * P.sb_rows;
  const uint32_t total = 3u * nsb;
  constexpr int TS = 73;
  while (true) {
    uint32_t tk = 0;
    if (lane == 0) tk = atomicAdd(&P.tickets[1], 1u);
    tk = __shfl_sync(0xffffffffu, tk, 0);
    if (tk >= total) break;
    const int plane = tk % 3, sb = P.order_lf[tk / 3];
    const int sbr = sb / P.sb_cols, sbc = sb % P.sb_cols;
    uint32_t *flags = P.flags_lf + plane * nsb;
    // dependencies: left (r,c-1) and above-right (r-1,c+1) [or (r-1,c) in the last column]
    if (lane < 2) {
      int r = sbr, c = sbc - 1;
      if (lane == 1) { r = sbr - 1; c = min(sbc + 1, P.sb_cols - 1); }
      if (r >= 0 && c >= 0) {
        const uint32_t *f = flags + r * P.sb_cols + c;
        while (ld_acquire(f) != P.epoch) __nanosleep(40);
      }
    }
    __syncwarp();
    const vp9b_lfm &M = P.lfm[sb];
    const bool sub = plane && P.ss_x;  // 4:2:0 chroma
    const int n = sub ? 32 : 64;       // tile size in samples
    const int nb = n >> 3;             // 8-sample columns/rows per tile
    const int stride = P.dst[plane].stride;
    const int x0 = sbc * n, y0 = sbr * n;
    Px *g = (Px *)P.dst[plane].ptr;
    uint16_t *T = tiles[warp].px;
    const int mi_row = sbr * 8;
    // any edge in this unit?
    uint64_t anym;
    if (!sub) {
      if (plane == 0 || !P.ss_x)
        anym = M.left_y[0] | M.left_y[1] | M.left_y[2] | M.above_y[0] | M.above_y[1] | M.above_y[2] | M.int_4x4_y;
      else anym = 0;
    } else {
      anym = (uint64_t)(M.left_uv[0] | M.left_uv[1] | M.left_uv[2] | M.above_uv[0] | M.above_uv[1] | M.above_uv[2] |
                        M.int_4x4_uv);
    }
    if (anym) {
      const int hx = x0 >= 8 ? 8 : 0, hy = y0 >= 8 ? 8 : 0;  // halo present?
      // ---- load tile: rows y0-hy .. y0+n-1, cols x0-hx .. x0+n-1 ; tile origin (8,8) = (x0,y0)
      for (int ry = 8 - hy; ry < 8 + n; ry++) {
        const Px *src = g + (size_t)(y0 + ry - 8) * stride + x0 - 8;
        for (int cx = 8 - hx + lane; cx < 8 + n; cx += 32) T[ry * TS + cx] = (uint16_t)__ldcg(src + cx);
      }
      __syncwarp();
      // ---- vertical edges: lane = sample row
      for (int rb = 0; rb < n; rb += 32) {
        const int y = rb + lane;          // row inside the tile
        const int r8 = y >> 3;            // 8-sample row index inside the unit
        const int pr = r8 & ~1;           // first row of the row pair (filter_selectively_vert_row2)
        const int mrow = sub ? 2 * pr : pr;  // mi row of the pair start, for the frame-height test
        const bool active = (mi_row + mrow < P.mi_rows);
        if (active) {
          uint16_t *row = T + (8 + y) * TS + 8;
          for (int c = 0; c < nb; c++) {
            int b0, b1;  // bit positions of this column in rows pr and pr+1
            uint32_t m16, m8, m4, m4i;
            int lvl_own, lvl_first;
            if (!sub) {
              b0 = pr * 8 + c; b1 = b0 + 8;
              const int bo = r8 * 8 + c;
              m16 = (uint32_t)((M.left_y[2] >> bo) & 1) | ((uint32_t)((M.left_y[2] >> b0) & (M.left_y[2] >> b1) & 1) << 1);
              m8 = (uint32_t)((M.left_y[1] >> bo) & 1);
              m4 = (uint32_t)((M.left_y[0] >> bo) & 1);
              m4i = (uint32_t)((M.int_4x4_y >> bo) & 1);
              lvl_own = M.lfl_y[bo];
              lvl_first = M.lfl_y[b0];
            } else {
              b0 = pr * 4 + c; b1 = b0 + 4;
              const int bo = r8 * 4 + c;
              m16 = (uint32_t)((M.left_uv[2] >> bo) & 1) | ((uint32_t)((M.left_uv[2] >> b0) & (M.left_uv[2] >> b1) & 1) << 1);
              m8 = (uint32_t)((M.left_uv[1] >> bo) & 1);
              m4 = (uint32_t)((M.left_uv[0] >> bo) & 1);
              m4i = (uint32_t)((M.int_4x4_uv >> bo) & 1);
              lvl_own = M.lfl_y[(2 * r8) * 8 + 2 * c];
              lvl_first = M.lfl_y[(2 * pr) * 8 + 2 * c];
            }
            uint16_t *s = row + c * 8;
            if (m16 & 1) {
              const int lv = (m16 & 2) ? lvl_first : lvl_own;  // dual 16: both bands use the first band's level
              lf_px(s, 1, 16, P.lf_mblim[lv], P.lf_lim[lv], lv >> 4, P.bd);
            }
            if (m8) lf_px(s, 1, 8, P.lf_mblim[lvl_own], P.lf_lim[lvl_own], lvl_own >> 4, P.bd);
            if (m4) lf_px(s, 1, 4, P.lf_mblim[lvl_own], P.lf_lim[lvl_own], lvl_own >> 4, P.bd);
            if (m4i) lf_px(s + 4, 1, 4, P.lf_mblim[lvl_own], P.lf_lim[lvl_own], lvl_own >> 4, P.bd);
          }
        }
      }
      __syncwarp();
      // ---- horizontal edges: lane = sample column
      for (int cb = 0; cb < n; cb += 32) {
        const int x = cb + lane;
        const int c = x >> 3;
        uint16_t *col = T + 8 * TS + 8 + x;
        for (int r = 0; r < nb; r++) {
          const int mrow = sub ? 2 * r : r;
          if (mi_row + mrow >= P.mi_rows) break;
          uint32_t row16, row8, row4, row4i;
          const uint8_t *lfl;
          int lstep;
          if (!sub) {
            row16 = (uint32_t)(M.above_y[2] >> (r * 8)) & 0xff;
            row8 = (uint32_t)(M.above_y[1] >> (r * 8)) & 0xff;
            row4 = (uint32_t)(M.above_y[0] >> (r * 8)) & 0xff;
            row4i = (uint32_t)(M.int_4x4_y >> (r * 8)) & 0xff;
            lfl = &M.lfl_y[r * 8];
            lstep = 1;
          } else {
            row16 = (uint32_t)(M.above_uv[2] >> (r * 4)) & 0xf;
            row8 = (uint32_t)(M.above_uv[1] >> (r * 4)) & 0xf;
            row4 = (uint32_t)(M.above_uv[0] >> (r * 4)) & 0xf;
            row4i = (uint32_t)(M.int_4x4_uv >> (r * 4)) & 0xf;
            if (mi_row + mrow == P.mi_rows - 1) row4i = 0;  // skip_border_4x4_r
            lfl = &M.lfl_y[(2 * r) * 8];
            lstep = 2;
          }
          if (mi_row + mrow == 0) row16 = row8 = row4 = 0;
          uint16_t *s = col + r * 8 * TS;
          const int lvl = lfl[c * lstep];
          if ((row16 >> c) & 1) {
            // greedy pairing from the left: the second block of a pair uses the first one's level
            int start = c;
            while (start > 0 && ((row16 >> (start - 1)) & 1)) start--;
            const int lv = ((c - start) & 1) ? lfl[(c - 1) * lstep] : lvl;
            lf_px(s, TS, 16, P.lf_mblim[lv], P.lf_lim[lv], lv >> 4, P.bd);
          } else if ((row8 >> c) & 1) {
            lf_px(s, TS, 8, P.lf_mblim[lvl], P.lf_lim[lvl], lvl >> 4, P.bd);
            if ((row4i >> c) & 1) lf_px(s + 4 * TS, TS, 4, P.lf_mblim[lvl], P.lf_lim[lvl], lvl >> 4, P.bd);
          } else if ((row4 >> c) & 1) {
            lf_px(s, TS, 4, P.lf_mblim[lvl], P.lf_lim[lvl], lvl >> 4, P.bd);
            if ((row4i >> c) & 1) lf_px(s + 4 * TS, TS, 4, P.lf_mblim[lvl], P.lf_lim[lvl], lvl >> 4, P.bd);
          } else if ((row4i >> c) & 1) {
            lf_px(s + 4 * TS, TS, 4, P.lf_mblim[lvl], P.lf_lim[lvl], lvl >> 4, P.bd);
          }
        }
      }
      __syncwarp();
      // ---- store: top halo rows (cols x0..), then body rows (cols x0-hx..)
      for (int ry = 8 - hy; ry < 8 + n; ry++) {
        Px *dstp = g + (size_t)(y0 + ry - 8) * stride + x0 - 8;
        const int c_lo = (ry < 8) ? 8 : 8 - hx;
        for (int cx = c_lo + lane; cx < 8 + n; cx += 32) __stcg(dstp + cx, (Px)T[ry * TS + cx]);
      }
      __syncwarp();
      __threadfence();
    }
    if (lane == 0) st_release(flags + sb, P.epoch);
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
  int ws_sb_cols = 0, ws_sb_rows = 0, ws_mi_cols = 0, ws_mi_rows = 0;
  MiEntry *d_mi = nullptr;
  uint32_t *d_order_intra = nullptr, *d_order_lf = nullptr, *d_flags = nullptr, *d_tickets = nullptr;
  uint32_t epoch = 0;
  int stop_stage = 0;
  uint64_t launches = 0;
  int num_sms = 148;
  bool profiling = false;
  cudaEvent_t prof_ev[6];
  float stage_ms[8] = { 0 };
  bool tables_ready = false;
};

#define CK(call)                                                                       \
  do {                                                                                 \
    cudaError_t e_ = (call);                                                           \
    if (e_ != cudaSuccess) {                                                           \
      ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_);                   \
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
  CK(cudaMemcpyToSymbol(d_filters, f8, sizeof(f8)));
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
  for (int i = 0; i < 6; i++) cudaEventCreate(&ctx->prof_ev[i]);
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
  for (int i = 0; i < 6; i++) cudaEventDestroy(ctx->prof_ev[i]);
  cudaFree(ctx->d_mi); cudaFree(ctx->d_order_intra); cudaFree(ctx->d_order_lf); cudaFree(ctx->d_flags);
  cudaFree(ctx->d_tickets);
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
extern "C" int vp9gpu_sync(vp9gpu_ctx *ctx) {
  if (!ctx) return VP9B_ERR_INVALID;
  CK(cudaStreamSynchronize(ctx->stream));
  return 0;
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
  return 0;
}

static int setup_workspace(vp9gpu_ctx *ctx, const vp9b_frame_hdr *h) {
  if ((int)h->mi_cols != ctx->ws_mi_cols || (int)h->mi_rows != ctx->ws_mi_rows) {
    CK(cudaStreamSynchronize(ctx->stream));
    cudaFree(ctx->d_mi); cudaFree(ctx->d_order_intra); cudaFree(ctx->d_order_lf); cudaFree(ctx->d_flags);
    ctx->d_mi = nullptr; ctx->d_order_intra = ctx->d_order_lf = ctx->d_flags = nullptr;
    const int nsb = h->sb_cols * h->sb_rows;
    CK(cudaMalloc(&ctx->d_mi, sizeof(MiEntry) * (size_t)h->mi_cols * h->mi_rows));
    CK(cudaMemsetAsync(ctx->d_mi, 0, sizeof(MiEntry) * (size_t)h->mi_cols * h->mi_rows, ctx->stream));
    CK(cudaMalloc(&ctx->d_order_intra, sizeof(uint32_t) * nsb));
    CK(cudaMalloc(&ctx->d_order_lf, sizeof(uint32_t) * nsb));
    CK(cudaMalloc(&ctx->d_flags, sizeof(uint32_t) * nsb * 6));
    CK(cudaMemsetAsync(ctx->d_flags, 0, sizeof(uint32_t) * nsb * 6, ctx->stream));
    if (!ctx->d_tickets) CK(cudaMalloc(&ctx->d_tickets, sizeof(uint32_t) * 4));
    // ticket orders: every dependency of a unit has a smaller ticket
    std::vector<uint32_t> oi, ol;
    const int R = h->sb_rows, Cc = h->sb_cols;
    for (int d = 0; d < R + Cc - 1; d++)
      for (int r = 0; r < R; r++) {
        const int c = d - r;
        if (c >= 0 && c < Cc) oi.push_back(r * Cc + c);
      }
    for (int d = 0; d < 2 * R + Cc; d++)
      for (int r = 0; r < R; r++) {
        const int c = d - 2 * r;
        if (c >= 0 && c < Cc) ol.push_back(r * Cc + c);
      }
    CK(cudaMemcpyAsync(ctx->d_order_intra, oi.data(), sizeof(uint32_t) * nsb, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_order_lf, ol.data(), sizeof(uint32_t) * nsb, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->ws_mi_cols = h->mi_cols; ctx->ws_mi_rows = h->mi_rows;
    ctx->ws_sb_cols = h->sb_cols; ctx->ws_sb_rows = h->sb_rows;
    ctx->epoch = 0;
  }
  return 0;
}

template <typename Px, typename PT>
static int launch_stages(vp9gpu_ctx *ctx, const FrameParams &P, const vp9b_frame_hdr *h) {
  cudaStream_t st = ctx->stream;
  const int stop = ctx->stop_stage;
  const bool prof = ctx->profiling;
  if (prof) cudaEventRecord(ctx->prof_ev[0], st);
  if (!h->intra_only) {
    if (P.n_blocks) {
      k_expand_blocks<<<(P.n_blocks + 127) / 128, 128, 0, st>>>(P.blocks, P.n_blocks, P.mi, P.mi_cols, P.mi_rows);
      ctx->launches++;
    }
    if (prof) cudaEventRecord(ctx->prof_ev[1], st);
    for (int p = 0; p < 3; p++) {
      const int sx = p ? P.ss_x : 0, sy = p ? P.ss_y : 0;
      const int pw = (P.mi_cols * 8) >> sx, ph = (P.mi_rows * 8) >> sy;
      dim3 blk(32, 8), grd((pw + 31) / 32, (ph + 7) / 8);
      k_inter_pred_generic<Px><<<grd, blk, 0, st>>>(P, p);
      ctx->launches++;
    }
    if (prof) cudaEventRecord(ctx->prof_ev[2], st);
    if (stop != 1 && P.n_txb_inter) {
      const int blocks = std::min((P.n_txb_inter + kResWarps - 1) / kResWarps, ctx->num_sms * 8);
      k_residual<Px, PT><<<blocks, kResWarps * 32, 0, st>>>(P);
      ctx->launches++;
    }
  } else if (prof) {
    cudaEventRecord(ctx->prof_ev[1], st);
    cudaEventRecord(ctx->prof_ev[2], st);
  }
  if (prof) cudaEventRecord(ctx->prof_ev[3], st);
  if (stop == 1 || stop == 2) return 0;
  const int nsb = P.sb_cols * P.sb_rows;
  const int wave_blocks = std::min((3 * nsb + kWaveWarps - 1) / kWaveWarps, ctx->num_sms * 4);
  {
    const uint32_t n_intra = h->sec[VP9B_SEC_TXB_INTRA_Y].count + h->sec[VP9B_SEC_TXB_INTRA_U].count +
                             h->sec[VP9B_SEC_TXB_INTRA_V].count;
    if (n_intra) {
      k_intra<Px, PT><<<wave_blocks, kWaveWarps * 32, 0, st>>>(P);
      ctx->launches++;
    }
  }
  if (prof) cudaEventRecord(ctx->prof_ev[4], st);
  if (stop == 3) return 0;
  if (h->lf_level) {
    k_loopfilter<Px><<<wave_blocks, kWaveWarps * 32, 0, st>>>(P);
    ctx->launches++;
  }
  if (prof) {
    cudaEventRecord(ctx->prof_ev[5], st);
    cudaEventSynchronize(ctx->prof_ev[5]);
    for (int i = 0; i < 5; i++) {
      float ms = 0;
      cudaEventElapsedTime(&ms, ctx->prof_ev[i], ctx->prof_ev[i + 1]);
      ctx->stage_ms[i] += ms;
    }
  }
  return 0;
}

static int submit_device(vp9gpu_ctx *ctx, const uint8_t *d_packed, const vp9b_frame_hdr *h) {
  if (h->magic != VP9B_MAGIC) { ctx->err = "bad packed buffer magic"; return VP9B_ERR_INVALID; }
  if (h->dst_slot < 0 || h->dst_slot >= VP9B_NUM_SLOTS) { ctx->err = "bad dst slot"; return VP9B_ERR_INVALID; }
  if (!((h->ss_x == 1 && h->ss_y == 1) || (h->ss_x == 0 && h->ss_y == 0))) {
    ctx->err = "only 4:2:0 / 4:4:4 supported";
    return VP9B_ERR_UNSUP;
  }
  CK(cudaSetDevice(ctx->device));
  int e = setup_slot(ctx, h->dst_slot, h->width, h->height, h->bit_depth, h->ss_x, h->ss_y);
  if (e) return e;
  e = setup_workspace(ctx, h);
  if (e) return e;
  FrameParams P;
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
  for (int p = 0; p < 3; p++) P.txb_intra[p] = (const vp9b_txb *)(d_packed + h->sec[VP9B_SEC_TXB_INTRA_Y + p].offset);
  P.sb_intra = (const vp9b_sbrange *)(d_packed + h->sec[VP9B_SEC_SB_INTRA].offset);
  P.coef = d_packed + h->sec[VP9B_SEC_COEF].offset;
  P.lfm = (const vp9b_lfm *)(d_packed + h->sec[VP9B_SEC_LFM].offset);
  P.mi = ctx->d_mi;
  P.order_intra = ctx->d_order_intra;
  P.order_lf = ctx->d_order_lf;
  const int nsb = h->sb_cols * h->sb_rows;
  P.flags_intra = ctx->d_flags;
  P.flags_lf = ctx->d_flags + 3 * nsb;
  P.tickets = ctx->d_tickets;
  P.epoch = ++ctx->epoch;
  memcpy(P.lf_lim, h->lf_lim, 64);
  memcpy(P.lf_mblim, h->lf_mblim, 64);
  CK(cudaMemsetAsync(ctx->d_tickets, 0, sizeof(uint32_t) * 4, ctx->stream));
  if (h->bit_depth == 8) e = launch_stages<uint8_t, int>(ctx, P, h);
  else e = launch_stages<uint16_t, long long>(ctx, P, h);
  if (e) return e;
  CK(cudaGetLastError());
  return 0;
}

extern "C" int vp9gpu_frame_submit_device(vp9gpu_ctx *ctx, const void *d_packed, const vp9b_frame_hdr *hdr) {
  if (!ctx || !d_packed || !hdr) return VP9B_ERR_INVALID;
  return submit_device(ctx, (const uint8_t *)d_packed, hdr);
}

extern "C" int vp9gpu_frame_submit(vp9gpu_ctx *ctx, const void *packed, size_t bytes) {
  if (!ctx || !packed || bytes < sizeof(vp9b_frame_hdr)) return VP9B_ERR_INVALID;
  const vp9b_frame_hdr *h = (const vp9b_frame_hdr *)packed;
  if (h->magic != VP9B_MAGIC || h->total_bytes > bytes) { ctx->err = "corrupt packed buffer"; return VP9B_ERR_INVALID; }
  CK(cudaSetDevice(ctx->device));
  const int k = ctx->ring_pos;
  ctx->ring_pos = (k + 1) % vp9gpu_ctx::kRing;
  CK(cudaEventSynchronize(ctx->ring_ev[k]));  // previous user of this ring entry is done
  if (ctx->ring_bytes[k] < bytes) {
    const size_t nb = (bytes * 3 / 2 + (1 << 20)) & ~(size_t)4095;
    cudaFreeHost(ctx->h_pinned[k]);
    cudaFree(ctx->d_packed[k]);
    ctx->h_pinned[k] = ctx->d_packed[k] = nullptr;
    ctx->ring_bytes[k] = 0;
    CK(cudaMallocHost(&ctx->h_pinned[k], nb));
    CK(cudaMalloc(&ctx->d_packed[k], nb));
    ctx->ring_bytes[k] = nb;
  }
  memcpy(ctx->h_pinned[k], packed, h->total_bytes);
  CK(cudaMemcpyAsync(ctx->d_packed[k], ctx->h_pinned[k], h->total_bytes, cudaMemcpyHostToDevice, ctx->stream));
  const int e = submit_device(ctx, (const uint8_t *)ctx->d_packed[k], (const vp9b_frame_hdr *)ctx->h_pinned[k]);
  cudaEventRecord(ctx->ring_ev[k], ctx->stream);
  return e;
}

extern "C" int vp9gpu_upload_packed(vp9gpu_ctx *ctx, const void *packed, size_t bytes, void **d_packed) {
  if (!ctx || !packed || !d_packed) return VP9B_ERR_INVALID;
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

extern "C" int vp9gpu_get_frame(vp9gpu_ctx *ctx, int slot, uint8_t *y, uint8_t *u, uint8_t *v, int y_stride,
                                int uv_stride) {
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
  CK(cudaStreamSynchronize(ctx->stream));
  return 0;
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
