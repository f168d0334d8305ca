/*
 * TEST INFRASTRUCTURE ONLY -- CPU restatement (plain C) of the VP9 reconstruction hot path.
 *
 * This file is the parity oracle for the CUDA kernels.  It consumes the same packed per-frame
 * buffer (include/vp9_b200.h) and reproduces, sample for sample, what the reference's portable C
 * path computes.  It is imported only by tests/, __graft_entry__.smoke() and bench.py's CPU
 * baseline leg; the product never links it.
 *
 * Parity status: PINNED -- whole-stream per-frame MD5s of (product parser -> this file) equal
 * those of libvpx v1.15.2 (tests/test_oracle_streams.py), and the intra predictors reproduce the
 * golden MD5 signatures of libvpx/test/test_intra_pred_speed.cc:106-155,388-440
 * (tests/test_oracle_golden.py).
 *
 * All samples are held as uint16_t with an explicit bit depth: the reference's high-bitdepth C
 * functions evaluated at bd = 8 are arithmetically identical to its 8-bit functions, so one
 * code path serves 8/10/12 bit.
 *
 * Reference routines followed (paths relative to /root/reference/libvpx):
 *   inter   vp9/decoder/vp9_decodeframe.c:556-800, vp9/common/vp9_reconinter.c:61-124,
 *           vp9/common/vp9_scale.c:16-125, vpx_dsp/vpx_convolve.c:292-535, vp9_filter.c:15-82
 *   intra   vp9/common/vp9_reconintra.c:113-424, vpx_dsp/intrapred.c:457-850
 *   txfm    vp9/common/vp9_idct.c:20-204,308-396, vpx_dsp/inv_txfm.c:18-1278
 *   lpf     vp9/common/vp9_loopfilter.c:297-560,1241-1469, vpx_dsp/loopfilter.c:360-743
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/vp9_b200.h"
#include "vp9_tables_ref.h"

typedef struct vp9ref_frame {
  uint16_t *plane[3];
  int stride[3];
  int alloc_h[3];
  uint32_t width, height; /* crop, luma */
  int ss_x, ss_y, bit_depth;
  int valid;
} vp9ref_frame;

typedef struct vp9ref_ctx {
  vp9ref_frame slot[VP9B_NUM_SLOTS];
} vp9ref_ctx;

#define ROUND_POW2(v, n) (((v) + (1 << ((n)-1))) >> (n))

static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
static inline uint16_t clip_px(int v, int bd) { return (uint16_t)clampi(v, 0, (1 << bd) - 1); }

/* ------------------------------------------------------------------ frames ---------------- */
vp9ref_ctx *vp9ref_create(void) { return (vp9ref_ctx *)calloc(1, sizeof(vp9ref_ctx)); }

void vp9ref_destroy(vp9ref_ctx *c) {
  if (!c) return;
  for (int s = 0; s < VP9B_NUM_SLOTS; s++)
    for (int p = 0; p < 3; p++) free(c->slot[s].plane[p]);
  free(c);
}

static int setup_slot(vp9ref_frame *f, uint32_t w, uint32_t h, int ss_x, int ss_y, int bd) {
  const int aw = (int)((w + 7) & ~7u), ah = (int)((h + 7) & ~7u);
  for (int p = 0; p < 3; p++) {
    const int sx = p ? ss_x : 0, sy = p ? ss_y : 0;
    const int stride = ((aw >> sx) + 64 + 31) & ~31;
    const int rows = (ah >> sy) + 64;
    if (f->stride[p] != stride || f->alloc_h[p] != rows || !f->plane[p]) {
      free(f->plane[p]);
      f->plane[p] = (uint16_t *)calloc((size_t)stride * rows, sizeof(uint16_t));
      if (!f->plane[p]) return VP9B_ERR_MEM;
      f->stride[p] = stride;
      f->alloc_h[p] = rows;
    }
  }
  f->width = w;
  f->height = h;
  f->ss_x = ss_x;
  f->ss_y = ss_y;
  f->bit_depth = bd;
  f->valid = 1;
  return 0;
}

/* ------------------------------------------------------------------ inverse transforms ---- */
static const int cospi[32] = { 16384, 16364, 16305, 16207, 16069, 15893, 15679, 15426, 15137, 14811, 14449,
                               14053, 13623, 13160, 12665, 12140, 11585, 11003, 10394, 9760,  9102,  8423,
                               7723,  7005,  6270,  5520,  4756,  3981,  3196,  2404,  1606,  804 };
#define C(k) ((int64_t)cospi[k])
static inline int32_t rs14(int64_t v) { return (int32_t)((v + (1 << 13)) >> 14); } /* dct_const_round_shift */

/* inv_txfm.c:133-153 */
static void idct4(const int32_t *in, int32_t *out) {
  int32_t s0 = rs14((in[0] + in[2]) * C(16)), s1 = rs14((in[0] - in[2]) * C(16));
  int32_t s2 = rs14(in[1] * C(24) - in[3] * C(8)), s3 = rs14(in[1] * C(8) + in[3] * C(24));
  out[0] = s0 + s3;
  out[1] = s1 + s2;
  out[2] = s1 - s2;
  out[3] = s0 - s3;
}

/* inv_txfm.c:267-324 */
static void idct8(const int32_t *in, int32_t *out) {
  int32_t a[8], b[8];
  a[0] = in[0]; a[2] = in[4]; a[1] = in[2]; a[3] = in[6];
  a[4] = rs14(in[1] * C(28) - in[7] * C(4));
  a[7] = rs14(in[1] * C(4) + in[7] * C(28));
  a[5] = rs14(in[5] * C(12) - in[3] * C(20));
  a[6] = rs14(in[5] * C(20) + in[3] * C(12));
  b[0] = rs14((a[0] + a[2]) * C(16));
  b[1] = rs14((a[0] - a[2]) * C(16));
  b[2] = rs14(a[1] * C(24) - a[3] * C(8));
  b[3] = rs14(a[1] * C(8) + a[3] * C(24));
  b[4] = a[4] + a[5]; b[5] = a[4] - a[5]; b[6] = -a[6] + a[7]; b[7] = a[6] + a[7];
  a[0] = b[0] + b[3]; a[1] = b[1] + b[2]; a[2] = b[1] - b[2]; a[3] = b[0] - b[3];
  a[4] = b[4];
  a[5] = rs14((b[6] - b[5]) * C(16));
  a[6] = rs14((b[5] + b[6]) * C(16));
  a[7] = b[7];
  out[0] = a[0] + a[7]; out[1] = a[1] + a[6]; out[2] = a[2] + a[5]; out[3] = a[3] + a[4];
  out[4] = a[3] - a[4]; out[5] = a[2] - a[5]; out[6] = a[1] - a[6]; out[7] = a[0] - a[7];
}

/* inv_txfm.c:556-722 */
static void idct16(const int32_t *in, int32_t *out) {
  int32_t s1[16], s2[16];
  static const int order[16] = { 0, 8, 4, 12, 2, 10, 6, 14, 1, 9, 5, 13, 3, 11, 7, 15 };
  for (int i = 0; i < 16; i++) s1[i] = in[order[i]];
  /* stage 2 */
  for (int i = 0; i < 8; i++) s2[i] = s1[i];
  s2[8] = rs14(s1[8] * C(30) - s1[15] * C(2));   s2[15] = rs14(s1[8] * C(2) + s1[15] * C(30));
  s2[9] = rs14(s1[9] * C(14) - s1[14] * C(18));  s2[14] = rs14(s1[9] * C(18) + s1[14] * C(14));
  s2[10] = rs14(s1[10] * C(22) - s1[13] * C(10)); s2[13] = rs14(s1[10] * C(10) + s1[13] * C(22));
  s2[11] = rs14(s1[11] * C(6) - s1[12] * C(26));  s2[12] = rs14(s1[11] * C(26) + s1[12] * C(6));
  /* stage 3 */
  for (int i = 0; i < 4; i++) s1[i] = s2[i];
  s1[4] = rs14(s2[4] * C(28) - s2[7] * C(4));  s1[7] = rs14(s2[4] * C(4) + s2[7] * C(28));
  s1[5] = rs14(s2[5] * C(12) - s2[6] * C(20)); s1[6] = rs14(s2[5] * C(20) + s2[6] * C(12));
  s1[8] = s2[8] + s2[9];    s1[9] = s2[8] - s2[9];
  s1[10] = -s2[10] + s2[11]; s1[11] = s2[10] + s2[11];
  s1[12] = s2[12] + s2[13]; s1[13] = s2[12] - s2[13];
  s1[14] = -s2[14] + s2[15]; s1[15] = s2[14] + s2[15];
  /* stage 4 */
  s2[0] = rs14((s1[0] + s1[1]) * C(16)); s2[1] = rs14((s1[0] - s1[1]) * C(16));
  s2[2] = rs14(s1[2] * C(24) - s1[3] * C(8)); s2[3] = rs14(s1[2] * C(8) + s1[3] * C(24));
  s2[4] = s1[4] + s1[5]; s2[5] = s1[4] - s1[5]; s2[6] = -s1[6] + s1[7]; s2[7] = s1[6] + s1[7];
  s2[8] = s1[8]; s2[15] = s1[15];
  s2[9] = rs14(-s1[9] * C(8) + s1[14] * C(24));   s2[14] = rs14(s1[9] * C(24) + s1[14] * C(8));
  s2[10] = rs14(-s1[10] * C(24) - s1[13] * C(8)); s2[13] = rs14(-s1[10] * C(8) + s1[13] * C(24));
  s2[11] = s1[11]; s2[12] = s1[12];
  /* stage 5 */
  s1[0] = s2[0] + s2[3]; s1[1] = s2[1] + s2[2]; s1[2] = s2[1] - s2[2]; s1[3] = s2[0] - s2[3];
  s1[4] = s2[4];
  s1[5] = rs14((s2[6] - s2[5]) * C(16)); s1[6] = rs14((s2[5] + s2[6]) * C(16));
  s1[7] = s2[7];
  s1[8] = s2[8] + s2[11]; s1[9] = s2[9] + s2[10]; s1[10] = s2[9] - s2[10]; s1[11] = s2[8] - s2[11];
  s1[12] = -s2[12] + s2[15]; s1[13] = -s2[13] + s2[14]; s1[14] = s2[13] + s2[14]; s1[15] = s2[12] + s2[15];
  /* stage 6 */
  s2[0] = s1[0] + s1[7]; s2[1] = s1[1] + s1[6]; s2[2] = s1[2] + s1[5]; s2[3] = s1[3] + s1[4];
  s2[4] = s1[3] - s1[4]; s2[5] = s1[2] - s1[5]; s2[6] = s1[1] - s1[6]; s2[7] = s1[0] - s1[7];
  s2[8] = s1[8]; s2[9] = s1[9];
  s2[10] = rs14((-s1[10] + s1[13]) * C(16)); s2[13] = rs14((s1[10] + s1[13]) * C(16));
  s2[11] = rs14((-s1[11] + s1[12]) * C(16)); s2[12] = rs14((s1[11] + s1[12]) * C(16));
  s2[14] = s1[14]; s2[15] = s1[15];
  /* stage 7 */
  for (int i = 0; i < 8; i++) {
    out[i] = s2[i] + s2[15 - i];
    out[15 - i] = s2[i] - s2[15 - i];
  }
}

/* inv_txfm.c:816-1185 */
static void idct32(const int32_t *in, int32_t *out) {
  int32_t s1[32], s2[32];
  static const int order[16] = { 0, 16, 8, 24, 4, 20, 12, 28, 2, 18, 10, 26, 6, 22, 14, 30 };
  for (int i = 0; i < 16; i++) s1[i] = in[order[i]];
  s1[16] = rs14(in[1] * C(31) - in[31] * C(1));   s1[31] = rs14(in[1] * C(1) + in[31] * C(31));
  s1[17] = rs14(in[17] * C(15) - in[15] * C(17)); s1[30] = rs14(in[17] * C(17) + in[15] * C(15));
  s1[18] = rs14(in[9] * C(23) - in[23] * C(9));   s1[29] = rs14(in[9] * C(9) + in[23] * C(23));
  s1[19] = rs14(in[25] * C(7) - in[7] * C(25));   s1[28] = rs14(in[25] * C(25) + in[7] * C(7));
  s1[20] = rs14(in[5] * C(27) - in[27] * C(5));   s1[27] = rs14(in[5] * C(5) + in[27] * C(27));
  s1[21] = rs14(in[21] * C(11) - in[11] * C(21)); s1[26] = rs14(in[21] * C(21) + in[11] * C(11));
  s1[22] = rs14(in[13] * C(19) - in[19] * C(13)); s1[25] = rs14(in[13] * C(13) + in[19] * C(19));
  s1[23] = rs14(in[29] * C(3) - in[3] * C(29));   s1[24] = rs14(in[29] * C(29) + in[3] * C(3));
  /* stage 2 */
  for (int i = 0; i < 8; i++) s2[i] = s1[i];
  s2[8] = rs14(s1[8] * C(30) - s1[15] * C(2));    s2[15] = rs14(s1[8] * C(2) + s1[15] * C(30));
  s2[9] = rs14(s1[9] * C(14) - s1[14] * C(18));   s2[14] = rs14(s1[9] * C(18) + s1[14] * C(14));
  s2[10] = rs14(s1[10] * C(22) - s1[13] * C(10)); s2[13] = rs14(s1[10] * C(10) + s1[13] * C(22));
  s2[11] = rs14(s1[11] * C(6) - s1[12] * C(26));  s2[12] = rs14(s1[11] * C(26) + s1[12] * C(6));
  for (int i = 16; i < 32; i += 4) {
    s2[i] = s1[i] + s1[i + 1];      s2[i + 1] = s1[i] - s1[i + 1];
    s2[i + 2] = -s1[i + 2] + s1[i + 3]; s2[i + 3] = s1[i + 2] + s1[i + 3];
  }
  /* stage 3 */
  for (int i = 0; i < 4; i++) s1[i] = s2[i];
  s1[4] = rs14(s2[4] * C(28) - s2[7] * C(4));  s1[7] = rs14(s2[4] * C(4) + s2[7] * C(28));
  s1[5] = rs14(s2[5] * C(12) - s2[6] * C(20)); s1[6] = rs14(s2[5] * C(20) + s2[6] * C(12));
  s1[8] = s2[8] + s2[9];     s1[9] = s2[8] - s2[9];
  s1[10] = -s2[10] + s2[11]; s1[11] = s2[10] + s2[11];
  s1[12] = s2[12] + s2[13];  s1[13] = s2[12] - s2[13];
  s1[14] = -s2[14] + s2[15]; s1[15] = s2[14] + s2[15];
  s1[16] = s2[16]; s1[31] = s2[31];
  s1[17] = rs14(-s2[17] * C(4) + s2[30] * C(28));  s1[30] = rs14(s2[17] * C(28) + s2[30] * C(4));
  s1[18] = rs14(-s2[18] * C(28) - s2[29] * C(4));  s1[29] = rs14(-s2[18] * C(4) + s2[29] * C(28));
  s1[19] = s2[19]; s1[20] = s2[20];
  s1[21] = rs14(-s2[21] * C(20) + s2[26] * C(12)); s1[26] = rs14(s2[21] * C(12) + s2[26] * C(20));
  s1[22] = rs14(-s2[22] * C(12) - s2[25] * C(20)); s1[25] = rs14(-s2[22] * C(20) + s2[25] * C(12));
  s1[23] = s2[23]; s1[24] = s2[24]; s1[27] = s2[27]; s1[28] = s2[28];
  /* stage 4 */
  s2[0] = rs14((s1[0] + s1[1]) * C(16)); s2[1] = rs14((s1[0] - s1[1]) * C(16));
  s2[2] = rs14(s1[2] * C(24) - s1[3] * C(8)); s2[3] = rs14(s1[2] * C(8) + s1[3] * C(24));
  s2[4] = s1[4] + s1[5]; s2[5] = s1[4] - s1[5]; s2[6] = -s1[6] + s1[7]; s2[7] = s1[6] + s1[7];
  s2[8] = s1[8]; s2[15] = s1[15];
  s2[9] = rs14(-s1[9] * C(8) + s1[14] * C(24));   s2[14] = rs14(s1[9] * C(24) + s1[14] * C(8));
  s2[10] = rs14(-s1[10] * C(24) - s1[13] * C(8)); s2[13] = rs14(-s1[10] * C(8) + s1[13] * C(24));
  s2[11] = s1[11]; s2[12] = s1[12];
  s2[16] = s1[16] + s1[19]; s2[17] = s1[17] + s1[18]; s2[18] = s1[17] - s1[18]; s2[19] = s1[16] - s1[19];
  s2[20] = -s1[20] + s1[23]; s2[21] = -s1[21] + s1[22]; s2[22] = s1[21] + s1[22]; s2[23] = s1[20] + s1[23];
  s2[24] = s1[24] + s1[27]; s2[25] = s1[25] + s1[26]; s2[26] = s1[25] - s1[26]; s2[27] = s1[24] - s1[27];
  s2[28] = -s1[28] + s1[31]; s2[29] = -s1[29] + s1[30]; s2[30] = s1[29] + s1[30]; s2[31] = s1[28] + s1[31];
  /* stage 5 */
  s1[0] = s2[0] + s2[3]; s1[1] = s2[1] + s2[2]; s1[2] = s2[1] - s2[2]; s1[3] = s2[0] - s2[3];
  s1[4] = s2[4];
  s1[5] = rs14((s2[6] - s2[5]) * C(16)); s1[6] = rs14((s2[5] + s2[6]) * C(16));
  s1[7] = s2[7];
  s1[8] = s2[8] + s2[11]; s1[9] = s2[9] + s2[10]; s1[10] = s2[9] - s2[10]; s1[11] = s2[8] - s2[11];
  s1[12] = -s2[12] + s2[15]; s1[13] = -s2[13] + s2[14]; s1[14] = s2[13] + s2[14]; s1[15] = s2[12] + s2[15];
  s1[16] = s2[16]; s1[17] = s2[17];
  s1[18] = rs14(-s2[18] * C(8) + s2[29] * C(24));  s1[29] = rs14(s2[18] * C(24) + s2[29] * C(8));
  s1[19] = rs14(-s2[19] * C(8) + s2[28] * C(24));  s1[28] = rs14(s2[19] * C(24) + s2[28] * C(8));
  s1[20] = rs14(-s2[20] * C(24) - s2[27] * C(8));  s1[27] = rs14(-s2[20] * C(8) + s2[27] * C(24));
  s1[21] = rs14(-s2[21] * C(24) - s2[26] * C(8));  s1[26] = rs14(-s2[21] * C(8) + s2[26] * C(24));
  s1[22] = s2[22]; s1[23] = s2[23]; s1[24] = s2[24]; s1[25] = s2[25]; s1[30] = s2[30]; s1[31] = s2[31];
  /* stage 6 */
  s2[0] = s1[0] + s1[7]; s2[1] = s1[1] + s1[6]; s2[2] = s1[2] + s1[5]; s2[3] = s1[3] + s1[4];
  s2[4] = s1[3] - s1[4]; s2[5] = s1[2] - s1[5]; s2[6] = s1[1] - s1[6]; s2[7] = s1[0] - s1[7];
  s2[8] = s1[8]; s2[9] = s1[9];
  s2[10] = rs14((-s1[10] + s1[13]) * C(16)); s2[13] = rs14((s1[10] + s1[13]) * C(16));
  s2[11] = rs14((-s1[11] + s1[12]) * C(16)); s2[12] = rs14((s1[11] + s1[12]) * C(16));
  s2[14] = s1[14]; s2[15] = s1[15];
  s2[16] = s1[16] + s1[23]; s2[17] = s1[17] + s1[22]; s2[18] = s1[18] + s1[21]; s2[19] = s1[19] + s1[20];
  s2[20] = s1[19] - s1[20]; s2[21] = s1[18] - s1[21]; s2[22] = s1[17] - s1[22]; s2[23] = s1[16] - s1[23];
  s2[24] = -s1[24] + s1[31]; s2[25] = -s1[25] + s1[30]; s2[26] = -s1[26] + s1[29]; s2[27] = -s1[27] + s1[28];
  s2[28] = s1[27] + s1[28]; s2[29] = s1[26] + s1[29]; s2[30] = s1[25] + s1[30]; s2[31] = s1[24] + s1[31];
  /* stage 7 */
  for (int i = 0; i < 8; i++) {
    s1[i] = s2[i] + s2[15 - i];
    s1[15 - i] = s2[i] - s2[15 - i];
  }
  s1[16] = s2[16]; s1[17] = s2[17]; s1[18] = s2[18]; s1[19] = s2[19];
  s1[20] = rs14((-s2[20] + s2[27]) * C(16)); s1[27] = rs14((s2[20] + s2[27]) * C(16));
  s1[21] = rs14((-s2[21] + s2[26]) * C(16)); s1[26] = rs14((s2[21] + s2[26]) * C(16));
  s1[22] = rs14((-s2[22] + s2[25]) * C(16)); s1[25] = rs14((s2[22] + s2[25]) * C(16));
  s1[23] = rs14((-s2[23] + s2[24]) * C(16)); s1[24] = rs14((s2[23] + s2[24]) * C(16));
  s1[28] = s2[28]; s1[29] = s2[29]; s1[30] = s2[30]; s1[31] = s2[31];
  for (int i = 0; i < 16; i++) {
    out[i] = s1[i] + s1[31 - i];
    out[31 - i] = s1[i] - s1[31 - i];
  }
}

/* inv_txfm.c:94-131 */
static void iadst4(const int32_t *in, int32_t *out) {
  const int64_t sinpi_1_9 = 5283, sinpi_2_9 = 9929, sinpi_3_9 = 13377, sinpi_4_9 = 15212;
  int64_t x0 = in[0], x1 = in[1], x2 = in[2], x3 = in[3];
  if (!(x0 | x1 | x2 | x3)) { out[0] = out[1] = out[2] = out[3] = 0; return; }
  int64_t s0 = sinpi_1_9 * x0, s1 = sinpi_2_9 * x0, s2 = sinpi_3_9 * x1, s3 = sinpi_4_9 * x2;
  int64_t s4 = sinpi_1_9 * x2, s5 = sinpi_2_9 * x3, s6 = sinpi_4_9 * x3, s7 = x0 - x2 + x3;
  s0 = s0 + s3 + s5;
  s1 = s1 - s4 - s6;
  s3 = s2;
  s2 = sinpi_3_9 * s7;
  out[0] = rs14(s0 + s3);
  out[1] = rs14(s1 + s3);
  out[2] = rs14(s2);
  out[3] = rs14(s0 + s1 - s3);
}

/* inv_txfm.c:189-265 */
static void iadst8(const int32_t *in, int32_t *out) {
  int64_t x0 = in[7], x1 = in[0], x2 = in[5], x3 = in[2], x4 = in[3], x5 = in[4], x6 = in[1], x7 = in[6];
  if (!(x0 | x1 | x2 | x3 | x4 | x5 | x6 | x7)) { memset(out, 0, 8 * sizeof(*out)); return; }
  int64_t s0 = C(2) * x0 + C(30) * x1, s1 = C(30) * x0 - C(2) * x1;
  int64_t s2 = C(10) * x2 + C(22) * x3, s3 = C(22) * x2 - C(10) * x3;
  int64_t s4 = C(18) * x4 + C(14) * x5, s5 = C(14) * x4 - C(18) * x5;
  int64_t s6 = C(26) * x6 + C(6) * x7, s7 = C(6) * x6 - C(26) * x7;
  x0 = rs14(s0 + s4); x1 = rs14(s1 + s5); x2 = rs14(s2 + s6); x3 = rs14(s3 + s7);
  x4 = rs14(s0 - s4); x5 = rs14(s1 - s5); x6 = rs14(s2 - s6); x7 = rs14(s3 - s7);
  s0 = x0; s1 = x1; s2 = x2; s3 = x3;
  s4 = C(8) * x4 + C(24) * x5;  s5 = C(24) * x4 - C(8) * x5;
  s6 = -C(24) * x6 + C(8) * x7; s7 = C(8) * x6 + C(24) * x7;
  x0 = s0 + s2; x1 = s1 + s3; x2 = s0 - s2; x3 = s1 - s3;
  x4 = rs14(s4 + s6); x5 = rs14(s5 + s7); x6 = rs14(s4 - s6); x7 = rs14(s5 - s7);
  s2 = C(16) * (x2 + x3); s3 = C(16) * (x2 - x3); s6 = C(16) * (x6 + x7); s7 = C(16) * (x6 - x7);
  x2 = rs14(s2); x3 = rs14(s3); x6 = rs14(s6); x7 = rs14(s7);
  out[0] = (int32_t)x0;  out[1] = (int32_t)-x4; out[2] = (int32_t)x6;  out[3] = (int32_t)-x2;
  out[4] = (int32_t)x3;  out[5] = (int32_t)-x7; out[6] = (int32_t)x5;  out[7] = (int32_t)-x1;
}

/* inv_txfm.c:385-554 */
static void iadst16(const int32_t *in, int32_t *out) {
  int64_t x[16], s[16];
  static const int perm[16] = { 15, 0, 13, 2, 11, 4, 9, 6, 7, 8, 5, 10, 3, 12, 1, 14 };
  int64_t any = 0;
  for (int i = 0; i < 16; i++) { x[i] = in[perm[i]]; any |= x[i]; }
  if (!any) { memset(out, 0, 16 * sizeof(*out)); return; }
  /* stage 1 */
  static const int c1[8] = { 1, 5, 9, 13, 17, 21, 25, 29 };
  for (int i = 0; i < 8; i++) {
    s[2 * i] = x[2 * i] * C(c1[i]) + x[2 * i + 1] * C(32 - c1[i]);
    s[2 * i + 1] = x[2 * i] * C(32 - c1[i]) - x[2 * i + 1] * C(c1[i]);
  }
  for (int i = 0; i < 8; i++) {
    x[i] = rs14(s[i] + s[i + 8]);
    x[i + 8] = rs14(s[i] - s[i + 8]);
  }
  /* stage 2 */
  for (int i = 0; i < 8; i++) s[i] = x[i];
  s[8] = x[8] * C(4) + x[9] * C(28);     s[9] = x[8] * C(28) - x[9] * C(4);
  s[10] = x[10] * C(20) + x[11] * C(12); s[11] = x[10] * C(12) - x[11] * C(20);
  s[12] = -x[12] * C(28) + x[13] * C(4); s[13] = x[12] * C(4) + x[13] * C(28);
  s[14] = -x[14] * C(12) + x[15] * C(20); s[15] = x[14] * C(20) + x[15] * C(12);
  for (int i = 0; i < 4; i++) {
    x[i] = s[i] + s[i + 4];
    x[i + 4] = s[i] - s[i + 4];
    x[i + 8] = rs14(s[i + 8] + s[i + 12]);
    x[i + 12] = rs14(s[i + 8] - s[i + 12]);
  }
  /* stage 3 */
  s[0] = x[0]; s[1] = x[1]; s[2] = x[2]; s[3] = x[3];
  s[4] = x[4] * C(8) + x[5] * C(24);   s[5] = x[4] * C(24) - x[5] * C(8);
  s[6] = -x[6] * C(24) + x[7] * C(8);  s[7] = x[6] * C(8) + x[7] * C(24);
  s[8] = x[8]; s[9] = x[9]; s[10] = x[10]; s[11] = x[11];
  s[12] = x[12] * C(8) + x[13] * C(24);  s[13] = x[12] * C(24) - x[13] * C(8);
  s[14] = -x[14] * C(24) + x[15] * C(8); s[15] = x[14] * C(8) + x[15] * C(24);
  x[0] = s[0] + s[2]; x[1] = s[1] + s[3]; x[2] = s[0] - s[2]; x[3] = s[1] - s[3];
  x[4] = rs14(s[4] + s[6]); x[5] = rs14(s[5] + s[7]); x[6] = rs14(s[4] - s[6]); x[7] = rs14(s[5] - s[7]);
  x[8] = s[8] + s[10]; x[9] = s[9] + s[11]; x[10] = s[8] - s[10]; x[11] = s[9] - s[11];
  x[12] = rs14(s[12] + s[14]); x[13] = rs14(s[13] + s[15]); x[14] = rs14(s[12] - s[14]); x[15] = rs14(s[13] - s[15]);
  /* stage 4 */
  s[2] = (-C(16)) * (x[2] + x[3]);   s[3] = C(16) * (x[2] - x[3]);
  s[6] = C(16) * (x[6] + x[7]);      s[7] = C(16) * (-x[6] + x[7]);
  s[10] = C(16) * (x[10] + x[11]);   s[11] = C(16) * (-x[10] + x[11]);
  s[14] = (-C(16)) * (x[14] + x[15]); s[15] = C(16) * (x[14] - x[15]);
  x[2] = rs14(s[2]); x[3] = rs14(s[3]); x[6] = rs14(s[6]); x[7] = rs14(s[7]);
  x[10] = rs14(s[10]); x[11] = rs14(s[11]); x[14] = rs14(s[14]); x[15] = rs14(s[15]);
  out[0] = (int32_t)x[0];   out[1] = (int32_t)-x[8];  out[2] = (int32_t)x[12];  out[3] = (int32_t)-x[4];
  out[4] = (int32_t)x[6];   out[5] = (int32_t)x[14];  out[6] = (int32_t)x[10];  out[7] = (int32_t)x[2];
  out[8] = (int32_t)x[3];   out[9] = (int32_t)x[11];  out[10] = (int32_t)x[15]; out[11] = (int32_t)x[7];
  out[12] = (int32_t)x[5];  out[13] = (int32_t)-x[13]; out[14] = (int32_t)x[9];  out[15] = (int32_t)-x[1];
}

typedef void (*txfm1d)(const int32_t *, int32_t *);

/* vpx_iwht4x4_16_add_c (inv_txfm.c:18-66) */
static void iwht4x4_add(const int32_t *input, uint16_t *dest, int stride, int bd) {
  int32_t output[16];
  const int32_t *ip = input;
  int32_t *op = output;
  for (int i = 0; i < 4; i++) {
    int32_t a1 = ip[0] >> 2, c1 = ip[1] >> 2, d1 = ip[2] >> 2, b1 = ip[3] >> 2, e1;
    a1 += c1; d1 -= b1; e1 = (a1 - d1) >> 1; b1 = e1 - b1; c1 = e1 - c1; a1 -= b1; d1 += c1;
    op[0] = a1; op[1] = b1; op[2] = c1; op[3] = d1;
    ip += 4; op += 4;
  }
  ip = output;
  for (int i = 0; i < 4; i++) {
    int32_t a1 = ip[0], c1 = ip[4], d1 = ip[8], b1 = ip[12], e1;
    a1 += c1; d1 -= b1; e1 = (a1 - d1) >> 1; b1 = e1 - b1; c1 = e1 - c1; a1 -= b1; d1 += c1;
    dest[stride * 0] = clip_px(dest[stride * 0] + a1, bd);
    dest[stride * 1] = clip_px(dest[stride * 1] + b1, bd);
    dest[stride * 2] = clip_px(dest[stride * 2] + c1, bd);
    dest[stride * 3] = clip_px(dest[stride * 3] + d1, bd);
    ip++; dest++;
  }
}

static const int16_t *scan_for(int tx_size, int tx_type) {
  switch (tx_size) {
    case 0: return tx_type == 1 ? kRowScan_4x4 : (tx_type == 2 ? kColScan_4x4 : kDefaultScan_4x4);
    case 1: return tx_type == 1 ? kRowScan_8x8 : (tx_type == 2 ? kColScan_8x8 : kDefaultScan_8x8);
    case 2: return tx_type == 1 ? kRowScan_16x16 : (tx_type == 2 ? kColScan_16x16 : kDefaultScan_16x16);
    default: return kDefaultScan_32x32;
  }
}

/* Inverse transform of one block + add to dst.  coefs: raster order, n*n, dequantised.
 * vp9_idct.c:20-204 (dispatch), vpx_dsp/inv_txfm.c (vpx_idctNxN_*_add_c): rows then columns,
 * final shift 4/5/6/6. */
void vp9ref_inverse_transform_add(const int32_t *coefs, int tx_size, int tx_type, int lossless,
                                  uint16_t *dst, int stride, int bd) {
  const int n = 4 << tx_size;
  if (lossless) { iwht4x4_add(coefs, dst, stride, bd); return; }
  static const txfm1d dct[4] = { idct4, idct8, idct16, idct32 };
  static const txfm1d adst[3] = { iadst4, iadst8, iadst16 };
  /* tx_type: 0 DCT_DCT, 1 ADST_DCT (cols adst, rows dct), 2 DCT_ADST, 3 ADST_ADST */
  const txfm1d rows = (tx_size < 3 && (tx_type & 2)) ? adst[tx_size] : dct[tx_size];
  const txfm1d cols = (tx_size < 3 && (tx_type & 1)) ? adst[tx_size] : dct[tx_size];
  const int shift = tx_size == 0 ? 4 : (tx_size == 1 ? 5 : 6);
  int32_t out[32 * 32], tin[32], tout[32];
  for (int i = 0; i < n; i++) rows(coefs + i * n, out + i * n);
  for (int i = 0; i < n; i++) {
    for (int j = 0; j < n; j++) tin[j] = out[j * n + i];
    cols(tin, tout);
    for (int j = 0; j < n; j++) dst[j * stride + i] = clip_px(dst[j * stride + i] + ROUND_POW2(tout[j], shift), bd);
  }
}

static void residual_add(const vp9b_frame_hdr *h, const uint8_t *base, const vp9b_txb *t, vp9ref_frame *f) {
  if (!t->eob) return;
  const int tx_size = t->tx & 3, tx_type = (t->tx >> 2) & 3, plane = (t->tx >> 4) & 3;
  const int n = 4 << tx_size;
  int32_t coefs[32 * 32];
  memset(coefs, 0, sizeof(int32_t) * n * n);
  const int16_t *scan = scan_for(tx_size, h->lossless ? 0 : tx_type);
  if (h->coef_bytes == 2) {
    const int16_t *c = (const int16_t *)(base + h->sec[VP9B_SEC_COEF].offset) + t->coef_off;
    for (int i = 0; i < t->eob; i++) coefs[scan[i]] = c[i];
  } else {
    const int32_t *c = (const int32_t *)(base + h->sec[VP9B_SEC_COEF].offset) + t->coef_off;
    for (int i = 0; i < t->eob; i++) coefs[scan[i]] = c[i];
  }
  uint16_t *dst = f->plane[plane] + (size_t)t->y4 * 4 * f->stride[plane] + t->x4 * 4;
  vp9ref_inverse_transform_add(coefs, tx_size, tx_type, h->lossless, dst, f->stride[plane], h->bit_depth);
}

/* ------------------------------------------------------------------ inter prediction ------ */
static inline int scaled_x(int v, const vp9b_ref_desc *r) { return (int)((int64_t)v * r->x_scale_fp >> 14); }
static inline int scaled_y(int v, const vp9b_ref_desc *r) { return (int)((int64_t)v * r->y_scale_fp >> 14); }

/* One w x h prediction of plane `plane` at block-relative (x,y) -- dec_build_inter_predictors
 * (vp9_decodeframe.c:556-690) + highbd_convolve (vpx_convolve.c:292-420).  Every reference
 * sample is fetched with coordinates clamped to the reference plane's crop size, which is what
 * build_mc_border (vp9_decodeframe.c:432-504) emulates. */
static void predict_block(const vp9b_frame_hdr *h, vp9ref_frame *dstf, const vp9ref_frame *reff,
                          const vp9b_ref_desc *rd, int plane, int mi_row, int mi_col, int bw_px, int bh_px,
                          int x, int y, int w, int h_, int mv_row, int mv_col, int filter, int is_second,
                          int mb_to_left, int mb_to_right, int mb_to_top, int mb_to_bottom, int vis_w, int vis_h) {
  const int ss_x = plane ? h->ss_x : 0, ss_y = plane ? h->ss_y : 0;
  const int bd = h->bit_depth;
  const int is_scaled = rd->x_step_q4 != 16 || rd->y_step_q4 != 16;
  const int ref_w = plane ? (int)((rd->width + ss_x) >> ss_x) : (int)rd->width;
  const int ref_h = plane ? (int)((rd->height + ss_y) >> ss_y) : (int)rd->height;
  int x0, y0, xs, ys, mvr, mvc;
  const int mi_x = mi_col * 8, mi_y = mi_row * 8;
  if (is_scaled) {
    /* clamp_mv_to_umv_border_sb (vp9_reconinter.c:90-110) */
    const int spel_left = (4 + bw_px) << 4, spel_right = spel_left - 16;
    const int spel_top = (4 + bh_px) << 4, spel_bottom = spel_top - 16;
    int qrow = (int16_t)(mv_row * (1 << (1 - ss_y))), qcol = (int16_t)(mv_col * (1 << (1 - ss_x)));
    qcol = clampi(qcol, mb_to_left * (1 << (1 - ss_x)) - spel_left, mb_to_right * (1 << (1 - ss_x)) + spel_right);
    qrow = clampi(qrow, mb_to_top * (1 << (1 - ss_y)) - spel_top, mb_to_bottom * (1 << (1 - ss_y)) + spel_bottom);
    const int x_start = (-mb_to_left >> (3 + ss_x)), y_start = (-mb_to_top >> (3 + ss_y));
    x0 = scaled_x(x_start + x, rd);
    y0 = scaled_y(y_start + y, rd);
    /* vp9_scale_mv (vp9_scale.c:37-43) */
    const int x_off_q4 = scaled_x((mi_x + x) << 4, rd) & 15;
    const int y_off_q4 = scaled_y((mi_y + y) << 4, rd) & 15;
    mvr = scaled_y(qrow, rd) + y_off_q4;
    mvc = scaled_x(qcol, rd) + x_off_q4;
    xs = rd->x_step_q4;
    ys = rd->y_step_q4;
  } else {
    x0 = (-mb_to_left >> (3 + ss_x)) + x;
    y0 = (-mb_to_top >> (3 + ss_y)) + y;
    mvr = mv_row * (1 << (1 - ss_y));
    mvc = mv_col * (1 << (1 - ss_x));
    xs = ys = 16;
  }
  const int subpel_x = mvc & 15, subpel_y = mvr & 15;
  x0 += mvc >> 4;
  y0 += mvr >> 4;
  const int16_t *taps = kSubpelFilters + filter * 128;
  const uint16_t *ref = reff->plane[plane];
  const int rstride = reff->stride[plane];
  /* horizontal pass into temp (with clip), rows y0-3 .. */
  static uint16_t temp[64 * 140];
  const int inter_h = (((h_ - 1) * ys + subpel_y) >> 4) + 8;
  for (int r = 0; r < inter_h; r++) {
    const int ry = clampi(y0 - 3 + r, 0, ref_h - 1);
    int x_q4 = subpel_x;
    for (int c = 0; c < w; c++) {
      const int16_t *f = taps + (x_q4 & 15) * 8;
      const int bx = x0 - 3 + (x_q4 >> 4);
      int sum = 0;
      for (int k = 0; k < 8; k++) sum += ref[(size_t)ry * rstride + clampi(bx + k, 0, ref_w - 1)] * f[k];
      temp[r * 64 + c] = clip_px(ROUND_POW2(sum, 7), bd);
      x_q4 += xs;
    }
  }
  uint16_t *dst = dstf->plane[plane] + (size_t)((-mb_to_top >> (3 + ss_y)) + y) * dstf->stride[plane] +
                  ((-mb_to_left >> (3 + ss_x)) + x);
  const int dstride = dstf->stride[plane];
  for (int c = 0; c < w && c < vis_w; c++) {
    int y_q4 = subpel_y;
    for (int r = 0; r < h_; r++) {
      const int16_t *f = taps + (y_q4 & 15) * 8;
      const uint16_t *s = temp + (y_q4 >> 4) * 64 + c;
      int sum = 0;
      for (int k = 0; k < 8; k++) sum += s[k * 64] * f[k];
      const int v = clip_px(ROUND_POW2(sum, 7), bd);
      if (r < vis_h) {
        if (is_second) dst[r * dstride + c] = (uint16_t)ROUND_POW2(dst[r * dstride + c] + v, 1);
        else dst[r * dstride + c] = (uint16_t)v;
      }
      y_q4 += ys;
    }
  }
}

static const uint8_t kBw8[13] = { 1, 1, 1, 1, 1, 2, 2, 2, 4, 4, 4, 8, 8 };
static const uint8_t kBh8[13] = { 1, 1, 1, 1, 2, 1, 2, 4, 2, 4, 8, 4, 8 };

/* dec_build_inter_predictors_sb (vp9_decodeframe.c:703-800) */
static int inter_predict_frame(vp9ref_ctx *c, const vp9b_frame_hdr *h, const uint8_t *base) {
  const vp9b_block *blk = (const vp9b_block *)(base + h->sec[VP9B_SEC_BLOCKS].offset);
  const vp9b_submv *sub = (const vp9b_submv *)(base + h->sec[VP9B_SEC_SUBMV].offset);
  vp9ref_frame *dst = &c->slot[h->dst_slot];
  for (uint32_t i = 0; i < h->sec[VP9B_SEC_BLOCKS].count; i++) {
    const vp9b_block *b = &blk[i];
    if (!b->ref0) continue;
    const int bw = kBw8[b->bsize], bh = kBh8[b->bsize];
    const int mb_to_left = -(b->mi_col * 64), mb_to_top = -(b->mi_row * 64);
    const int mb_to_right = ((int)h->mi_cols - bw - b->mi_col) * 64;
    const int mb_to_bottom = ((int)h->mi_rows - bh - b->mi_row) * 64;
    for (int r = 0; r < (b->ref1 ? 2 : 1); r++) {
      const int refn = r ? b->ref1 : b->ref0;
      const vp9b_ref_desc *rd = &h->ref[refn - 1];
      if (rd->slot < 0 || !c->slot[rd->slot].valid || rd->x_scale_fp < 0) return VP9B_ERR_CORRUPT;
      const vp9ref_frame *rf = &c->slot[rd->slot];
      for (int plane = 0; plane < 3; plane++) {
        const int ss_x = plane ? h->ss_x : 0, ss_y = plane ? h->ss_y : 0;
        const int n4w_x4 = 4 * ((bw << 1) >> ss_x), n4h_x4 = 4 * ((bh << 1) >> ss_y);
        /* visible (inside the mi grid) part of the block in this plane */
        const int vis_w = (((int)h->mi_cols - b->mi_col) * 8) >> ss_x;
        const int vis_h = (((int)h->mi_rows - b->mi_row) * 8) >> ss_y;
        if (b->bsize < 3) {
          uint32_t si;
          memcpy(&si, &b->mv[0][0], 4);
          const vp9b_submv *s = &sub[si];
          const int nw = (bw << 1) >> ss_x, nh = (bh << 1) >> ss_y;
          int k = 0;
          for (int y = 0; y < nh; y++)
            for (int x = 0; x < nw; x++, k++) {
              /* average_split_mvs (vp9_reconinter.c:90-113 / dec_build_inter_predictors_sb): the block
               * index is the running count k of the plane's own 4x4 grid; 4:2:2 averages blocks k, k+1,
               * 4:4:0 blocks k, k+2 (mi_mv_pred_q2), 4:2:0 all four (q4, precomputed), 4:4:4 none */
              int mr, mc;
              if (ss_x && ss_y) { mr = s->chroma[r][0]; mc = s->chroma[r][1]; }
              else if (ss_x || ss_y) {
                const int k2 = k + (ss_x ? 1 : 2);
                const int sr = s->luma[k][r][0] + s->luma[k2][r][0], sc = s->luma[k][r][1] + s->luma[k2][r][1];
                mr = (sr < 0 ? sr - 1 : sr + 1) / 2;
                mc = (sc < 0 ? sc - 1 : sc + 1) / 2;
              } else { mr = s->luma[k][r][0]; mc = s->luma[k][r][1]; }
              predict_block(h, dst, rf, rd, plane, b->mi_row, b->mi_col, n4w_x4, n4h_x4, 4 * x, 4 * y, 4, 4, mr, mc,
                            b->filter, r, mb_to_left, mb_to_right, mb_to_top, mb_to_bottom, vis_w - 4 * x, vis_h - 4 * y);
            }
        } else {
          predict_block(h, dst, rf, rd, plane, b->mi_row, b->mi_col, n4w_x4, n4h_x4, 0, 0, n4w_x4, n4h_x4, b->mv[r][0],
                        b->mv[r][1], b->filter, r, mb_to_left, mb_to_right, mb_to_top, mb_to_bottom, vis_w, vis_h);
        }
      }
    }
  }
  return 0;
}

/* ------------------------------------------------------------------ intra prediction ------ */
#define AVG2(a, b) (((a) + (b) + 1) >> 1)
#define AVG3(a, b, c) (((a) + 2 * (b) + (c) + 2) >> 2)

/* vpx_dsp/intrapred.c:457-850 (highbd_*_predictor + dedicated 4x4 d45/d63).  above[-1] valid. */
void vp9ref_intra_predictor(int mode, int bs, int have_left, int have_top, const uint16_t *above,
                            const uint16_t *left, uint16_t *dst, int stride, int bd) {
  int r, c;
  switch (mode) {
    case 0: { /* DC family: dc_pred[left_available][up_available] */
      int sum = 0, dc;
      if (have_left && have_top) {
        for (c = 0; c < bs; c++) sum += above[c] + left[c];
        dc = (sum + bs) / (2 * bs);
      } else if (have_top) {
        for (c = 0; c < bs; c++) sum += above[c];
        dc = (sum + (bs >> 1)) / bs;
      } else if (have_left) {
        for (c = 0; c < bs; c++) sum += left[c];
        dc = (sum + (bs >> 1)) / bs;
      } else {
        dc = 128 << (bd - 8);
      }
      for (r = 0; r < bs; r++)
        for (c = 0; c < bs; c++) dst[r * stride + c] = (uint16_t)dc;
      break;
    }
    case 1: /* V */
      for (r = 0; r < bs; r++)
        for (c = 0; c < bs; c++) dst[r * stride + c] = above[c];
      break;
    case 2: /* H */
      for (r = 0; r < bs; r++)
        for (c = 0; c < bs; c++) dst[r * stride + c] = left[r];
      break;
    case 3: /* D45 */
      if (bs == 4) {
        const int A = above[0], B = above[1], Cc = above[2], D = above[3], E = above[4], F = above[5], G = above[6],
                  H = above[7];
#define DST(x, y) dst[(x) + (y)*stride]
        DST(0, 0) = AVG3(A, B, Cc);
        DST(1, 0) = DST(0, 1) = AVG3(B, Cc, D);
        DST(2, 0) = DST(1, 1) = DST(0, 2) = AVG3(Cc, D, E);
        DST(3, 0) = DST(2, 1) = DST(1, 2) = DST(0, 3) = AVG3(D, E, F);
        DST(3, 1) = DST(2, 2) = DST(1, 3) = AVG3(E, F, G);
        DST(3, 2) = DST(2, 3) = AVG3(F, G, H);
        DST(3, 3) = H;
      } else {
        const uint16_t above_right = above[bs - 1];
        for (c = 0; c < bs - 1; ++c) dst[c] = AVG3(above[c], above[c + 1], above[c + 2]);
        dst[bs - 1] = above_right;
        for (r = 1; r < bs; ++r) {
          const int size = bs - 1 - r;
          for (c = 0; c < size; c++) dst[r * stride + c] = dst[c + r];
          for (c = size > 0 ? size : 0; c < bs; c++) dst[r * stride + c] = above_right;
        }
      }
      break;
    case 8: /* D63 */
      if (bs == 4) {
        const int A = above[0], B = above[1], Cc = above[2], D = above[3], E = above[4], F = above[5], G = above[6];
        DST(0, 0) = AVG2(A, B);
        DST(1, 0) = DST(0, 2) = AVG2(B, Cc);
        DST(2, 0) = DST(1, 2) = AVG2(Cc, D);
        DST(3, 0) = DST(2, 2) = AVG2(D, E);
        DST(3, 2) = AVG2(E, F);
        DST(0, 1) = AVG3(A, B, Cc);
        DST(1, 1) = DST(0, 3) = AVG3(B, Cc, D);
        DST(2, 1) = DST(1, 3) = AVG3(Cc, D, E);
        DST(3, 1) = DST(2, 3) = AVG3(D, E, F);
        DST(3, 3) = AVG3(E, F, G);
      } else {
        for (c = 0; c < bs; ++c) {
          dst[c] = AVG2(above[c], above[c + 1]);
          dst[stride + c] = AVG3(above[c], above[c + 1], above[c + 2]);
        }
        int size;
        for (r = 2, size = bs - 2; r < bs; r += 2, --size) {
          for (c = 0; c < size; c++) dst[r * stride + c] = dst[(r >> 1) + c];
          for (c = size; c < bs; c++) dst[r * stride + c] = above[bs - 1];
          for (c = 0; c < size; c++) dst[(r + 1) * stride + c] = dst[stride + (r >> 1) + c];
          for (c = size; c < bs; c++) dst[(r + 1) * stride + c] = above[bs - 1];
        }
      }
      break;
    case 7: { /* D207 */
      for (r = 0; r < bs - 1; ++r) dst[r * stride] = AVG2(left[r], left[r + 1]);
      dst[(bs - 1) * stride] = left[bs - 1];
      for (r = 0; r < bs - 2; ++r) dst[r * stride + 1] = AVG3(left[r], left[r + 1], left[r + 2]);
      dst[(bs - 2) * stride + 1] = AVG3(left[bs - 2], left[bs - 1], left[bs - 1]);
      dst[(bs - 1) * stride + 1] = left[bs - 1];
      for (c = 0; c < bs - 2; ++c) dst[(bs - 1) * stride + c + 2] = left[bs - 1];
      for (r = bs - 2; r >= 0; --r)
        for (c = 0; c < bs - 2; ++c) dst[r * stride + c + 2] = dst[(r + 1) * stride + c];
      break;
    }
    case 5: { /* D117 */
      for (c = 0; c < bs; c++) dst[c] = AVG2(above[c - 1], above[c]);
      dst[stride] = AVG3(left[0], above[-1], above[0]);
      for (c = 1; c < bs; c++) dst[stride + c] = AVG3(above[c - 2], above[c - 1], above[c]);
      dst[2 * stride] = AVG3(above[-1], left[0], left[1]);
      for (r = 3; r < bs; ++r) dst[r * stride] = AVG3(left[r - 3], left[r - 2], left[r - 1]);
      for (r = 2; r < bs; ++r)
        for (c = 1; c < bs; c++) dst[r * stride + c] = dst[(r - 2) * stride + c - 1];
      break;
    }
    case 4: { /* D135 */
      uint16_t border[32 + 32 - 1];
      int i;
      for (i = 0; i < bs - 2; ++i) border[i] = AVG3(left[bs - 3 - i], left[bs - 2 - i], left[bs - 1 - i]);
      border[bs - 2] = AVG3(above[-1], left[0], left[1]);
      border[bs - 1] = AVG3(left[0], above[-1], above[0]);
      border[bs - 0] = AVG3(above[-1], above[0], above[1]);
      for (i = 0; i < bs - 2; ++i) border[bs + 1 + i] = AVG3(above[i], above[i + 1], above[i + 2]);
      for (i = 0; i < bs; ++i)
        for (c = 0; c < bs; c++) dst[i * stride + c] = border[bs - 1 - i + c];
      break;
    }
    case 6: { /* D153 */
      dst[0] = AVG2(above[-1], left[0]);
      for (r = 1; r < bs; r++) dst[r * stride] = AVG2(left[r - 1], left[r]);
      dst[1] = AVG3(left[0], above[-1], above[0]);
      dst[stride + 1] = AVG3(above[-1], left[0], left[1]);
      for (r = 2; r < bs; r++) dst[r * stride + 1] = AVG3(left[r - 2], left[r - 1], left[r]);
      for (c = 0; c < bs - 2; c++) dst[c + 2] = AVG3(above[c - 1], above[c], above[c + 1]);
      for (r = 1; r < bs; ++r)
        for (c = 0; c < bs - 2; c++) dst[r * stride + c + 2] = dst[(r - 1) * stride + c];
      break;
    }
    case 9: { /* TM */
      const int tl = above[-1];
      for (r = 0; r < bs; r++)
        for (c = 0; c < bs; c++) dst[r * stride + c] = clip_px(left[r] + above[c] - tl, bd);
      break;
    }
  }
}

/* build_intra_predictors_high (vp9_reconintra.c:113-259) reduced to its sample semantics:
 * a sample is taken from the frame when the availability rules allow it and its coordinate is
 * below the aligned frame size, otherwise the last available one is replicated. */
static void intra_predict_txb(const vp9b_frame_hdr *h, vp9ref_frame *f, const vp9b_txb *t) {
  const int tx_size = t->tx & 3, plane = (t->tx >> 4) & 3;
  const int mode = t->aux & 15, have_top = (t->aux >> 4) & 1, have_left = (t->aux >> 5) & 1,
            have_right = (t->aux >> 6) & 1;
  const int bs = 4 << tx_size, bd = h->bit_depth;
  const int ss_x = plane ? h->ss_x : 0, ss_y = plane ? h->ss_y : 0;
  const int frame_w = (int)(h->mi_cols * 8) >> ss_x, frame_h = (int)(h->mi_rows * 8) >> ss_y;
  const int x0 = t->x4 * 4, y0 = t->y4 * 4;
  const int stride = f->stride[plane];
  uint16_t *dst = f->plane[plane] + (size_t)y0 * stride + x0;
  uint16_t left[32], above_data[64 + 16];
  uint16_t *above = above_data + 16;
  const int base = 128 << (bd - 8);
  int i;
  /* left column */
  if (have_left) {
    for (i = 0; i < bs; i++) {
      const int yy = (y0 + i < frame_h) ? i : frame_h - y0 - 1;
      left[i] = dst[yy * stride - 1];
    }
  } else {
    for (i = 0; i < bs; i++) left[i] = (uint16_t)(base + 1);
  }
  /* above row + above right */
  if (have_top) {
    const uint16_t *ar = dst - stride;
    for (i = 0; i < 2 * bs; i++) {
      int avail_px; /* number of real samples readable from x0 */
      if (i < bs) avail_px = 1; else avail_px = (have_right && bs == 4);
      if (avail_px) {
        const int xx = (x0 + i < frame_w) ? i : frame_w - x0 - 1;
        above[i] = ar[xx];
      } else {
        above[i] = above[bs - 1];
      }
    }
    above[-1] = have_left ? ar[-1] : (uint16_t)(base + 1);
  } else {
    for (i = 0; i < 2 * bs; i++) above[i] = (uint16_t)(base - 1);
    above[-1] = (uint16_t)(base - 1);
  }
  vp9ref_intra_predictor(mode, bs, have_left, have_top, above, left, dst, stride, bd);
}

/* ------------------------------------------------------------------ loop filter ------------ */
static inline int sc_clamp(int t, int bd) { /* signed_char_clamp_high */
  const int lim = 128 << (bd - 8);
  return clampi(t, -lim, lim - 1);
}

/* highbd_filter4 (loopfilter.c:415-448) */
static void lf_filter4(int mask, int thresh, uint16_t *op1, uint16_t *op0, uint16_t *oq0, uint16_t *oq1, int bd) {
  const int shift = bd - 8, off = 0x80 << shift;
  const int ps1 = (int)*op1 - off, ps0 = (int)*op0 - off, qs0 = (int)*oq0 - off, qs1 = (int)*oq1 - off;
  const int t16 = thresh << shift;
  const int hev = (abs((int)*op1 - (int)*op0) > t16 || abs((int)*oq1 - (int)*oq0) > t16) ? -1 : 0;
  int filter = sc_clamp(ps1 - qs1, bd) & hev;
  filter = sc_clamp(filter + 3 * (qs0 - ps0), bd) & mask;
  const int filter1 = sc_clamp(filter + 4, bd) >> 3;
  const int filter2 = sc_clamp(filter + 3, bd) >> 3;
  *oq0 = (uint16_t)(sc_clamp(qs0 - filter1, bd) + off);
  *op0 = (uint16_t)(sc_clamp(ps0 + filter2, bd) + off);
  filter = ROUND_POW2(filter1, 1) & ~hev;
  *oq1 = (uint16_t)(sc_clamp(qs1 - filter, bd) + off);
  *op1 = (uint16_t)(sc_clamp(ps1 + filter, bd) + off);
}

/* One sample position across an edge.  s points at q0; step is the distance between taps.
 * size: 4, 8 or 16 (vpx_highbd_lpf_{horizontal,vertical}_{4,8,16}_c, loopfilter.c:450-743) */
static void lf_edge_px(uint16_t *s, int step, int size, int blimit, int limit, int thresh, int bd) {
  const int sh = bd - 8;
  const int p3 = s[-4 * step], p2 = s[-3 * step], p1 = s[-2 * step], p0 = s[-step];
  const int q0 = s[0], q1 = s[step], q2 = s[2 * step], q3 = s[3 * step];
  const int lim16 = limit << sh, blim16 = blimit << sh, one = 1 << sh;
  int mask = 0;
  mask |= abs(p3 - p2) > lim16; mask |= abs(p2 - p1) > lim16; mask |= abs(p1 - p0) > lim16;
  mask |= abs(q1 - q0) > lim16; mask |= abs(q2 - q1) > lim16; mask |= abs(q3 - q2) > lim16;
  mask |= abs(p0 - q0) * 2 + abs(p1 - q1) / 2 > blim16;
  mask = mask ? 0 : -1;
  if (size == 4) {
    lf_filter4(mask, thresh, s - 2 * step, s - step, s, s + step, bd);
    return;
  }
  const int flat = !(abs(p1 - p0) > one || abs(q1 - q0) > one || abs(p2 - p0) > one || abs(q2 - q0) > one ||
                     abs(p3 - p0) > one || abs(q3 - q0) > one);
  if (size == 16) {
    const int p7 = s[-8 * step], p6 = s[-7 * step], p5 = s[-6 * step], p4 = s[-5 * step];
    const int q4 = s[4 * step], q5 = s[5 * step], q6 = s[6 * step], q7 = s[7 * step];
    const int flat2 = !(abs(p4 - p0) > one || abs(q4 - q0) > one || abs(p5 - p0) > one || abs(q5 - q0) > one ||
                        abs(p6 - p0) > one || abs(q6 - q0) > one || abs(p7 - p0) > one || abs(q7 - q0) > one);
    if (flat2 && flat && mask) {
      s[-7 * step] = (uint16_t)ROUND_POW2(p7 * 7 + p6 * 2 + p5 + p4 + p3 + p2 + p1 + p0 + q0, 4);
      s[-6 * step] = (uint16_t)ROUND_POW2(p7 * 6 + p6 + p5 * 2 + p4 + p3 + p2 + p1 + p0 + q0 + q1, 4);
      s[-5 * step] = (uint16_t)ROUND_POW2(p7 * 5 + p6 + p5 + p4 * 2 + p3 + p2 + p1 + p0 + q0 + q1 + q2, 4);
      s[-4 * step] = (uint16_t)ROUND_POW2(p7 * 4 + p6 + p5 + p4 + p3 * 2 + p2 + p1 + p0 + q0 + q1 + q2 + q3, 4);
      s[-3 * step] = (uint16_t)ROUND_POW2(p7 * 3 + p6 + p5 + p4 + p3 + p2 * 2 + p1 + p0 + q0 + q1 + q2 + q3 + q4, 4);
      s[-2 * step] = (uint16_t)ROUND_POW2(p7 * 2 + p6 + p5 + p4 + p3 + p2 + p1 * 2 + p0 + q0 + q1 + q2 + q3 + q4 + q5, 4);
      s[-1 * step] = (uint16_t)ROUND_POW2(p7 + p6 + p5 + p4 + p3 + p2 + p1 + p0 * 2 + q0 + q1 + q2 + q3 + q4 + q5 + q6, 4);
      s[0 * step] = (uint16_t)ROUND_POW2(p6 + p5 + p4 + p3 + p2 + p1 + p0 + q0 * 2 + q1 + q2 + q3 + q4 + q5 + q6 + q7, 4);
      s[1 * step] = (uint16_t)ROUND_POW2(p5 + p4 + p3 + p2 + p1 + p0 + q0 + q1 * 2 + q2 + q3 + q4 + q5 + q6 + q7 * 2, 4);
      s[2 * step] = (uint16_t)ROUND_POW2(p4 + p3 + p2 + p1 + p0 + q0 + q1 + q2 * 2 + q3 + q4 + q5 + q6 + q7 * 3, 4);
      s[3 * step] = (uint16_t)ROUND_POW2(p3 + p2 + p1 + p0 + q0 + q1 + q2 + q3 * 2 + q4 + q5 + q6 + q7 * 4, 4);
      s[4 * step] = (uint16_t)ROUND_POW2(p2 + p1 + p0 + q0 + q1 + q2 + q3 + q4 * 2 + q5 + q6 + q7 * 5, 4);
      s[5 * step] = (uint16_t)ROUND_POW2(p1 + p0 + q0 + q1 + q2 + q3 + q4 + q5 * 2 + q6 + q7 * 6, 4);
      s[6 * step] = (uint16_t)ROUND_POW2(p0 + q0 + q1 + q2 + q3 + q4 + q5 + q6 * 2 + q7 * 7, 4);
      return;
    }
  }
  if (flat && mask) {
    s[-3 * step] = (uint16_t)ROUND_POW2(p3 + p3 + p3 + 2 * p2 + p1 + p0 + q0, 3);
    s[-2 * step] = (uint16_t)ROUND_POW2(p3 + p3 + p2 + 2 * p1 + p0 + q0 + q1, 3);
    s[-1 * step] = (uint16_t)ROUND_POW2(p3 + p2 + p1 + 2 * p0 + q0 + q1 + q2, 3);
    s[0 * step] = (uint16_t)ROUND_POW2(p2 + p1 + p0 + 2 * q0 + q1 + q2 + q3, 3);
    s[1 * step] = (uint16_t)ROUND_POW2(p1 + p0 + q0 + 2 * q1 + q2 + q3 + q3, 3);
    s[2 * step] = (uint16_t)ROUND_POW2(p0 + q0 + q1 + 2 * q2 + q3 + q3 + q3, 3);
  } else {
    lf_filter4(mask, thresh, s - 2 * step, s - step, s, s + step, bd);
  }
}

/* Filter `count` consecutive sample positions of one edge.  vertical edge: positions advance
 * along rows (pitch), taps along x (1); horizontal edge: the other way round. */
static void lf_edge(uint16_t *s, int pitch, int vertical, int size, int count, const vp9b_frame_hdr *h, int level) {
  const int blimit = h->lf_mblim[level], limit = h->lf_lim[level], thresh = level >> 4;
  for (int i = 0; i < count; i++) {
    if (vertical) lf_edge_px(s + i * pitch, 1, size, blimit, limit, thresh, h->bit_depth);
    else lf_edge_px(s + i, pitch, size, blimit, limit, thresh, h->bit_depth);
  }
}

/* highbd_filter_selectively_vert_row2 (vp9_loopfilter.c:375-451): two 8-row bands at once. */
static void filter_selectively_vert_row2(int ss, uint16_t *s, int pitch, unsigned m16, unsigned m8, unsigned m4,
                                         unsigned m4i, const vp9b_frame_hdr *h, const uint8_t *lfl) {
  const unsigned dual_mask_cutoff = ss ? 0xff : 0xffff;
  const int lfl_forward = ss ? 4 : 8;
  const unsigned dual_one = 1 | (1u << lfl_forward);
  unsigned mask;
  uint16_t *ss0 = s;
  for (mask = (m16 | m8 | m4 | m4i) & dual_mask_cutoff; mask; mask = (mask & ~dual_one) >> 1) {
    if (mask & dual_one) {
      const int l0 = lfl[0], l1 = lfl[lfl_forward];
      uint16_t *ss1 = ss0 + 8 * pitch;
      if (m16 & dual_one) {
        if ((m16 & dual_one) == dual_one) lf_edge(ss0, pitch, 1, 16, 16, h, l0); /* both bands use lfis[0] */
        else if (m16 & 1) lf_edge(ss0, pitch, 1, 16, 8, h, l0);
        else lf_edge(ss1, pitch, 1, 16, 8, h, l1);
      }
      if (m8 & dual_one) {
        if (m8 & 1) lf_edge(ss0, pitch, 1, 8, 8, h, l0);
        if (m8 & (1u << lfl_forward)) lf_edge(ss1, pitch, 1, 8, 8, h, l1);
      }
      if (m4 & dual_one) {
        if (m4 & 1) lf_edge(ss0, pitch, 1, 4, 8, h, l0);
        if (m4 & (1u << lfl_forward)) lf_edge(ss1, pitch, 1, 4, 8, h, l1);
      }
      if (m4i & dual_one) {
        if (m4i & 1) lf_edge(ss0 + 4, pitch, 1, 4, 8, h, l0);
        if (m4i & (1u << lfl_forward)) lf_edge(ss1 + 4, pitch, 1, 4, 8, h, l1);
      }
    }
    ss0 += 8;
    lfl += 1;
    m16 >>= 1; m8 >>= 1; m4 >>= 1; m4i >>= 1;
  }
}

/* highbd_filter_selectively_horiz (vp9_loopfilter.c:546-650) */
static void filter_selectively_horiz(uint16_t *s, int pitch, unsigned m16, unsigned m8, unsigned m4, unsigned m4i,
                                     const vp9b_frame_hdr *h, const uint8_t *lfl) {
  unsigned mask;
  int count;
  for (mask = m16 | m8 | m4 | m4i; mask; mask >>= count) {
    count = 1;
    if (mask & 1) {
      const int l0 = lfl[0];
      if (m16 & 1) {
        if ((m16 & 3) == 3) {
          lf_edge(s, pitch, 0, 16, 16, h, l0); /* dual: first block's thresholds for both */
          count = 2;
        } else {
          lf_edge(s, pitch, 0, 16, 8, h, l0);
        }
      } else if (m8 & 1) {
        if ((m8 & 3) == 3) {
          const int l1 = lfl[1];
          lf_edge(s, pitch, 0, 8, 8, h, l0);
          lf_edge(s + 8, pitch, 0, 8, 8, h, l1);
          if (m4i & 1) lf_edge(s + 4 * pitch, pitch, 0, 4, 8, h, l0);
          if (m4i & 2) lf_edge(s + 8 + 4 * pitch, pitch, 0, 4, 8, h, l1);
          count = 2;
        } else {
          lf_edge(s, pitch, 0, 8, 8, h, l0);
          if (m4i & 1) lf_edge(s + 4 * pitch, pitch, 0, 4, 8, h, l0);
        }
      } else if (m4 & 1) {
        if ((m4 & 3) == 3) {
          const int l1 = lfl[1];
          lf_edge(s, pitch, 0, 4, 8, h, l0);
          lf_edge(s + 8, pitch, 0, 4, 8, h, l1);
          if (m4i & 1) lf_edge(s + 4 * pitch, pitch, 0, 4, 8, h, l0);
          if (m4i & 2) lf_edge(s + 8 + 4 * pitch, pitch, 0, 4, 8, h, l1);
          count = 2;
        } else {
          lf_edge(s, pitch, 0, 4, 8, h, l0);
          if (m4i & 1) lf_edge(s + 4 * pitch, pitch, 0, 4, 8, h, l0);
        }
      } else {
        lf_edge(s + 4 * pitch, pitch, 0, 4, 8, h, l0);
      }
    }
    s += 8 * count;
    lfl += count;
    m16 >>= count; m8 >>= count; m4 >>= count; m4i >>= count;
  }
}

/* vp9_filter_block_plane_ss00 (vp9_loopfilter.c:1241-1324) */
static void filter_block_plane_ss00(const vp9b_frame_hdr *h, uint16_t *buf, int pitch, int mi_row, const vp9b_lfm *lfm) {
  uint64_t m16 = lfm->left_y[2], m8 = lfm->left_y[1], m4 = lfm->left_y[0], m4i = lfm->int_4x4_y;
  uint16_t *b = buf;
  int r;
  for (r = 0; r < 8 && mi_row + r < (int)h->mi_rows; r += 2) {
    filter_selectively_vert_row2(0, b, pitch, (unsigned)m16, (unsigned)m8, (unsigned)m4, (unsigned)m4i, h,
                                 &lfm->lfl_y[r << 3]);
    b += 16 * pitch;
    m16 >>= 16; m8 >>= 16; m4 >>= 16; m4i >>= 16;
  }
  b = buf;
  m16 = lfm->above_y[2]; m8 = lfm->above_y[1]; m4 = lfm->above_y[0]; m4i = lfm->int_4x4_y;
  for (r = 0; r < 8 && mi_row + r < (int)h->mi_rows; r++) {
    unsigned a16, a8, a4;
    if (mi_row + r == 0) { a16 = a8 = a4 = 0; }
    else { a16 = m16 & 0xff; a8 = m8 & 0xff; a4 = m4 & 0xff; }
    filter_selectively_horiz(b, pitch, a16, a8, a4, m4i & 0xff, h, &lfm->lfl_y[r << 3]);
    b += 8 * pitch;
    m16 >>= 8; m8 >>= 8; m4 >>= 8; m4i >>= 8;
  }
}

/* vp9_filter_block_plane_ss11 (vp9_loopfilter.c:1326-1422) */
static void filter_block_plane_ss11(const vp9b_frame_hdr *h, uint16_t *buf, int pitch, int mi_row, const vp9b_lfm *lfm) {
  uint16_t m16 = lfm->left_uv[2], m8 = lfm->left_uv[1], m4 = lfm->left_uv[0], m4i = lfm->int_4x4_uv;
  uint8_t lfl_uv[16];
  uint16_t *b = buf;
  int r, c;
  memset(lfl_uv, 0, sizeof(lfl_uv));
  for (r = 0; r < 8 && mi_row + r < (int)h->mi_rows; r += 4) {
    for (c = 0; c < 4; c++) {
      lfl_uv[(r << 1) + c] = lfm->lfl_y[(r << 3) + (c << 1)];
      lfl_uv[((r + 2) << 1) + c] = lfm->lfl_y[((r + 2) << 3) + (c << 1)];
    }
    filter_selectively_vert_row2(1, b, pitch, m16, m8, m4, m4i, h, &lfl_uv[r << 1]);
    b += 16 * pitch;
    m16 >>= 8; m8 >>= 8; m4 >>= 8; m4i >>= 8;
  }
  b = buf;
  m16 = lfm->above_uv[2]; m8 = lfm->above_uv[1]; m4 = lfm->above_uv[0]; m4i = lfm->int_4x4_uv;
  for (r = 0; r < 8 && mi_row + r < (int)h->mi_rows; r += 2) {
    const int skip_border_4x4_r = mi_row + r == (int)h->mi_rows - 1;
    const unsigned m4i_r = skip_border_4x4_r ? 0 : (m4i & 0xf);
    unsigned a16, a8, a4;
    if (mi_row + r == 0) { a16 = a8 = a4 = 0; }
    else { a16 = m16 & 0xf; a8 = m8 & 0xf; a4 = m4 & 0xf; }
    filter_selectively_horiz(b, pitch, a16, a8, a4, m4i_r, h, &lfl_uv[r << 1]);
    b += 8 * pitch;
    m16 >>= 4; m8 >>= 4; m4 >>= 4; m4i >>= 4;
  }
}

/* filter_selectively_vert (vp9_loopfilter.c:652-681, high bit depth :683-714): one 8-row band */
static void filter_selectively_vert(uint16_t *s, int pitch, unsigned m16, unsigned m8, unsigned m4, unsigned m4i,
                                    const vp9b_frame_hdr *h, const uint8_t *lfl) {
  unsigned mask;
  for (mask = m16 | m8 | m4 | m4i; mask; mask >>= 1) {
    if (mask & 1) {
      if (m16 & 1) lf_edge(s, pitch, 1, 16, 8, h, *lfl);
      else if (m8 & 1) lf_edge(s, pitch, 1, 8, 8, h, *lfl);
      else if (m4 & 1) lf_edge(s, pitch, 1, 4, 8, h, *lfl);
    }
    if (m4i & 1) lf_edge(s + 4, pitch, 1, 4, 8, h, *lfl);
    s += 8;
    lfl += 1;
    m16 >>= 1; m8 >>= 1; m4 >>= 1; m4i >>= 1;
  }
}

/* the two filter passes of vp9_filter_block_plane_non420 (vp9_loopfilter.c:1180-1239); the masks
 * were built by the packer (vp9b_lfm_uv), frame-edge rules included */
static void filter_block_plane_non420(const vp9b_frame_hdr *h, uint16_t *buf, int pitch, int mi_row, const vp9b_lfm_uv *m) {
  const int row_step = 1 << h->ss_y;
  uint16_t *b = buf;
  int r;
  for (r = 0; r < 8 && mi_row + r < (int)h->mi_rows; r += row_step) {
    filter_selectively_vert(b, pitch, m->v16[r], m->v8[r], m->v4[r], m->v4i[r], h, &m->lfl[r << 3]);
    b += 8 * pitch;
  }
  b = buf;
  for (r = 0; r < 8 && mi_row + r < (int)h->mi_rows; r += row_step) {
    filter_selectively_horiz(b, pitch, m->h16[r], m->h8[r], m->h4[r], m->h4i[r], h, &m->lfl[r << 3]);
    b += 8 * pitch;
  }
}

/* loop_filter_rows (vp9_loopfilter.c:1424-1469): superblocks in raster order, masks already
 * adjusted (vp9_adjust_mask) by the packer. */
static int loop_filter_frame(vp9ref_ctx *c, const vp9b_frame_hdr *h, const uint8_t *base) {
  if (!h->lf_level) return 0;
  const int non420 = h->ss_x != h->ss_y;   /* LF_PATH_SLOW: luma ss00, chroma non420 */
  if (non420 && h->sec[VP9B_SEC_LFM_UV].count != h->sb_rows * h->sb_cols) return VP9B_ERR_INVALID;
  vp9ref_frame *f = &c->slot[h->dst_slot];
  const vp9b_lfm *lfm = (const vp9b_lfm *)(base + h->sec[VP9B_SEC_LFM].offset);
  const vp9b_lfm_uv *lfm_uv = (const vp9b_lfm_uv *)(base + h->sec[VP9B_SEC_LFM_UV].offset);
  for (uint32_t sbr = 0; sbr < h->sb_rows; sbr++)
    for (uint32_t sbc = 0; sbc < h->sb_cols; sbc++) {
      const vp9b_lfm *m = &lfm[sbr * h->sb_cols + sbc];
      filter_block_plane_ss00(h, f->plane[0] + (size_t)sbr * 64 * f->stride[0] + sbc * 64, f->stride[0], sbr * 8, m);
      for (int p = 1; p < 3; p++) {
        if (non420)
          filter_block_plane_non420(h, f->plane[p] + (size_t)sbr * (64 >> h->ss_y) * f->stride[p] + sbc * (64 >> h->ss_x),
                                    f->stride[p], sbr * 8, &lfm_uv[sbr * h->sb_cols + sbc]);
        else if (h->ss_x)
          filter_block_plane_ss11(h, f->plane[p] + (size_t)sbr * 32 * f->stride[p] + sbc * 32, f->stride[p], sbr * 8, m);
        else
          filter_block_plane_ss00(h, f->plane[p] + (size_t)sbr * 64 * f->stride[p] + sbc * 64, f->stride[p], sbr * 8, m);
      }
    }
  return 0;
}

/* ------------------------------------------------------------------ frame driver ----------- */
/* stop_stage: 0 = everything; 1 after inter prediction; 2 after inter residual; 3 after intra;
 * 4 after loop filter. */
int vp9ref_recon_frame(vp9ref_ctx *c, const void *packed, size_t bytes, int stop_stage) {
  const uint8_t *base = (const uint8_t *)packed;
  const vp9b_frame_hdr *h = (const vp9b_frame_hdr *)packed;
  if (!c || !packed || bytes < sizeof(*h) || h->magic != VP9B_MAGIC || h->total_bytes > bytes) return VP9B_ERR_INVALID;
  if (h->dst_slot < 0 || h->dst_slot >= VP9B_NUM_SLOTS) return VP9B_ERR_INVALID;
  vp9ref_frame *f = &c->slot[h->dst_slot];
  int e = setup_slot(f, h->width, h->height, h->ss_x, h->ss_y, h->bit_depth);
  if (e) return e;
  if (!h->intra_only) {
    e = inter_predict_frame(c, h, base);
    if (e) return e;
    if (stop_stage == 1) return 0;
    const vp9b_txb *t = (const vp9b_txb *)(base + h->sec[VP9B_SEC_TXB_INTER].offset);
    for (uint32_t i = 0; i < h->sec[VP9B_SEC_TXB_INTER].count; i++) residual_add(h, base, &t[i], f);
  }
  if (stop_stage == 1 || stop_stage == 2) return 0;
  for (int p = 0; p < 3; p++) {
    const vp9b_txb *t = (const vp9b_txb *)(base + h->sec[VP9B_SEC_TXB_INTRA_Y + p].offset);
    for (uint32_t i = 0; i < h->sec[VP9B_SEC_TXB_INTRA_Y + p].count; i++) {
      intra_predict_txb(h, f, &t[i]);
      residual_add(h, base, &t[i], f);
    }
  }
  if (stop_stage == 3) return 0;
  return loop_filter_frame(c, h, base);
}

/* Copy the cropped planes out: 1 byte/sample for 8-bit, 2 bytes little endian otherwise. */
int vp9ref_get_frame(vp9ref_ctx *c, int slot, uint8_t *y, uint8_t *u, uint8_t *v, int y_stride, int uv_stride) {
  if (!c || slot < 0 || slot >= VP9B_NUM_SLOTS || !c->slot[slot].valid) return VP9B_ERR_INVALID;
  const vp9ref_frame *f = &c->slot[slot];
  uint8_t *out[3] = { y, u, v };
  for (int p = 0; p < 3; p++) {
    const int w = p ? (int)((f->width + f->ss_x) >> f->ss_x) : (int)f->width;
    const int hh = p ? (int)((f->height + f->ss_y) >> f->ss_y) : (int)f->height;
    const int ostride = p ? uv_stride : y_stride;
    for (int r = 0; r < hh; r++) {
      const uint16_t *s = f->plane[p] + (size_t)r * f->stride[p];
      if (f->bit_depth == 8) {
        for (int x = 0; x < w; x++) out[p][(size_t)r * ostride + x] = (uint8_t)s[x];
      } else {
        memcpy(out[p] + (size_t)r * ostride, s, (size_t)w * 2);
      }
    }
  }
  return 0;
}

int vp9ref_slot_info(vp9ref_ctx *c, int slot, uint32_t *w, uint32_t *h, int *bd, int *ss_x, int *ss_y) {
  if (!c || slot < 0 || slot >= VP9B_NUM_SLOTS || !c->slot[slot].valid) return VP9B_ERR_INVALID;
  *w = c->slot[slot].width; *h = c->slot[slot].height; *bd = c->slot[slot].bit_depth;
  *ss_x = c->slot[slot].ss_x; *ss_y = c->slot[slot].ss_y;
  return 0;
}

/* Test hook: (re)initialise a slot with given samples (8-bit: 1 byte/sample, else 2 LE). */
int vp9ref_set_frame(vp9ref_ctx *c, int slot, uint32_t w, uint32_t h, int bd, int ss_x, int ss_y, const uint8_t *y,
                     const uint8_t *u, const uint8_t *v, int y_stride, int uv_stride) {
  if (!c || slot < 0 || slot >= VP9B_NUM_SLOTS) return VP9B_ERR_INVALID;
  vp9ref_frame *f = &c->slot[slot];
  int e = setup_slot(f, w, h, ss_x, ss_y, bd);
  if (e) return e;
  const uint8_t *in[3] = { y, u, v };
  for (int p = 0; p < 3; p++) {
    const int pw = p ? (int)((w + ss_x) >> ss_x) : (int)w, ph = p ? (int)((h + ss_y) >> ss_y) : (int)h;
    const int istride = p ? uv_stride : y_stride;
    memset(f->plane[p], 0, (size_t)f->stride[p] * f->alloc_h[p] * sizeof(uint16_t));
    for (int r = 0; r < ph; r++)
      for (int x = 0; x < pw; x++)
        f->plane[p][(size_t)r * f->stride[p] + x] =
            bd == 8 ? in[p][(size_t)r * istride + x] : ((const uint16_t *)(in[p] + (size_t)r * istride))[x];
  }
  return 0;
}

/* Unit-test hooks: one loop-filter edge of 8 positions, and a w x h convolution. */
void vp9ref_lpf_edge8(uint16_t *s, int pitch, int vertical, int size, int blimit, int limit, int thresh, int bd) {
  for (int i = 0; i < 8; i++) {
    if (vertical) lf_edge_px(s + i * pitch, 1, size, blimit, limit, thresh, bd);
    else lf_edge_px(s + i, pitch, size, blimit, limit, thresh, bd);
  }
}

/* Test hook: the raw sample array of a plane (aligned size + padding), for stage-level checks
 * that need the samples between the crop size and the aligned size. */
uint16_t *vp9ref_plane(vp9ref_ctx *c, int slot, int plane, int *stride, int *rows) {
  if (!c || slot < 0 || slot >= VP9B_NUM_SLOTS || plane < 0 || plane > 2 || !c->slot[slot].valid) return NULL;
  *stride = c->slot[slot].stride[plane];
  *rows = c->slot[slot].alloc_h[plane];
  return c->slot[slot].plane[plane];
}
