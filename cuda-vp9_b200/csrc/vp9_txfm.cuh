// Inverse transforms for sm_100a: 1-D integer butterflies evaluated entirely in registers, one
// lane per row / column of a transform block.  Written as a recursive even/odd decomposition
// (idct2N = idctN on the even inputs + an odd-half rotation network), which is the structure of
// the VP9 inverse DCT (reference arithmetic: vpx_dsp/inv_txfm.c:133-153 idct4, :267-324 idct8,
// :556-722 idct16, :816-1185 idct32, :94-131 iadst4, :189-265 iadst8, :385-554 iadst16,
// :18-66 iwht).  Every rotation rounds with (x + 2^13) >> 14 exactly where libvpx does.
//
// PT is the product type: int for 8-bit streams (|coef| < 2^15 so products fit 32 bits),
// long long for 10/12-bit.  No tensor cores: this is not a dense contraction and must be
// bit-exact.
#pragma once
#include <stdint.h>

namespace vp9b {

#define VP9B_DEV __device__ __forceinline__

template <typename PT>
VP9B_DEV int rs14(PT v) {
  return (int)((v + (PT)8192) >> 14);
}

// cos(k*pi/64) * 2^14
__device__ __constant__ const int kCospi[32] = { 16384, 16364, 16305, 16207, 16069, 15893, 15679, 15426,
                                                 15137, 14811, 14449, 14053, 13623, 13160, 12665, 12140,
                                                 11585, 11003, 10394, 9760,  9102,  8423,  7723,  7005,
                                                 6270,  5520,  4756,  3981,  3196,  2404,  1606,  804 };
#define CP(k) ((PT)c##k)

// rotation pair:  lo = R(a*ca - b*cb),  hi = R(a*cb + b*ca)
template <typename PT>
VP9B_DEV void rot(int a, int b, int ca, int cb, int &lo, int &hi) {
  lo = rs14<PT>((PT)a * ca - (PT)b * cb);
  hi = rs14<PT>((PT)a * cb + (PT)b * ca);
}

enum : int {
  c1 = 16364, c2 = 16305, c3 = 16207, c4 = 16069, c5 = 15893, c6 = 15679, c7 = 15426, c8 = 15137,
  c9 = 14811, c10 = 14449, c11 = 14053, c12 = 13623, c13 = 13160, c14 = 12665, c15 = 12140, c16 = 11585,
  c17 = 11003, c18 = 10394, c19 = 9760, c20 = 9102, c21 = 8423, c22 = 7723, c23 = 7005, c24 = 6270,
  c25 = 5520, c26 = 4756, c27 = 3981, c28 = 3196, c29 = 2404, c30 = 1606, c31 = 804
};

// ---- inverse DCT, recursive ------------------------------------------------------------
template <typename PT>
VP9B_DEV void idct4(int i0, int i1, int i2, int i3, int *o) {
  const int s0 = rs14<PT>((PT)(i0 + i2) * c16), s1 = rs14<PT>((PT)(i0 - i2) * c16);
  int s2, s3;
  rot<PT>(i1, i3, c24, c8, s2, s3);
  o[0] = s0 + s3;
  o[1] = s1 + s2;
  o[2] = s1 - s2;
  o[3] = s0 - s3;
}

template <typename PT>
VP9B_DEV void idct8(const int *in, int *o) {  // in: 8 values in natural order
  int e[4];
  idct4<PT>(in[0], in[2], in[4], in[6], e);
  int s4, s7, s5, s6;
  rot<PT>(in[1], in[7], c28, c4, s4, s7);
  rot<PT>(in[5], in[3], c12, c20, s5, s6);
  const int t4 = s4 + s5, t5 = s4 - s5, t6 = s7 - s6, t7 = s6 + s7;
  const int u5 = rs14<PT>((PT)(t6 - t5) * c16), u6 = rs14<PT>((PT)(t5 + t6) * c16);
  o[0] = e[0] + t7; o[1] = e[1] + u6; o[2] = e[2] + u5; o[3] = e[3] + t4;
  o[4] = e[3] - t4; o[5] = e[2] - u5; o[6] = e[1] - u6; o[7] = e[0] - t7;
}

// odd half of the 16-point transform: inputs in[1],in[3],..,in[15] -> d[0..7] (= step2[8..15])
template <typename PT>
VP9B_DEV void idct16_odd(int i1, int i3, int i5, int i7, int i9, int i11, int i13, int i15, int *d) {
  int a8, a15, a9, a14, a10, a13, a11, a12;
  rot<PT>(i1, i15, c30, c2, a8, a15);
  rot<PT>(i9, i7, c14, c18, a9, a14);
  rot<PT>(i5, i11, c22, c10, a10, a13);
  rot<PT>(i13, i3, c6, c26, a11, a12);
  // stage 3
  const int b8 = a8 + a9, b9 = a8 - a9, b10 = a11 - a10, b11 = a10 + a11;
  const int b12 = a12 + a13, b13 = a12 - a13, b14 = a15 - a14, b15 = a14 + a15;
  // stage 4
  const int e9 = rs14<PT>(-(PT)b9 * c8 + (PT)b14 * c24), e14 = rs14<PT>((PT)b9 * c24 + (PT)b14 * c8);
  const int e10 = rs14<PT>(-(PT)b10 * c24 - (PT)b13 * c8), e13 = rs14<PT>(-(PT)b10 * c8 + (PT)b13 * c24);
  // stage 5
  const int f8 = b8 + b11, f9 = e9 + e10, f10 = e9 - e10, f11 = b8 - b11;
  const int f12 = b15 - b12, f13 = e14 - e13, f14 = e13 + e14, f15 = b12 + b15;
  // stage 6
  d[0] = f8;
  d[1] = f9;
  d[2] = rs14<PT>((PT)(f13 - f10) * c16);
  d[5] = rs14<PT>((PT)(f10 + f13) * c16);
  d[3] = rs14<PT>((PT)(f12 - f11) * c16);
  d[4] = rs14<PT>((PT)(f11 + f12) * c16);
  d[6] = f14;
  d[7] = f15;
}

template <typename PT>
VP9B_DEV void idct16(const int *in, int *o) {
  int ev[8], e[8], d[8];
#pragma unroll
  for (int i = 0; i < 8; i++) ev[i] = in[2 * i];
  idct8<PT>(ev, e);
  idct16_odd<PT>(in[1], in[3], in[5], in[7], in[9], in[11], in[13], in[15], d);
#pragma unroll
  for (int i = 0; i < 8; i++) {
    o[i] = e[i] + d[7 - i];
    o[15 - i] = e[i] - d[7 - i];
  }
}

template <typename PT>
VP9B_DEV void idct32(const int *in, int *o) {
  int ev[16], e[16];
#pragma unroll
  for (int i = 0; i < 16; i++) ev[i] = in[2 * i];
  idct16<PT>(ev, e);
  // odd half: step1[16..31]
  int s16, s31, s17, s30, s18, s29, s19, s28, s20, s27, s21, s26, s22, s25, s23, s24;
  rot<PT>(in[1], in[31], c31, c1, s16, s31);
  rot<PT>(in[17], in[15], c15, c17, s17, s30);
  rot<PT>(in[9], in[23], c23, c9, s18, s29);
  rot<PT>(in[25], in[7], c7, c25, s19, s28);
  rot<PT>(in[5], in[27], c27, c5, s20, s27);
  rot<PT>(in[21], in[11], c11, c21, s21, s26);
  rot<PT>(in[13], in[19], c19, c13, s22, s25);
  rot<PT>(in[29], in[3], c3, c29, s23, s24);
  // stage 2
  const int a16 = s16 + s17, a17 = s16 - s17, a18 = s19 - s18, a19 = s18 + s19;
  const int a20 = s20 + s21, a21 = s20 - s21, a22 = s23 - s22, a23 = s22 + s23;
  const int a24 = s24 + s25, a25 = s24 - s25, a26 = s27 - s26, a27 = s26 + s27;
  const int a28 = s28 + s29, a29 = s28 - s29, a30 = s31 - s30, a31 = s30 + s31;
  // stage 3
  const int b17 = rs14<PT>(-(PT)a17 * c4 + (PT)a30 * c28), b30 = rs14<PT>((PT)a17 * c28 + (PT)a30 * c4);
  const int b18 = rs14<PT>(-(PT)a18 * c28 - (PT)a29 * c4), b29 = rs14<PT>(-(PT)a18 * c4 + (PT)a29 * c28);
  const int b21 = rs14<PT>(-(PT)a21 * c20 + (PT)a26 * c12), b26 = rs14<PT>((PT)a21 * c12 + (PT)a26 * c20);
  const int b22 = rs14<PT>(-(PT)a22 * c12 - (PT)a25 * c20), b25 = rs14<PT>(-(PT)a22 * c20 + (PT)a25 * c12);
  // stage 4
  const int d16 = a16 + a19, d17 = b17 + b18, d18 = b17 - b18, d19 = a16 - a19;
  const int d20 = a23 - a20, d21 = b22 - b21, d22 = b21 + b22, d23 = a20 + a23;
  const int d24 = a24 + a27, d25 = b25 + b26, d26 = b25 - b26, d27 = a24 - a27;
  const int d28 = a31 - a28, d29 = b30 - b29, d30 = b29 + b30, d31 = a28 + a31;
  // stage 5
  const int g18 = rs14<PT>(-(PT)d18 * c8 + (PT)d29 * c24), g29 = rs14<PT>((PT)d18 * c24 + (PT)d29 * c8);
  const int g19 = rs14<PT>(-(PT)d19 * c8 + (PT)d28 * c24), g28 = rs14<PT>((PT)d19 * c24 + (PT)d28 * c8);
  const int g20 = rs14<PT>(-(PT)d20 * c24 - (PT)d27 * c8), g27 = rs14<PT>(-(PT)d20 * c8 + (PT)d27 * c24);
  const int g21 = rs14<PT>(-(PT)d21 * c24 - (PT)d26 * c8), g26 = rs14<PT>(-(PT)d21 * c8 + (PT)d26 * c24);
  // stage 6
  const int h16 = d16 + d23, h17 = d17 + d22, h18 = g18 + g21, h19 = g19 + g20;
  const int h20 = g19 - g20, h21 = g18 - g21, h22 = d17 - d22, h23 = d16 - d23;
  const int h24 = d31 - d24, h25 = d30 - d25, h26 = g29 - g26, h27 = g28 - g27;
  const int h28 = g27 + g28, h29 = g26 + g29, h30 = d25 + d30, h31 = d24 + d31;
  // stage 7
  int q[16];
  q[0] = h16; q[1] = h17; q[2] = h18; q[3] = h19;
  q[4] = rs14<PT>((PT)(h27 - h20) * c16);  q[11] = rs14<PT>((PT)(h20 + h27) * c16);
  q[5] = rs14<PT>((PT)(h26 - h21) * c16);  q[10] = rs14<PT>((PT)(h21 + h26) * c16);
  q[6] = rs14<PT>((PT)(h25 - h22) * c16);  q[9] = rs14<PT>((PT)(h22 + h25) * c16);
  q[7] = rs14<PT>((PT)(h24 - h23) * c16);  q[8] = rs14<PT>((PT)(h23 + h24) * c16);
  q[12] = h28; q[13] = h29; q[14] = h30; q[15] = h31;
#pragma unroll
  for (int i = 0; i < 16; i++) {
    o[i] = e[i] + q[15 - i];
    o[31 - i] = e[i] - q[15 - i];
  }
}

// ---- inverse ADST --------------------------------------------------------------------------
template <typename PT>
VP9B_DEV void iadst4(const int *in, int *o) {
  const int x0 = in[0], x1 = in[1], x2 = in[2], x3 = in[3];
  if (!(x0 | x1 | x2 | x3)) { o[0] = o[1] = o[2] = o[3] = 0; return; }
  const PT sp1 = 5283, sp2 = 9929, sp3 = 13377, sp4 = 15212;
  PT s0 = sp1 * x0, s1 = sp2 * x0, s2 = sp3 * x1, s3 = sp4 * x2;
  const PT s4 = sp1 * x2, s5 = sp2 * x3, s6 = sp4 * x3;
  const PT s7 = (PT)x0 - x2 + x3;
  s0 = s0 + s3 + s5;
  s1 = s1 - s4 - s6;
  s3 = s2;
  s2 = sp3 * s7;
  o[0] = rs14<PT>(s0 + s3);
  o[1] = rs14<PT>(s1 + s3);
  o[2] = rs14<PT>(s2);
  o[3] = rs14<PT>(s0 + s1 - s3);
}

template <typename PT>
VP9B_DEV void iadst8(const int *in, int *o) {
  int x0 = in[7], x1 = in[0], x2 = in[5], x3 = in[2], x4 = in[3], x5 = in[4], x6 = in[1], x7 = in[6];
  if (!(x0 | x1 | x2 | x3 | x4 | x5 | x6 | x7)) {
#pragma unroll
    for (int i = 0; i < 8; i++) o[i] = 0;
    return;
  }
  PT s0 = (PT)c2 * x0 + (PT)c30 * x1, s1 = (PT)c30 * x0 - (PT)c2 * x1;
  PT s2 = (PT)c10 * x2 + (PT)c22 * x3, s3 = (PT)c22 * x2 - (PT)c10 * x3;
  PT s4 = (PT)c18 * x4 + (PT)c14 * x5, s5 = (PT)c14 * x4 - (PT)c18 * x5;
  PT s6 = (PT)c26 * x6 + (PT)c6 * x7, s7 = (PT)c6 * x6 - (PT)c26 * x7;
  x0 = rs14<PT>(s0 + s4); x1 = rs14<PT>(s1 + s5); x2 = rs14<PT>(s2 + s6); x3 = rs14<PT>(s3 + s7);
  x4 = rs14<PT>(s0 - s4); x5 = rs14<PT>(s1 - s5); x6 = rs14<PT>(s2 - s6); x7 = rs14<PT>(s3 - s7);
  const int t0 = x0, t1 = x1, t2 = x2, t3 = x3;
  s4 = (PT)c8 * x4 + (PT)c24 * x5;
  s5 = (PT)c24 * x4 - (PT)c8 * x5;
  s6 = -(PT)c24 * x6 + (PT)c8 * x7;
  s7 = (PT)c8 * x6 + (PT)c24 * x7;
  x0 = t0 + t2; x1 = t1 + t3; x2 = t0 - t2; x3 = t1 - t3;
  x4 = rs14<PT>(s4 + s6); x5 = rs14<PT>(s5 + s7); x6 = rs14<PT>(s4 - s6); x7 = rs14<PT>(s5 - s7);
  const int y2 = rs14<PT>((PT)c16 * (x2 + x3)), y3 = rs14<PT>((PT)c16 * (x2 - x3));
  const int y6 = rs14<PT>((PT)c16 * (x6 + x7)), y7 = rs14<PT>((PT)c16 * (x6 - x7));
  o[0] = x0; o[1] = -x4; o[2] = y6; o[3] = -y2; o[4] = y3; o[5] = -y7; o[6] = x5; o[7] = -x1;
}

template <typename PT>
VP9B_DEV void iadst16(const int *in, int *o) {
  int x[16];
  x[0] = in[15]; x[1] = in[0]; x[2] = in[13]; x[3] = in[2]; x[4] = in[11]; x[5] = in[4]; x[6] = in[9]; x[7] = in[6];
  x[8] = in[7]; x[9] = in[8]; x[10] = in[5]; x[11] = in[10]; x[12] = in[3]; x[13] = in[12]; x[14] = in[1]; x[15] = in[14];
  int any = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) any |= x[i];
  if (!any) {
#pragma unroll
    for (int i = 0; i < 16; i++) o[i] = 0;
    return;
  }
  PT s[16];
  // stage 1
  s[0] = (PT)x[0] * c1 + (PT)x[1] * c31;   s[1] = (PT)x[0] * c31 - (PT)x[1] * c1;
  s[2] = (PT)x[2] * c5 + (PT)x[3] * c27;   s[3] = (PT)x[2] * c27 - (PT)x[3] * c5;
  s[4] = (PT)x[4] * c9 + (PT)x[5] * c23;   s[5] = (PT)x[4] * c23 - (PT)x[5] * c9;
  s[6] = (PT)x[6] * c13 + (PT)x[7] * c19;  s[7] = (PT)x[6] * c19 - (PT)x[7] * c13;
  s[8] = (PT)x[8] * c17 + (PT)x[9] * c15;  s[9] = (PT)x[8] * c15 - (PT)x[9] * c17;
  s[10] = (PT)x[10] * c21 + (PT)x[11] * c11; s[11] = (PT)x[10] * c11 - (PT)x[11] * c21;
  s[12] = (PT)x[12] * c25 + (PT)x[13] * c7;  s[13] = (PT)x[12] * c7 - (PT)x[13] * c25;
  s[14] = (PT)x[14] * c29 + (PT)x[15] * c3;  s[15] = (PT)x[14] * c3 - (PT)x[15] * c29;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    x[i] = rs14<PT>(s[i] + s[i + 8]);
    x[i + 8] = rs14<PT>(s[i] - s[i + 8]);
  }
  // stage 2
  s[8] = (PT)x[8] * c4 + (PT)x[9] * c28;     s[9] = (PT)x[8] * c28 - (PT)x[9] * c4;
  s[10] = (PT)x[10] * c20 + (PT)x[11] * c12; s[11] = (PT)x[10] * c12 - (PT)x[11] * c20;
  s[12] = -(PT)x[12] * c28 + (PT)x[13] * c4; s[13] = (PT)x[12] * c4 + (PT)x[13] * c28;
  s[14] = -(PT)x[14] * c12 + (PT)x[15] * c20; s[15] = (PT)x[14] * c20 + (PT)x[15] * c12;
  int y[16];
#pragma unroll
  for (int i = 0; i < 4; i++) {
    y[i] = x[i] + x[i + 4];
    y[i + 4] = x[i] - x[i + 4];
    y[i + 8] = rs14<PT>(s[i + 8] + s[i + 12]);
    y[i + 12] = rs14<PT>(s[i + 8] - s[i + 12]);
  }
  // stage 3
  s[4] = (PT)y[4] * c8 + (PT)y[5] * c24;    s[5] = (PT)y[4] * c24 - (PT)y[5] * c8;
  s[6] = -(PT)y[6] * c24 + (PT)y[7] * c8;   s[7] = (PT)y[6] * c8 + (PT)y[7] * c24;
  s[12] = (PT)y[12] * c8 + (PT)y[13] * c24;  s[13] = (PT)y[12] * c24 - (PT)y[13] * c8;
  s[14] = -(PT)y[14] * c24 + (PT)y[15] * c8; s[15] = (PT)y[14] * c8 + (PT)y[15] * c24;
  x[0] = y[0] + y[2]; x[1] = y[1] + y[3]; x[2] = y[0] - y[2]; x[3] = y[1] - y[3];
  x[4] = rs14<PT>(s[4] + s[6]); x[5] = rs14<PT>(s[5] + s[7]); x[6] = rs14<PT>(s[4] - s[6]); x[7] = rs14<PT>(s[5] - s[7]);
  x[8] = y[8] + y[10]; x[9] = y[9] + y[11]; x[10] = y[8] - y[10]; x[11] = y[9] - y[11];
  x[12] = rs14<PT>(s[12] + s[14]); x[13] = rs14<PT>(s[13] + s[15]);
  x[14] = rs14<PT>(s[12] - s[14]); x[15] = rs14<PT>(s[13] - s[15]);
  // stage 4
  const int z2 = rs14<PT>(-(PT)c16 * (x[2] + x[3])), z3 = rs14<PT>((PT)c16 * (x[2] - x[3]));
  const int z6 = rs14<PT>((PT)c16 * (x[6] + x[7])), z7 = rs14<PT>((PT)c16 * (x[7] - x[6]));
  const int z10 = rs14<PT>((PT)c16 * (x[10] + x[11])), z11 = rs14<PT>((PT)c16 * (x[11] - x[10]));
  const int z14 = rs14<PT>(-(PT)c16 * (x[14] + x[15])), z15 = rs14<PT>((PT)c16 * (x[14] - x[15]));
  o[0] = x[0]; o[1] = -x[8]; o[2] = x[12]; o[3] = -x[4]; o[4] = z6; o[5] = z14; o[6] = z10; o[7] = z2;
  o[8] = z3; o[9] = z11; o[10] = z15; o[11] = z7; o[12] = x[5]; o[13] = -x[13]; o[14] = x[9]; o[15] = -x[1];
}

// 1-D dispatch: N in {4,8,16,32}; adst selects the ADST (N <= 16)
template <typename PT, int N>
VP9B_DEV void txfm1d(const int *in, int *out, bool adst) {
  if (N == 4) {
    if (adst) iadst4<PT>(in, out); else idct4<PT>(in[0], in[1], in[2], in[3], out);
  } else if (N == 8) {
    if (adst) iadst8<PT>(in, out); else idct8<PT>(in, out);
  } else if (N == 16) {
    if (adst) iadst16<PT>(in, out); else idct16<PT>(in, out);
  } else {
    idct32<PT>(in, out);
  }
}

// eob classes of the DCT (vp9_idct.c:119-204: vpx_idct8x8_12, idct16x16_10 / _38, idct32x32_34 / _135):
// with the default scan the first eob coefficients of a block lie in its top-left K x K corner, so
// both passes only see K leading inputs.  eob_lead = that K (N when the block is not DCT_DCT or is
// dense).  The reduced transforms are the full butterflies with literal zeros for the inputs K..N-1:
// the compiler folds away every operation that only sees zeros, the surviving arithmetic and its
// rounding are those of the full transform.
template <int N>
VP9B_DEV int eob_lead(int tx_type, int eob) {
  if (tx_type != 0) return N;
  if (N == 8) return eob <= 12 ? 4 : 8;
  if (N == 16) return eob <= 10 ? 4 : (eob <= 38 ? 8 : 16);
  if (N == 32) return eob <= 34 ? 8 : (eob <= 135 ? 16 : 32);
  return N;
}
// one line (row or column) of a block held in shared memory: element j at tb[j * sj]; K leading
// entries are read, the rest are zero
template <typename PT, int N, int K>
VP9B_DEV void itx_line(const int *tb, int sj, bool adst, int *out) {
  int in[N];
#pragma unroll
  for (int j = 0; j < N; j++) in[j] = j < K ? tb[j * sj] : 0;
  txfm1d<PT, N>(in, out, K < N ? false : adst);
}
// dispatch on the (warp-uniform) number of leading entries
template <typename PT, int N>
VP9B_DEV void itx_line_lead(const int *tb, int sj, bool adst, int lead, int *out) {
  if (N >= 8 && lead <= N / 4) itx_line<PT, N, (N >= 8 ? N / 4 : N)>(tb, sj, adst, out);
  else if (N >= 8 && lead <= N / 2) itx_line<PT, N, (N >= 8 ? N / 2 : N)>(tb, sj, adst, out);
  else itx_line<PT, N, N>(tb, sj, adst, out);
}

// one 1-D pass of the 4-point inverse Walsh-Hadamard transform (lossless path)
VP9B_DEV void iwht4(int a1, int c1_, int d1, int b1, int *o) {
  a1 += c1_;
  d1 -= b1;
  const int e1 = (a1 - d1) >> 1;
  b1 = e1 - b1;
  c1_ = e1 - c1_;
  a1 -= b1;
  d1 += c1_;
  o[0] = a1; o[1] = b1; o[2] = c1_; o[3] = d1;
}

}  // namespace vp9b
