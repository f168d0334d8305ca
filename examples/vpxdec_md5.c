/* vpxdec --md5 on top of the C ABI (include/vp9_b200.h), no Python involved:
 *     vpxdec_md5 clip.ivf            one MD5 per shown frame, same bytes hashed as vpxdec.c:285-302
 * Build:  gcc -O2 -Iinclude examples/vpxdec_md5.c -Lcuda-vp9_b200 -lvp9b200 -o vpxdec_md5
 * This is the whole-decoder binding of INTEGRATION.md section A. */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "vp9_b200.h"

/* ---- MD5 (RFC 1321) ---- */
typedef struct { uint32_t a, b, c, d; uint64_t len; uint8_t buf[64]; size_t fill; } md5_t;
static const uint32_t K[64] = {
  0xd76aa478, 0xe8c7b756, 0x242070db, 0xc1bdceee, 0xf57c0faf, 0x4787c62a, 0xa8304613, 0xfd469501, 0x698098d8, 0x8b44f7af,
  0xffff5bb1, 0x895cd7be, 0x6b901122, 0xfd987193, 0xa679438e, 0x49b40821, 0xf61e2562, 0xc040b340, 0x265e5a51, 0xe9b6c7aa,
  0xd62f105d, 0x02441453, 0xd8a1e681, 0xe7d3fbc8, 0x21e1cde6, 0xc33707d6, 0xf4d50d87, 0x455a14ed, 0xa9e3e905, 0xfcefa3f8,
  0x676f02d9, 0x8d2a4c8a, 0xfffa3942, 0x8771f681, 0x6d9d6122, 0xfde5380c, 0xa4beea44, 0x4bdecfa9, 0xf6bb4b60, 0xbebfbc70,
  0x289b7ec6, 0xeaa127fa, 0xd4ef3085, 0x04881d05, 0xd9d4d039, 0xe6db99e5, 0x1fa27cf8, 0xc4ac5665, 0xf4292244, 0x432aff97,
  0xab9423a7, 0xfc93a039, 0x655b59c3, 0x8f0ccc92, 0xffeff47d, 0x85845dd1, 0x6fa87e4f, 0xfe2ce6e0, 0xa3014314, 0x4e0811a1,
  0xf7537e82, 0xbd3af235, 0x2ad7d2bb, 0xeb86d391 };
static const uint8_t S[64] = { 7, 12, 17, 22, 7, 12, 17, 22, 7, 12, 17, 22, 7, 12, 17, 22, 5, 9, 14, 20, 5, 9, 14, 20,
                               5, 9, 14, 20, 5, 9, 14, 20, 4, 11, 16, 23, 4, 11, 16, 23, 4, 11, 16, 23, 4, 11, 16, 23,
                               6, 10, 15, 21, 6, 10, 15, 21, 6, 10, 15, 21, 6, 10, 15, 21 };
static void md5_block(md5_t *m, const uint8_t *p) {
  uint32_t w[16], a = m->a, b = m->b, c = m->c, d = m->d;
  for (int i = 0; i < 16; i++) w[i] = p[4 * i] | (p[4 * i + 1] << 8) | (p[4 * i + 2] << 16) | ((uint32_t)p[4 * i + 3] << 24);
  for (int i = 0; i < 64; i++) {
    uint32_t f; int g;
    if (i < 16) { f = (b & c) | (~b & d); g = i; }
    else if (i < 32) { f = (d & b) | (~d & c); g = (5 * i + 1) & 15; }
    else if (i < 48) { f = b ^ c ^ d; g = (3 * i + 5) & 15; }
    else { f = c ^ (b | ~d); g = (7 * i) & 15; }
    const uint32_t t = a + f + K[i] + w[g];
    a = d; d = c; c = b; b += (t << S[i]) | (t >> (32 - S[i]));
  }
  m->a += a; m->b += b; m->c += c; m->d += d;
}
static void md5_init(md5_t *m) { m->a = 0x67452301; m->b = 0xefcdab89; m->c = 0x98badcfe; m->d = 0x10325476; m->len = 0; m->fill = 0; }
static void md5_add(md5_t *m, const uint8_t *p, size_t n) {
  m->len += n;
  while (n) {
    const size_t k = 64 - m->fill < n ? 64 - m->fill : n;
    memcpy(m->buf + m->fill, p, k);
    m->fill += k; p += k; n -= k;
    if (m->fill == 64) { md5_block(m, m->buf); m->fill = 0; }
  }
}
static void md5_hex(md5_t *m, char *out) {
  const uint64_t bits = m->len * 8;
  uint8_t pad[72] = { 0x80 };
  const size_t padlen = (m->fill < 56 ? 56 : 120) - m->fill;
  uint8_t lenb[8];
  for (int i = 0; i < 8; i++) lenb[i] = (uint8_t)(bits >> (8 * i));
  md5_add(m, pad, padlen);
  md5_add(m, lenb, 8);
  const uint32_t v[4] = { m->a, m->b, m->c, m->d };
  for (int i = 0; i < 16; i++) sprintf(out + 2 * i, "%02x", (v[i / 4] >> (8 * (i & 3))) & 0xff);
}

int main(int argc, char **argv) {
  if (argc < 2) { fprintf(stderr, "usage: %s clip.ivf\n", argv[0]); return 2; }
  FILE *f = fopen(argv[1], "rb");
  if (!f) { perror(argv[1]); return 2; }
  uint8_t hdr[32];
  if (fread(hdr, 1, 32, f) != 32 || memcmp(hdr, "DKIF", 4)) { fprintf(stderr, "not an IVF file\n"); return 2; }
  vp9b_decoder *d = vp9b_decoder_create(0);
  if (!d) { fprintf(stderr, "no CUDA device: this decoder has no CPU fallback\n"); return 3; }
  uint8_t fh[12], *chunk = NULL, *planes = NULL;
  size_t cap = 0, pcap = 0;
  int shown = 0;
  while (fread(fh, 1, 12, f) == 12) {
    const uint32_t sz = fh[0] | (fh[1] << 8) | (fh[2] << 16) | ((uint32_t)fh[3] << 24);
    if (sz > cap) { chunk = (uint8_t *)realloc(chunk, sz); cap = sz; }
    if (fread(chunk, 1, sz, f) != sz) break;
    if (vp9b_decoder_decode(d, chunk, sz) < 0) { fprintf(stderr, "decode error: %s\n", vp9b_decoder_last_error(d)); return 1; }
    vp9b_image im;
    if (vp9b_decoder_get_frame(d, &im) != 1) continue;
    const int bps = im.bit_depth > 8 ? 2 : 1;
    const int cw = (im.width + im.ss_x) >> im.ss_x, chh = (im.height + im.ss_y) >> im.ss_y;
    const size_t ysz = (size_t)im.width * im.height * bps, csz = (size_t)cw * chh * bps;
    if (ysz + 2 * csz > pcap) { planes = (uint8_t *)realloc(planes, ysz + 2 * csz); pcap = ysz + 2 * csz; }
    if (vp9b_decoder_read_frame(d, &im, planes, planes + ysz, planes + ysz + csz, im.width * bps, cw * bps) < 0) {
      fprintf(stderr, "read error: %s\n", vp9b_decoder_last_error(d));
      return 1;
    }
    md5_t m;
    char hex[33];
    md5_init(&m);
    md5_add(&m, planes, ysz + 2 * csz);   /* cropped Y, U, V rows back to back, as vpxdec hashes them */
    md5_hex(&m, hex);
    printf("%s  img-%ux%u-%04d.i420\n", hex, im.width, im.height, ++shown);
  }
  vp9b_decoder_destroy(d);
  free(chunk); free(planes); fclose(f);
  return 0;
}
