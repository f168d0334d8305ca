/* Host-parser micro benchmark: parses every frame of an IVF clip N times through the C ABI
 * (vp9b_parser_*) and prints ms per frame and an FNV hash of the packed buffers (so that an
 * optimisation of the parser can be checked for identical output).
 *   gcc -O2 -Iinclude tools/parse_bench.c -Lcuda-vp9_b200 -lvp9b200_host -o /tmp/parse_bench */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "vp9_b200.h"

static double now(void) {
  struct timespec t;
  clock_gettime(CLOCK_MONOTONIC, &t);
  return t.tv_sec + 1e-9 * t.tv_nsec;
}

int main(int argc, char **argv) {
  if (argc < 2) { fprintf(stderr, "usage: parse_bench clip.ivf [passes] [threads]\n"); return 2; }
  const int passes = argc > 2 ? atoi(argv[2]) : 3, threads = argc > 3 ? atoi(argv[3]) : 1;
  FILE *f = fopen(argv[1], "rb");
  if (!f) { perror(argv[1]); return 1; }
  fseek(f, 0, SEEK_END);
  long n = ftell(f);
  fseek(f, 0, SEEK_SET);
  uint8_t *buf = malloc(n);
  if (fread(buf, 1, n, f) != (size_t)n) return 1;
  fclose(f);
  double best = 1e30;
  int frames = 0;
  uint64_t hash = 0;
  for (int pass = 0; pass < passes; pass++) {
    vp9b_parser *p = vp9b_parser_create();
    if (!p) { fprintf(stderr, "vp9b_parser_create failed\n"); return 1; }
    if (threads > 1) vp9b_parser_set_threads(p, threads);
    long pos = buf[6] | (buf[7] << 8);
    frames = 0;
    if (pass == 0) hash = 1469598103934665603ull;
    double t0 = now(), hash_s = 0;
    while (pos + 12 <= n) {
      uint32_t sz;
      memcpy(&sz, buf + pos, 4);
      const uint8_t *chunk = buf + pos + 12;
      pos += 12 + sz;
      uint32_t sizes[8];
      int k = vp9b_parse_superframe(chunk, sz, sizes);
      if (k < 0) { fprintf(stderr, "bad superframe\n"); return 1; }
      size_t off = 0;
      for (int i = 0; i < k; i++) {
        const void *packed;
        size_t bytes;
        int show;
        int rc = vp9b_parser_parse(p, chunk + off, sizes[i], &packed, &bytes, &show);
        if (rc) { fprintf(stderr, "parse error %d: %s\n", rc, vp9b_parser_last_error(p)); return 1; }
        off += sizes[i];
        frames++;
        if (pass == 0 && packed) {
          double h0 = now();
          const uint8_t *q = packed;
          for (size_t j = 0; j < bytes; j++) hash = (hash ^ q[j]) * 1099511628211ull;
          hash_s += now() - h0;
        }
      }
    }
    double dt = now() - t0 - hash_s;
    if (dt < best) best = dt;
    vp9b_parser_destroy(p);
  }
  printf("%s: %d frames, %.3f ms/frame (best of %d, %d thread%s), packed hash %016llx\n", argv[1], frames,
         best / frames * 1e3, passes, threads, threads > 1 ? "s" : "", (unsigned long long)hash);
  return 0;
}
