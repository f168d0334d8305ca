// C ABI, host-only part: parser + superframe splitter (include/vp9_b200.h section 2).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <exception>
#include <new>
#include <string>

#include "vp9_b200.h"
#include "vp9_parser.h"

struct vp9b_parser {
  vp9b::Parser impl;
  std::string err;
};

extern "C" {

vp9b_parser *vp9b_parser_create(void) {
#if defined(__x86_64__) && defined(__AVX2__)
  // the host code is built for x86-64-v3 (see __graft_entry__.py): fail cleanly on an older CPU
  if (!__builtin_cpu_supports("avx2") || !__builtin_cpu_supports("bmi2")) {
    fprintf(stderr, "vp9b_parser_create: this build needs an x86-64-v3 CPU (AVX2/BMI2); rebuild with VP9B_HOST_ARCH=x86-64\n");
    return nullptr;
  }
#endif
  vp9b_parser *p = new (std::nothrow) vp9b_parser();
  if (p) {
    const char *t = getenv("VP9B_PARSE_THREADS");     // default for every parser of the process
    if (t && atoi(t) > 0) p->impl.set_threads(atoi(t));
  }
  return p;
}

void vp9b_parser_destroy(vp9b_parser *p) { delete p; }

const char *vp9b_parser_last_error(vp9b_parser *p) {
  if (!p) return "null parser";
  return p->err.empty() ? p->impl.error().c_str() : p->err.c_str();
}

int vp9b_parser_parse(vp9b_parser *p, const uint8_t *data, size_t size, const void **packed,
                      size_t *bytes, int *show_slot) {
  if (!p || !packed || !bytes || !show_slot) return VP9B_ERR_INVALID;
  *packed = nullptr;
  *bytes = 0;
  *show_slot = -1;
  p->err.clear();
  int e;
  try {   // nothing may propagate through the C ABI (a hostile header can still ask for a lot of memory)
    e = p->impl.parse(data, size);
  } catch (const std::bad_alloc &) {
    p->err = "out of memory";
    return VP9B_ERR_MEM;
  } catch (const std::exception &ex) {
    p->err = ex.what();
    return VP9B_ERR_INVALID;
  }
  if (e) return e;
  *packed = p->impl.packed();
  *bytes = p->impl.packed_bytes();
  *show_slot = p->impl.show_slot();
  return VP9B_OK;
}

int vp9b_parser_set_threads(vp9b_parser *p, int threads) {
  if (!p || threads < 1) return VP9B_ERR_INVALID;
  p->impl.set_threads(threads);
  return VP9B_OK;
}

int vp9b_parser_set_output_allocator(vp9b_parser *p, void *(*alloc)(void *priv, size_t bytes), void *priv) {
  if (!p) return VP9B_ERR_INVALID;
  p->impl.set_output_allocator(alloc, priv);
  return VP9B_OK;
}

int vp9b_parser_slot_info(vp9b_parser *p, int slot, uint32_t *width, uint32_t *height,
                          int *bit_depth, int *ss_x, int *ss_y) {
  if (!p || slot < 0 || slot >= VP9B_NUM_SLOTS) return VP9B_ERR_INVALID;
  const vp9b::SlotInfo &s = p->impl.slot(slot);
  if (!s.valid) return VP9B_ERR_INVALID;
  if (width) *width = s.width;
  if (height) *height = s.height;
  if (bit_depth) *bit_depth = s.bit_depth;
  if (ss_x) *ss_x = s.ss_x;
  if (ss_y) *ss_y = s.ss_y;
  return VP9B_OK;
}

// vp9_parse_superframe_index (vp9_dx_iface.c / vp9_decoder.c): marker byte 0b110xxyyy at the
// end, xx+1 = bytes per size, yyy+1 = frame count; the index is bracketed by the marker.
int vp9b_parse_superframe(const uint8_t *data, size_t size, uint32_t sizes[8]) {
  if (!data || !size || !sizes) return VP9B_ERR_INVALID;
  const uint8_t marker = data[size - 1];
  if ((marker & 0xe0) == 0xc0) {
    const int frames = (marker & 0x7) + 1;
    const int mag = ((marker >> 3) & 0x3) + 1;
    const size_t index_sz = 2 + (size_t)mag * frames;
    if (size >= index_sz && data[size - index_sz] == marker) {
      const uint8_t *x = data + size - index_sz + 1;
      size_t total = 0;
      for (int i = 0; i < frames; ++i) {
        uint32_t s = 0;
        for (int j = 0; j < mag; ++j) s |= (uint32_t)(*x++) << (j * 8);
        sizes[i] = s;
        total += s;
      }
      if (total + index_sz > size) return VP9B_ERR_CORRUPT;
      return frames;
    }
  }
  sizes[0] = (uint32_t)size;
  return 1;
}

}  // extern "C"
