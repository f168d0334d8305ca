// TEST INFRASTRUCTURE -- CPU emulation of the warp-per-column loop filter (k_lf_warp, 8-bit 4:2:0).
// Compiles the kernel's per-lane functions (cuda-vp9_b200/csrc/vp9_lf_warp.cuh) for the host:
//   * lfw_edge_test: the packed filter arithmetic on four sample positions, compared by
//     tests/test_lf_warp.py with the scalar reference filter;
//   * lfwemu_frame: the whole decomposition, worker by worker, in schedules that are legal under
//     the kernel's flag protocol (lock step, one worker completely before the next, seeded random).
// Not part of the product.
#define VP9B_EMUL
#include "../../cuda-vp9_b200/csrc/vp9_lf_warp.cuh"

#include <cstdio>
#include <random>
#include <vector>

using namespace vp9b;

// taps[16][4]: taps p7..q7 of four sample positions; filtered in place.  Returns 1 if changed.
extern "C" int lfw_edge_test(int size, uint8_t *taps, int blimit, int limit, int thresh) {
  uint32_t v[16];
  for (int i = 0; i < 16; i++) memcpy(&v[i], taps + 4 * i, 4);
  bool ch;
  if (size == 16) ch = lfw_edge<16>(v, blimit, limit, thresh);
  else if (size == 8) ch = lfw_edge<8>(v, blimit, limit, thresh);
  else ch = lfw_edge<4>(v, blimit, limit, thresh);
  for (int i = 0; i < 16; i++) memcpy(taps + 4 * i, &v[i], 4);
  return ch;
}

struct Worker {
  std::vector<uint8_t> buf[2];
  uint16_t dset[24];
  LfwCtx C;
  uint8_t *base[3];
  int r = 0, part = 0;
  bool chain = false, early = false;
};

extern "C" int lfwemu_frame(const void *packed, uint16_t *y, uint16_t *u, uint16_t *v, const int *strides, const int *rows,
                            unsigned schedule, int *n_chain) {
  const vp9b_frame_hdr *h = (const vp9b_frame_hdr *)packed;
  const uint8_t *pbase = (const uint8_t *)packed;
  *n_chain = 0;
  if (!h->lf_level) return 0;
  if (h->bit_depth != 8 || h->ss_x != 1 || h->ss_y != 1) return -2;
  uint16_t *const planes[3] = { y, u, v };
  const int aw = (int)h->mi_cols * 8, ah = (int)h->mi_rows * 8;
  std::vector<uint8_t> buf[3];
  int stride[3], prow[3];
  for (int p = 0; p < 3; p++) {
    const int s = p ? 1 : 0;
    stride[p] = (((aw >> s) + 64) + 127) & ~127;
    prow[p] = (ah >> s) + 64;
    buf[p].assign((size_t)stride[p] * prow[p] + 256, (uint8_t)0x55);
    for (int r = 0; r < std::min(prow[p], rows[p]); r++)
      for (int x = 0; x < std::min(stride[p], strides[p]); x++) buf[p][(size_t)r * stride[p] + x] = (uint8_t)planes[p][(size_t)r * strides[p] + x];
  }
  const int sb_cols = h->sb_cols, sb_rows = h->sb_rows;
  const vp9b_lfm *lfm = (const vp9b_lfm *)(pbase + h->sec[VP9B_SEC_LFM].offset);
  std::vector<Worker> W(sb_cols);
  std::vector<int> edge_done(sb_cols, 0);
  auto row_ptrs = [&](Worker &w, int r, uint8_t *g[3]) {
    for (int p = 0; p < 3; p++) g[p] = w.base[p] + (size_t)r * (p ? 32 : 64) * stride[p];
  };
  auto set_row = [&](Worker &w, int r) {
    LfwCtx &C = w.C;
    C.tile = w.buf[0].data();
    C.sbr = r;
    C.has_top = r > 0;
    const int t = r * sb_cols + C.sbc;
    C.M = &lfm[t];
    C.Mn = &lfm[C.has_next ? t + 1 : t];
    row_ptrs(w, r, C.g);
  };
  for (int c = 0; c < sb_cols; c++) {
    Worker &w = W[c];
    for (int k = 0; k < 2; k++) w.buf[k].assign(LfwGeom::kTileBytes, (uint8_t)0xAA);
    LfwCtx &C = w.C;
    C.dset = w.dset;
    C.lf_lim = h->lf_lim;
    C.lf_mblim = h->lf_mblim;
    C.sbc = c; C.sb_cols = sb_cols; C.mi_rows = h->mi_rows;
    C.has_next = c + 1 < sb_cols;
    for (int p = 0; p < 3; p++) {
      C.stride[p] = stride[p];
      w.base[p] = buf[p].data() + c * (p ? 32 : 64);
    }
    set_row(w, 0);
    for (int lane = 0; lane < 32; lane++) lfw_phase_fetch(C.tile, C.g, C.stride, C.has_next, lane);
  }
  auto dependent = [&](Worker &w) {
    if (w.C.sbc == 0) return;
    for (int lane = 0; lane < 32; lane++) lfw_phase_merge_left(w.C, lane);
    for (int lane = 0; lane < 32; lane++) lfw_phase_vertical(w.C, lane, true);
  };
  auto step = [&](int c) -> bool {
    Worker &w = W[c];
    if (w.r >= sb_rows) return false;
    LfwCtx &C = w.C;
    const bool next_row = w.r + 1 < sb_rows;
    const int keep = next_row ? 8 : 0;
    if (w.part == 0) {
      set_row(w, w.r);
      bool ch = false;
      for (int u = 0; u < kLfwPlanUnits; u++) ch |= lfw_plan_unit(*C.M, c > 0, u, C.dset);
      if (ch && c > 0 && edge_done[c - 1] <= w.r) return false;
      w.chain = ch;
      *n_chain += ch;
      for (int lane = 0; lane < 32; lane++) lfw_phase_vertical(C, lane, false);
      if (w.chain) dependent(w);
      for (int lane = 0; lane < 32; lane++) lfw_phase_horizontal(C, lane, 0);
      if (C.has_next)
        for (int lane = 0; lane < 32; lane++) lfw_phase_next_left_edge(C, lane);
      for (int lane = 0; lane < 32; lane++) lfw_phase_store(C, lane, 0, keep);
      edge_done[c] = w.r + 1;
      // as in the kernel: the block columns the left edge cannot reach are filtered BEFORE the wait
      // for the left neighbour (i.e. possibly before that neighbour has even started this row)
      w.early = !w.chain && c > 0 && !lfw_plan_deep(C.dset);
      if (w.early)
        for (int lane = 0; lane < 32; lane++) lfw_phase_horizontal(C, lane, 2);
      w.part = 1;
      return true;
    }
    if (!w.chain && c > 0 && edge_done[c - 1] <= w.r) return false;
    if (!w.chain) dependent(w);
    for (int lane = 0; lane < 32; lane++) lfw_phase_horizontal(C, lane, w.early ? 3 : 1);
    for (int lane = 0; lane < 32; lane++) lfw_phase_store(C, lane, 1, keep);
    if (next_row) {      // single tile buffer: bottom rows to the top, then the next superblock over the interior
      for (int lane = 0; lane < 32; lane++) lfw_phase_carry(C, C.tile, lane);
      uint8_t *g[3];
      row_ptrs(w, w.r + 1, g);
      for (int lane = 0; lane < 32; lane++) lfw_phase_fetch(C.tile, g, C.stride, C.has_next, lane);
    }
    w.part = 0;
    w.r++;
    return true;
  };
  if (schedule == 0) {
    for (int r = 0; r < sb_rows; r++)
      for (int part = 0; part < 2; part++)
        for (int c = 0; c < sb_cols; c++)
          if (!step(c)) return -10;
  } else if (schedule == 1) {
    for (int c = 0; c < sb_cols; c++)
      for (int k = 0; k < 2 * sb_rows; k++)
        if (!step(c)) return -11;
  } else {
    std::mt19937 rng(schedule);
    int left = 2 * sb_rows * sb_cols, stuck = 0;
    while (left > 0) {
      const int c = (int)(rng() % (unsigned)sb_cols);
      if (step(c)) { left--; stuck = 0; }
      else if (++stuck > 100000) return -12;
    }
  }
  for (int p = 0; p < 3; p++)
    for (int r = 0; r < std::min(prow[p], rows[p]); r++)
      for (int x = 0; x < std::min(stride[p], strides[p]); x++) planes[p][(size_t)r * strides[p] + x] = buf[p][(size_t)r * stride[p] + x];
  return 0;
}
