// TEST INFRASTRUCTURE -- CPU emulation of the superblock data-flow loop filter (k_lf_sb).
// Compiles the kernel's own per-thread phase functions (cuda-vp9_b200/csrc/vp9_lf_sb.cuh) for the
// host and runs them "thread by thread" per CTA, in schedules that are legal under the kernel's
// flag protocol but as adversarial as possible (every CTA loads and pre-filters its tile before
// any neighbour has published anything, or a random mix of early and late CTAs).  The result is
// compared with the oracle by tests/test_lf_emul.py.  Not part of the product.
#define VP9B_EMUL
#include "../../cuda-vp9_b200/csrc/vp9_lf_sb.cuh"

#include <cstdio>
#include <random>
#include <vector>

using namespace vp9b;

template <typename Px, bool SUB>
static int run(const vp9b_frame_hdr *h, const uint8_t *base, uint16_t *const planes[3], const int strides[3],
               const int rows[3], unsigned schedule, int *n_chain) {
  typedef LfGeom<Px, SUB> G;
  // planes in the device geometry: aligned size + 64 samples of padding right / bottom
  const int aw = (int)h->mi_cols * 8, ah = (int)h->mi_rows * 8;
  std::vector<Px> buf[3];
  int stride[3], prow[3];
  for (int p = 0; p < 3; p++) {
    const int sx = p ? h->ss_x : 0, sy = p ? h->ss_y : 0;
    stride[p] = ((((aw >> sx) + 64) * (int)sizeof(Px) + 127) & ~127) / (int)sizeof(Px);
    prow[p] = (ah >> sy) + 64;
    buf[p].assign((size_t)stride[p] * prow[p] + 256, (Px)0x55);
    for (int r = 0; r < std::min(prow[p], rows[p]); r++)
      for (int x = 0; x < std::min(stride[p], strides[p]); x++) buf[p][(size_t)r * stride[p] + x] = (Px)planes[p][(size_t)r * strides[p] + x];
  }
  const int sb_cols = h->sb_cols, sb_rows = h->sb_rows, nsb = sb_cols * sb_rows;
  const vp9b_lfm *lfm = (const vp9b_lfm *)(base + h->sec[VP9B_SEC_LFM].offset);
  std::vector<std::vector<Px>> tiles(nsb, std::vector<Px>(G::kTileSamples, (Px)0xAA));
  std::vector<LfCtx<Px, SUB>> ctx(nsb);
  std::vector<char> chain(nsb, 0), early(nsb, 0);
  std::vector<uint16_t> dsets((size_t)nsb * 24, 0);
  std::mt19937 rng(schedule);
  auto phase1 = [&](int t) {
    LfCtx<Px, SUB> &C = ctx[t];
    const int sbr = t / sb_cols, sbc = t % sb_cols;
    C.tile = tiles[t].data();
    C.dset = &dsets[(size_t)t * 24];
    for (int p = 0; p < 3; p++) {
      const int n = p ? G::nC : 64;
      C.g[p] = buf[p].data() + (size_t)sbr * n * stride[p] + sbc * n;
      C.stride[p] = stride[p];
    }
    C.M = &lfm[t];
    C.Mn = &lfm[sbc + 1 < sb_cols ? t + 1 : t];
    C.lf_lim = h->lf_lim;
    C.lf_mblim = h->lf_mblim;
    C.sbr = sbr; C.sbc = sbc; C.sb_cols = sb_cols; C.mi_rows = h->mi_rows; C.bd = h->bit_depth;
    C.has_next = sbc + 1 < sb_cols;
    C.has_top = sbr > 0;
    bool ch = false;
    for (int tid = 0; tid < G::NT; tid++) ch |= lf_phase_dsets(C, tid);
    chain[t] = ch;
    for (int tid = 0; tid < G::NT; tid++) lf_phase_load_interior(C, tid);
    for (int tid = 0; tid < G::NT; tid++) lf_phase_vertical(C, tid, false);
  };
  auto dependent = [&](int t) {
    LfCtx<Px, SUB> &C = ctx[t];
    if (C.sbc == 0) return;
    for (int tid = 0; tid < G::NT; tid++) { lf_phase_merge_left(C, tid); lf_phase_vertical(C, tid, true); }
  };
  auto edge_part = [&](int t) {
    LfCtx<Px, SUB> &C = ctx[t];
    if (chain[t]) dependent(t);
    for (int tid = 0; tid < G::NT; tid++) lf_phase_load_top(C, tid, 0);
    for (int tid = 0; tid < G::NT; tid++) lf_phase_horizontal(C, tid, 0);
    if (C.has_next)
      for (int tid = 0; tid < G::NT; tid++) lf_phase_next_left_edge(C, tid);
    for (int tid = 0; tid < G::NT; tid++) lf_phase_store(C, tid, 0);
  };
  auto rest_part = [&](int t) {
    LfCtx<Px, SUB> &C = ctx[t];
    if (!chain[t]) dependent(t);
    for (int tid = 0; tid < G::NT; tid++) lf_phase_load_top(C, tid, 1);
    for (int tid = 0; tid < G::NT; tid++) lf_phase_horizontal(C, tid, 1);
    for (int tid = 0; tid < G::NT; tid++) lf_phase_store(C, tid, 1);
  };
  // schedule 0: every CTA runs phase 1 before anything is published; 1: strictly one CTA after the
  // other in ticket order; otherwise a seeded random mix of early and late CTAs
  for (int t = 0; t < nsb; t++) early[t] = schedule == 0 ? 1 : (schedule == 1 ? 0 : (rng() & 1));
  for (int t = 0; t < nsb; t++)
    if (early[t]) phase1(t);
  for (int r = 0; r < sb_rows; r++) {
    if (schedule == 1) {
      for (int c = 0; c < sb_cols; c++) { const int t = r * sb_cols + c; phase1(t); edge_part(t); rest_part(t); }
    } else {
      for (int c = 0; c < sb_cols; c++) { const int t = r * sb_cols + c; if (!early[t]) phase1(t); edge_part(t); }
      for (int c = 0; c < sb_cols; c++) rest_part(r * sb_cols + c);
    }
  }
  *n_chain = 0;
  for (int t = 0; t < nsb; t++) *n_chain += chain[t];
  for (int p = 0; p < 3; p++)
    for (int r = 0; r < std::min(prow[p], rows[p]); r++)
      for (int x = 0; x < std::min(stride[p], strides[p]); x++) planes[p][(size_t)r * strides[p] + x] = buf[p][(size_t)r * stride[p] + x];
  return 0;
}

// planes: uint16 samples in the oracle's geometry (aligned size + padding), filtered in place
extern "C" int lfemu_frame(const void *packed, uint16_t *y, uint16_t *u, uint16_t *v, const int *strides, const int *rows,
                           unsigned schedule, int *n_chain) {
  const vp9b_frame_hdr *h = (const vp9b_frame_hdr *)packed;
  const uint8_t *base = (const uint8_t *)packed;
  if (!h->lf_level) { *n_chain = 0; return 0; }
  uint16_t *const planes[3] = { y, u, v };
  const bool sub = h->ss_x == 1 && h->ss_y == 1;
  if (!sub && !(h->ss_x == 0 && h->ss_y == 0)) return -2;
  if (h->bit_depth == 8) return sub ? run<uint8_t, true>(h, base, planes, strides, rows, schedule, n_chain)
                                    : run<uint8_t, false>(h, base, planes, strides, rows, schedule, n_chain);
  return sub ? run<uint16_t, true>(h, base, planes, strides, rows, schedule, n_chain)
             : run<uint16_t, false>(h, base, planes, strides, rows, schedule, n_chain);
}
