// TEST INFRASTRUCTURE -- CPU emulation of the column-worker loop filter (k_lf_sb).
// Compiles the kernel's own per-thread phase functions (cuda-vp9_b200/csrc/vp9_lf_sb.cuh) for the
// host and runs them "thread by thread" per worker, in schedules that are legal under the kernel's
// flag protocol but as adversarial as possible: all workers in lock step, one worker completely
// before the next, or a seeded random interleaving.  The result is compared with the oracle by
// tests/test_lf_emul.py.  Not part of the product.
#define VP9B_EMUL
#include "../../cuda-vp9_b200/csrc/vp9_lf_sb.cuh"

#include <cstdio>
#include <random>
#include <vector>

using namespace vp9b;

// threads of one phase (between two barriers) run in index order, or -- schedule bit 30 -- in reverse
// order: a phase whose threads touch disjoint samples must not care (this is what catches a wrong
// split of a luma row / column between its two threads)
static bool g_rev = false;
#define FOR_TID for (int t_ = 0, tid = 0; t_ < G::NT && ((tid = g_rev ? G::NT - 1 - t_ : t_), true); t_++)

template <typename Px, bool SUB>
struct Worker {
  typedef LfGeom<Px, SUB> G;
  typedef typename LfVecT<Px>::T Vec;
  std::vector<Px> buf[2];
  std::vector<Vec> pre;       // [NT][kPre]
  uint16_t dset[24];
  LfCtx<Px, SUB> C;
  Px *base[3];
  int r = 0, part = 0;        // next step: (row r, part 0 or 1)
  bool chain = false;
};

template <typename Px, bool SUB>
static int run(const vp9b_frame_hdr *h, const uint8_t *base, uint16_t *const planes[3], const int strides[3],
               const int rows[3], unsigned schedule, int *n_chain) {
  typedef LfGeom<Px, SUB> G;
  typedef typename LfVecT<Px>::T Vec;
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
  const int sb_cols = h->sb_cols, sb_rows = h->sb_rows;
  const vp9b_lfm *lfm = (const vp9b_lfm *)(base + h->sec[VP9B_SEC_LFM].offset);
  std::vector<Worker<Px, SUB>> W(sb_cols);
  std::vector<int> edge_done(sb_cols, 0);     // rows whose F_edge this worker has published
  *n_chain = 0;
  g_rev = (schedule >> 30) & 1;
  schedule &= ~(1u << 30);
  auto set_row = [&](Worker<Px, SUB> &w, int r) {
    LfCtx<Px, SUB> &C = w.C;
    C.tile = w.buf[r & 1].data();
    C.sbr = r;
    C.has_top = r > 0;
    const int t = r * sb_cols + C.sbc;
    C.M = &lfm[t];
    C.Mn = &lfm[C.has_next ? t + 1 : t];
    for (int p = 0; p < 3; p++) C.g[p] = w.base[p] + (size_t)r * (p ? G::nC : 64) * stride[p];
  };
  for (int c = 0; c < sb_cols; c++) {
    Worker<Px, SUB> &w = W[c];
    for (int k = 0; k < 2; k++) w.buf[k].assign(G::kTileSamples, (Px)0xAA);
    w.pre.resize((size_t)G::NT * G::kPre);
    LfCtx<Px, SUB> &C = w.C;
    C.dset = w.dset;
    C.lf_lim = h->lf_lim;
    C.lf_mblim = h->lf_mblim;
    C.sbc = c; C.sb_cols = sb_cols; C.mi_rows = h->mi_rows; C.bd = h->bit_depth;
    C.has_next = c + 1 < sb_cols;
    for (int p = 0; p < 3; p++) {
      C.stride[p] = stride[p];
      w.base[p] = buf[p].data() + c * (p ? G::nC : 64);
    }
    // prologue: the first superblock is fetched and committed before the loop
    set_row(w, 0);
    FOR_TID {
      Vec(&pre)[G::kPre] = *(Vec(*)[G::kPre]) & w.pre[(size_t)tid * G::kPre];
      lf_phase_prefetch<Px, SUB>(C.g, C.stride, C.has_next, tid, pre);
      lf_phase_commit<Px, SUB>(C, tid, pre);
    }
  }
  auto dependent = [&](Worker<Px, SUB> &w) {
    if (w.C.sbc == 0) return;
    FOR_TID { lf_phase_merge_left(w.C, tid); lf_phase_vertical(w.C, tid, true); }
  };
  // one step of a worker; returns false if its wait is not satisfied yet
  auto step = [&](int c) -> bool {
    Worker<Px, SUB> &w = W[c];
    if (w.r >= sb_rows) return false;
    LfCtx<Px, SUB> &C = w.C;
    const bool next_row = w.r + 1 < sb_rows;
    const int keep = next_row ? 8 : 0;
    if (w.part == 0) {
      set_row(w, w.r);
      bool ch = false;
      for (int u = 0; u < LfPlanUnits<Px, SUB>::value; u++) ch |= lf_plan_unit<Px, SUB>(*C.M, c > 0, u, C.dset);
      if (ch && c > 0 && edge_done[c - 1] <= w.r) return false;      // would wait for F_edge(r, c-1)
      w.chain = ch;
      *n_chain += ch;
      if (next_row) {      // fetch of the next superblock is issued first
        Px *g[3];
        for (int p = 0; p < 3; p++) g[p] = w.base[p] + (size_t)(w.r + 1) * (p ? G::nC : 64) * stride[p];
        FOR_TID
          lf_phase_prefetch<Px, SUB>(g, C.stride, C.has_next, tid, *(Vec(*)[G::kPre]) & w.pre[(size_t)tid * G::kPre]);
      }
      FOR_TID lf_phase_vertical(C, tid, false);
      if (w.chain) dependent(w);
      FOR_TID lf_phase_horizontal(C, tid, 0);
      if (C.has_next)
        FOR_TID lf_phase_next_left_edge(C, tid);
      FOR_TID lf_phase_store(C, tid, 0, keep);
      edge_done[c] = w.r + 1;
      w.part = 1;
      return true;
    }
    if (!w.chain && c > 0 && edge_done[c - 1] <= w.r) return false;
    if (!w.chain) dependent(w);
    FOR_TID lf_phase_horizontal(C, tid, 1);
    FOR_TID lf_phase_store(C, tid, 1, keep);
    if (next_row) {
      Px *nt = w.buf[(w.r + 1) & 1].data();
      FOR_TID lf_phase_carry(C, nt, tid);
      LfCtx<Px, SUB> N = C;
      N.tile = nt;
      FOR_TID lf_phase_commit<Px, SUB>(N, tid, *(Vec(*)[G::kPre]) & w.pre[(size_t)tid * G::kPre]);
    }
    w.part = 0;
    w.r++;
    return true;
  };
  // schedule 0: lock step (all workers part 0 of a row, then all part 1); 1: worker 0 completely,
  // then worker 1, ...; otherwise a seeded random choice among the workers that can advance
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
