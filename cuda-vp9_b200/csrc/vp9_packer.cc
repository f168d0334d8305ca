// vp9b_packer_*: builds the packed per-frame buffer (vp9_b200.h) from what a host decoder has
// parsed -- the role of the reference's createBuffers / createBuffersTr
// (vpx-master/inter_cuda_kernel.cu:897-1028, intra_cuda_kernel.cu:673,1252-1304), but fed record by
// record through a C ABI instead of walking libvpx structures on the CUDA side.  A libvpx host calls
// it from decode_block() (integration/libvpx/vp9_b200_shim.inc shows the binding); the sections it
// produces are the ones vp9b_parser_* emits.
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <new>
#include <string>
#include <vector>

#include "vp9_b200.h"

namespace {
const uint8_t kBw8[13] = { 1, 1, 1, 1, 1, 2, 2, 2, 4, 4, 4, 8, 8 };
const uint8_t kBh8[13] = { 1, 1, 1, 1, 2, 1, 2, 4, 2, 4, 8, 4, 8 };
}  // namespace

struct vp9b_packer {
  std::string err;
  vp9b_frame_hdr hdr;
  bool open = false;
  std::vector<vp9b_block> blocks;
  std::vector<vp9b_submv> submv;
  std::vector<vp9b_txb> txb_inter[4];
  std::vector<vp9b_txb> txb_intra[3];
  std::vector<uint32_t> intra_sb[3];     // superblock index of every intra record (call order)
  std::vector<int32_t> coef;             // widened; narrowed at finish for 8-bit streams
  std::vector<vp9b_lfm> lfm;
  std::vector<vp9b_lfm_uv> lfm_uv;       // 4:2:2 / 4:4:0 frames only
  std::vector<uint8_t> out;
  int fail(const char *m) { err = m; return VP9B_ERR_INVALID; }
};

extern "C" vp9b_packer *vp9b_packer_create(void) { return new (std::nothrow) vp9b_packer(); }
extern "C" void vp9b_packer_destroy(vp9b_packer *p) { delete p; }
extern "C" const char *vp9b_packer_last_error(vp9b_packer *p) { return p ? p->err.c_str() : "null packer"; }

extern "C" int vp9b_packer_begin(vp9b_packer *p, const vp9b_frame_hdr *in) {
  if (!p || !in) return VP9B_ERR_INVALID;
  try {
    if (!in->width || !in->height || in->width > 65536 || in->height > 65536) return p->fail("bad frame size");
    if (in->bit_depth != 8 && in->bit_depth != 10 && in->bit_depth != 12) return p->fail("bad bit depth");
    if (in->ss_x > 1 || in->ss_y > 1) return p->fail("bad chroma subsampling");
    if (in->dst_slot < 0 || in->dst_slot >= VP9B_NUM_SLOTS) return p->fail("bad dst slot");
    vp9b_frame_hdr &h = p->hdr;
    h = *in;
    h.magic = VP9B_MAGIC;
    h.mi_cols = (h.width + 7) / 8; h.mi_rows = (h.height + 7) / 8;
    h.sb_cols = (h.mi_cols + 7) / 8; h.sb_rows = (h.mi_rows + 7) / 8;
    h.coef_bytes = h.bit_depth > 8 ? 4 : 2;
    memset(h.sec, 0, sizeof(h.sec));
    memset(h.n_inter_px, 0, sizeof(h.n_inter_px));
    memset(h.n_compound_px, 0, sizeof(h.n_compound_px));
    memset(h.n_txb_inter_size, 0, sizeof(h.n_txb_inter_size));
    p->blocks.clear(); p->submv.clear(); p->coef.clear();
    for (auto &v : p->txb_inter) v.clear();
    for (int i = 0; i < 3; i++) { p->txb_intra[i].clear(); p->intra_sb[i].clear(); }
    p->lfm.assign(h.lf_level ? (size_t)h.sb_cols * h.sb_rows : 0, vp9b_lfm());
    if (!p->lfm.empty()) memset(p->lfm.data(), 0, p->lfm.size() * sizeof(vp9b_lfm));
    p->lfm_uv.assign((h.lf_level && h.ss_x != h.ss_y) ? (size_t)h.sb_cols * h.sb_rows : 0, vp9b_lfm_uv());
    if (!p->lfm_uv.empty()) memset(p->lfm_uv.data(), 0, p->lfm_uv.size() * sizeof(vp9b_lfm_uv));
    p->open = true;
    return VP9B_OK;
  } catch (const std::bad_alloc &) {
    p->err = "out of memory";
    return VP9B_ERR_MEM;
  }
}

extern "C" int vp9b_packer_add_block(vp9b_packer *p, const vp9b_block *b, const int16_t sub_mv[4][2][2]) {
  if (!p || !b || !p->open) return VP9B_ERR_INVALID;
  try {
    const vp9b_frame_hdr &h = p->hdr;
    if (b->bsize > 12 || b->mi_col >= h.mi_cols || b->mi_row >= h.mi_rows || b->ref0 > 3 || b->ref1 > 3 || b->filter > 3)
      return p->fail("block record out of range");
    if (b->ref0 && h.intra_only) return p->fail("inter block in an intra-only frame");
    vp9b_block o = *b;
    if (o.ref0 && o.bsize < 3) {
      if (!sub_mv) return p->fail("sub-8x8 inter block without its four motion vectors");
      vp9b_submv s;
      memset(&s, 0, sizeof(s));
      memcpy(s.luma, sub_mv, sizeof(s.luma));
      for (int rf = 0; rf < 2; rf++) {   // 4:2:0 chroma MV: average of the four (vp9_reconinter.c:61-87)
        const int sr = sub_mv[0][rf][0] + sub_mv[1][rf][0] + sub_mv[2][rf][0] + sub_mv[3][rf][0];
        const int sc = sub_mv[0][rf][1] + sub_mv[1][rf][1] + sub_mv[2][rf][1] + sub_mv[3][rf][1];
        s.chroma[rf][0] = (int16_t)((sr < 0 ? sr - 2 : sr + 2) / 4);
        s.chroma[rf][1] = (int16_t)((sc < 0 ? sc - 2 : sc + 2) / 4);
      }
      const uint32_t idx = (uint32_t)p->submv.size();
      p->submv.push_back(s);
      memset(o.mv, 0, sizeof(o.mv));
      memcpy(&o.mv[0][0], &idx, 4);
    }
    if (o.ref0) {
      const int x_mis = std::min<int>(kBw8[o.bsize], h.mi_cols - o.mi_col), y_mis = std::min<int>(kBh8[o.bsize], h.mi_rows - o.mi_row);
      for (int pl = 0; pl < 3; pl++) {
        const uint32_t px = (uint32_t)((x_mis * 8) >> (pl ? h.ss_x : 0)) * ((y_mis * 8) >> (pl ? h.ss_y : 0));
        p->hdr.n_inter_px[pl] += px;
        if (o.ref1) p->hdr.n_compound_px[pl] += px;
      }
    }
    p->blocks.push_back(o);
    return VP9B_OK;
  } catch (const std::bad_alloc &) {
    p->err = "out of memory";
    return VP9B_ERR_MEM;
  }
}

extern "C" int vp9b_packer_add_txb(vp9b_packer *p, int plane, int inter, int x4, int y4, int tx_size, int tx_type,
                                   int mode, int have_top, int have_left, int have_right, const void *dqcoeff,
                                   int coef_bytes, const int16_t *scan, int eob) {
  if (!p || !p->open) return VP9B_ERR_INVALID;
  try {
    const vp9b_frame_hdr &h = p->hdr;
    const int n = 16 << (2 * tx_size);
    if (plane < 0 || plane > 2 || tx_size < 0 || tx_size > 3 || tx_type < 0 || tx_type > 3 || eob < 0 || eob > n)
      return p->fail("transform block record out of range");
    const int pw4 = (int)(h.mi_cols * 2) >> (plane ? h.ss_x : 0), ph4 = (int)(h.mi_rows * 2) >> (plane ? h.ss_y : 0);
    // a transform block may hang over the aligned frame edge (into the padding), never start beyond it
    if (x4 < 0 || y4 < 0 || x4 >= std::max(pw4, 1) + 15 || y4 >= std::max(ph4, 1) + 15) return p->fail("transform block outside the frame");
    if (eob && (!dqcoeff || !scan || (coef_bytes != 2 && coef_bytes != 4))) return p->fail("transform block without coefficients");
    if (inter && !eob) return VP9B_OK;
    if (!inter && (mode < 0 || mode > 9)) return p->fail("bad intra mode");
    vp9b_txb t;
    t.coef_off = (uint32_t)p->coef.size();
    t.x4 = (uint16_t)x4; t.y4 = (uint16_t)y4;
    t.eob = (uint16_t)eob;
    t.tx = (uint8_t)(tx_size | (tx_type << 2) | (plane << 4));
    t.aux = inter ? 0 : (uint8_t)(mode | (have_top ? 16 : 0) | (have_left ? 32 : 0) | (have_right ? 64 : 0));
    for (int i = 0; i < eob; i++) {
      const int pos = scan[i];
      if (pos < 0 || pos >= n) return p->fail("scan position out of range");
      p->coef.push_back(coef_bytes == 2 ? (int32_t)((const int16_t *)dqcoeff)[pos] : ((const int32_t *)dqcoeff)[pos]);
    }
    if (inter) {
      p->txb_inter[tx_size].push_back(t);
    } else {
      const int sb_sh = plane && h.ss_x ? 3 : 4, sb_shy = plane && h.ss_y ? 3 : 4;   // 4-sample units per superblock: 16 or 8
      const uint32_t sbc = std::min<uint32_t>((uint32_t)x4 >> sb_sh, h.sb_cols - 1), sbr = std::min<uint32_t>((uint32_t)y4 >> sb_shy, h.sb_rows - 1);
      p->txb_intra[plane].push_back(t);
      p->intra_sb[plane].push_back(sbr * h.sb_cols + sbc);
    }
    return VP9B_OK;
  } catch (const std::bad_alloc &) {
    p->err = "out of memory";
    return VP9B_ERR_MEM;
  }
}

extern "C" int vp9b_packer_set_lfm(vp9b_packer *p, int sb_row, int sb_col, const vp9b_lfm *m) {
  if (!p || !m || !p->open) return VP9B_ERR_INVALID;
  const vp9b_frame_hdr &h = p->hdr;
  if (!h.lf_level) return VP9B_OK;
  if (sb_row < 0 || sb_col < 0 || sb_row >= (int)h.sb_rows || sb_col >= (int)h.sb_cols) return p->fail("loop-filter mask outside the frame");
  p->lfm[(size_t)sb_row * h.sb_cols + sb_col] = *m;
  return VP9B_OK;
}

extern "C" int vp9b_packer_set_lfm_uv(vp9b_packer *p, int sb_row, int sb_col, const vp9b_lfm_uv *m) {
  if (!p || !m || !p->open) return VP9B_ERR_INVALID;
  const vp9b_frame_hdr &h = p->hdr;
  if (p->lfm_uv.empty()) return VP9B_OK;      // only 4:2:2 / 4:4:0 frames with the filter on carry the section
  if (sb_row < 0 || sb_col < 0 || sb_row >= (int)h.sb_rows || sb_col >= (int)h.sb_cols) return p->fail("loop-filter mask outside the frame");
  p->lfm_uv[(size_t)sb_row * h.sb_cols + sb_col] = *m;
  return VP9B_OK;
}

extern "C" int vp9b_packer_finish(vp9b_packer *p, const void **packed, size_t *bytes) {
  if (!p || !packed || !bytes || !p->open) return VP9B_ERR_INVALID;
  try {
    vp9b_frame_hdr &h = p->hdr;
    const uint32_t nsb = h.sb_cols * h.sb_rows;
    // intra records: grouped by superblock (stable, so decode order inside a superblock is kept --
    // tiles may have been parsed in any order), with one range per (plane, superblock)
    std::vector<vp9b_txb> intra[3];
    std::vector<vp9b_sbrange> ranges((size_t)3 * nsb, vp9b_sbrange{ 0, 0 });
    for (int pl = 0; pl < 3; pl++) {
      const size_t n = p->txb_intra[pl].size();
      std::vector<uint32_t> cnt(nsb + 1, 0);
      for (size_t i = 0; i < n; i++) cnt[p->intra_sb[pl][i] + 1]++;
      for (uint32_t s = 0; s < nsb; s++) cnt[s + 1] += cnt[s];
      for (uint32_t s = 0; s < nsb; s++) ranges[(size_t)pl * nsb + s] = vp9b_sbrange{ cnt[s], cnt[s + 1] - cnt[s] };
      intra[pl].resize(n);
      std::vector<uint32_t> pos(cnt.begin(), cnt.end() - 1);
      for (size_t i = 0; i < n; i++) intra[pl][pos[p->intra_sb[pl][i]]++] = p->txb_intra[pl][i];
    }
    std::vector<vp9b_txb> inter_all;
    for (int s = 0; s < 4; s++) {
      h.n_txb_inter_size[s] = (uint32_t)p->txb_inter[s].size();
      inter_all.insert(inter_all.end(), p->txb_inter[s].begin(), p->txb_inter[s].end());
    }
    std::vector<int16_t> coef16;
    if (h.coef_bytes == 2) {
      coef16.resize(p->coef.size());
      for (size_t i = 0; i < p->coef.size(); i++) coef16[i] = (int16_t)p->coef[i];
    }
    struct Src { const void *ptr; size_t elem, count; } src[VP9B_SEC_COUNT];
    memset(src, 0, sizeof(src));
    src[VP9B_SEC_BLOCKS] = { p->blocks.data(), sizeof(vp9b_block), p->blocks.size() };
    src[VP9B_SEC_SUBMV] = { p->submv.data(), sizeof(vp9b_submv), p->submv.size() };
    src[VP9B_SEC_TXB_INTER] = { inter_all.data(), sizeof(vp9b_txb), inter_all.size() };
    for (int pl = 0; pl < 3; pl++) src[VP9B_SEC_TXB_INTRA_Y + pl] = { intra[pl].data(), sizeof(vp9b_txb), intra[pl].size() };
    src[VP9B_SEC_SB_INTRA] = { ranges.data(), sizeof(vp9b_sbrange), ranges.size() };
    if (h.coef_bytes == 2) src[VP9B_SEC_COEF] = { coef16.data(), 2, coef16.size() };
    else src[VP9B_SEC_COEF] = { p->coef.data(), 4, p->coef.size() };
    src[VP9B_SEC_LFM] = { p->lfm.data(), sizeof(vp9b_lfm), p->lfm.size() };
    src[VP9B_SEC_LFM_UV] = { p->lfm_uv.data(), sizeof(vp9b_lfm_uv), p->lfm_uv.size() };
    uint64_t off = (sizeof(vp9b_frame_hdr) + 15) & ~(size_t)15;
    for (int i = 0; i < VP9B_SEC_COUNT; i++) {
      h.sec[i].offset = (uint32_t)off;
      h.sec[i].count = (uint32_t)src[i].count;
      off += (src[i].elem * src[i].count + 15) & ~(size_t)15;
    }
    if (off > 0xffffffffull) return p->fail("packed frame too large");
    h.total_bytes = (uint32_t)off;
    p->out.assign((size_t)off, 0);
    memcpy(p->out.data(), &h, sizeof(h));
    for (int i = 0; i < VP9B_SEC_COUNT; i++)
      if (src[i].count) memcpy(p->out.data() + h.sec[i].offset, src[i].ptr, src[i].elem * src[i].count);
    p->open = false;
    *packed = p->out.data();
    *bytes = p->out.size();
    return VP9B_OK;
  } catch (const std::bad_alloc &) {
    p->err = "out of memory";
    return VP9B_ERR_MEM;
  }
}
