// Host-side VP9 bitstream parser -> packed per-frame buffer.  See vp9_parser.h.
// Every function cites the reference routine whose behaviour it reproduces (paths relative to
// /root/reference/libvpx).
#include "vp9_parser.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "vp9_tables.h"

namespace vp9b {
// Largest frame dimension accepted from a header (role of libvpx's DECODE_WIDTH_LIMIT /
// DECODE_HEIGHT_LIMIT, vp9_decodeframe.c:1705-1711): the header is untrusted and every per-frame
// table is sized from it before any tile data has been validated.  16384 covers 16K video; the
// GPU side has the same limit (vp9gpu_frame_submit).
constexpr int kMaxFrameDim = 16384;
}  // namespace vp9b
namespace vp9b {

namespace {

enum { BLOCK_4X4 = 0, BLOCK_4X8, BLOCK_8X4, BLOCK_8X8, BLOCK_8X16, BLOCK_16X8, BLOCK_16X16,
       BLOCK_16X32, BLOCK_32X16, BLOCK_32X32, BLOCK_32X64, BLOCK_64X32, BLOCK_64X64,
       BLOCK_INVALID = 13 };
enum { DC_PRED = 0, V_PRED, H_PRED, D45_PRED, D135_PRED, D117_PRED, D153_PRED, D207_PRED,
       D63_PRED, TM_PRED, NEARESTMV = 10, NEARMV, ZEROMV, NEWMV };
enum { TX_4X4 = 0, TX_8X8, TX_16X16, TX_32X32 };
enum { ONLY_4X4 = 0, ALLOW_8X8, ALLOW_16X16, ALLOW_32X32, TX_MODE_SELECT };
enum { INTRA_FRAME = 0, LAST_FRAME, GOLDEN_FRAME, ALTREF_FRAME, REF_NONE = -1 };
enum { SINGLE_REFERENCE = 0, COMPOUND_REFERENCE, REFERENCE_MODE_SELECT };
enum { EIGHTTAP = 0, EIGHTTAP_SMOOTH, EIGHTTAP_SHARP, BILINEAR, SWITCHABLE = 4,
       SWITCHABLE_FILTERS = 3 };
enum { PARTITION_NONE = 0, PARTITION_HORZ, PARTITION_VERT, PARTITION_SPLIT };
enum { SEG_LVL_ALT_Q = 0, SEG_LVL_ALT_LF, SEG_LVL_REF_FRAME, SEG_LVL_SKIP };
enum { KEY_FRAME = 0, INTER_FRAME = 1 };
enum { MV_JOINT_ZERO = 0, MV_JOINT_HNZVZ, MV_JOINT_HZVNZ, MV_JOINT_HNZVNZ };

// Coding trees (vp9_entropymode.c:245-266,337; vp9_entropymv.c:14-28; vp9_seg_common.c:58).
// value <= 0: leaf (-value); value > 0: index of next node pair.
const int8_t kIntraModeTree[18] = { -DC_PRED, 2, -TM_PRED, 4, -V_PRED, 6, 8, 12, -H_PRED, 10,
                                    -D135_PRED, -D117_PRED, -D45_PRED, 14, -D63_PRED, 16,
                                    -D153_PRED, -D207_PRED };
const int8_t kInterModeTree[6] = { -2, 2, 0, 4, -1, -3 };  // leaves: NEAREST 0 NEAR 1 ZERO 2 NEW 3
const int8_t kPartitionTree[6] = { 0, 2, -1, 4, -2, -3 };
const int8_t kSwitchableInterpTree[4] = { 0, 2, -1, -2 };
const int8_t kSegmentTree[14] = { 2, 4, 6, 8, 10, 12, 0, -1, -2, -3, -4, -5, -6, -7 };
const int8_t kMvJointTree[6] = { 0, 2, -1, 4, -2, -3 };
const int8_t kMvClassTree[20] = { 0, 2, -1, 4, 6, 8, -2, -3, 10, 12, -4, -5, -6, 14, 16, 18,
                                  -7, -8, -9, -10 };
const int8_t kMvClass0Tree[2] = { 0, -1 };
const int8_t kMvFpTree[6] = { 0, 2, -1, 4, -2, -3 };

// prob.h:79 MODE_MV_MAX_UPDATE_FACTOR(128) * count / MODE_MV_COUNT_SAT(20)
const int kCountToUpdateFactor[21] = { 0, 6, 12, 19, 25, 32, 38, 44, 51, 57, 64,
                                       70, 76, 83, 89, 96, 102, 108, 115, 121, 128 };

inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

}  // namespace

// ------------------------------------------------------------------------------------------
// Boolean (arithmetic) decoder.  Same arithmetic as vpx_dsp/bitreader.h vpx_read():
// split = (range * prob + (256 - prob)) >> 8, renormalise by leading zeros of range.  Bytes
// past the end of the partition read as zero (bitreader.c:54-84).
class BoolReader {
 public:
  int init(const uint8_t *buf, size_t size) {
    buf_ = buf;
    end_ = buf + size;
    value_ = 0;
    count_ = -8;
    range_ = 255;
    overrun_ = 0;
    fill();
    return read(128);  // marker bit, must be 0
  }
  inline void fill() {
    int shift = 64 - 8 - (count_ + 8);
    if (end_ - buf_ >= 8 && shift >= 0) {   // whole bytes from one big-endian 64-bit load
      uint64_t w;
      memcpy(&w, buf_, 8);
      w = __builtin_bswap64(w);
      const int nbytes = (shift >> 3) + 1;            // 1..8 bytes fit
      const int nbits = nbytes * 8;
      value_ |= (w >> (64 - nbits)) << (shift + 8 - nbits);
      buf_ += nbytes;
      count_ += nbits;
      return;
    }
    while (shift >= 0) {
      uint64_t b = 0;
      if (buf_ < end_) b = *buf_++; else overrun_++;
      count_ += 8;
      value_ |= b << shift;
      shift -= 8;
    }
  }
  static inline int norm_shift(uint32_t range) {
    return __builtin_clz(range) - 24;   // range is 1..255 (LZCNT; the table of bitreader.c costs a load)
  }
  inline int read(int prob) {
    const uint32_t split = (range_ * (uint32_t)prob + (256 - (uint32_t)prob)) >> 8;
    if (count_ < 0) fill();
    const uint64_t bigsplit = (uint64_t)split << 56;
    int bit;
    if (value_ >= bigsplit) {
      range_ -= split;
      value_ -= bigsplit;
      bit = 1;
    } else {
      range_ = split;
      bit = 0;
    }
    const int shift = norm_shift(range_);
    range_ <<= shift;
    value_ <<= shift;
    count_ -= shift;
    return bit;
  }
  // Same arithmetic without a data-dependent branch: for symbols whose value only feeds arithmetic
  // (signs, the extra bits of the category tokens) -- a coin-flip branch there costs more than the
  // select.
  inline int read_nb(int prob) {
    const uint32_t split = (range_ * (uint32_t)prob + (256 - (uint32_t)prob)) >> 8;
    if (count_ < 0) fill();
    const uint64_t bigsplit = (uint64_t)split << 56;
    const uint64_t m = (uint64_t)0 - (uint64_t)(value_ >= bigsplit);
    range_ = split + ((range_ - 2 * split) & (uint32_t)m);
    value_ -= bigsplit & m;
    const int shift = norm_shift(range_);
    range_ <<= shift;
    value_ <<= shift;
    count_ -= shift;
    return (int)(m & 1);
  }
  inline int read_bit() { return read(128); }
  inline int read_literal(int bits) {
    int v = 0;
    for (int b = bits - 1; b >= 0; --b) v |= read_bit() << b;
    return v;
  }
  inline int read_tree(const int8_t *tree, const uint8_t *probs) {
    int i = 0;
    while ((i = tree[i + read(probs[i >> 1])]) > 0) {
    }
    return -i;
  }
  // vpx_reader_has_error (bitreader.h:60-76): bits were consumed past the end of the data.
  long bytes_left() const { return (long)(end_ - buf_) + (count_ >= 0 ? (count_ + 8) / 8 : 0); }
  bool has_error() const {
    if (!overrun_) return false;
    return (long)overrun_ * 8 - (count_ + 8) > 0;
  }

 private:
  const uint8_t *buf_, *end_;
  uint64_t value_;
  int count_;
  uint32_t range_;
  int overrun_;
};

// Plain MSB-first bit reader for the uncompressed header (vpx_dsp/bitreader_buffer.c).
struct BitBuf {
  const uint8_t *p;
  size_t size_bits, pos;
  bool err;
  int bit() {
    if (pos >= size_bits) { err = true; return 0; }
    int b = (p[pos >> 3] >> (7 - (pos & 7))) & 1;
    pos++;
    return b;
  }
  int lit(int n) {
    int v = 0;
    for (int i = 0; i < n; i++) v = (v << 1) | bit();
    return v;
  }
  int signed_lit(int n) {  // vpx_rb_read_signed_literal
    int v = lit(n);
    return bit() ? -v : v;
  }
};

// ------------------------------------------------------------------------------------------
Parser::Parser() {
  for (int i = 0; i < 8; i++) ref_frame_map_[i] = -1;
  memset(frame_contexts_, 0, sizeof(frame_contexts_));
  memset(&fc_, 0, sizeof(fc_));
  memset(&counts_, 0, sizeof(counts_));
  memset(&seg_, 0, sizeof(seg_));
  memset(&lf_, 0, sizeof(lf_));
  memset(ref_slot_, -1, sizeof(ref_slot_));
  memset(ref_sign_bias_, 0, sizeof(ref_sign_bias_));
  memset(ref_desc_, 0, sizeof(ref_desc_));
}
Parser::~Parser() { adapt_worker_.wait(); }

int Parser::fail(int code, const char *msg) {
  err_ = msg;
  return code;
}

int Parser::find_free_slot() {
  for (int s = 0; s < VP9B_NUM_SLOTS; s++) {
    bool used = (s == last_shown_slot_);
    for (int i = 0; i < 8 && !used; i++) used = (ref_frame_map_[i] == s);
    if (!used) return s;
  }
  return -1;
}

// resize_context_buffers (vp9_decodeframe.c:1705-1741) + vp9_set_mb_mi
void Parser::setup_frame_size(int w, int h) {
  if (w != width_ || h != height_) {
    width_ = w;
    height_ = h;
    mi_cols_ = (w + 7) >> 3;
    mi_rows_ = (h + 7) >> 3;
    sb_cols_ = (mi_cols_ + 7) >> 3;
    sb_rows_ = (mi_rows_ + 7) >> 3;
    grid_.assign((size_t)mi_cols_ * mi_rows_, 0);
    seg_map_cur_.assign((size_t)mi_cols_ * mi_rows_, 0);
    seg_map_last_.assign((size_t)mi_cols_ * mi_rows_, 0);
    cur_mvs_.assign((size_t)mi_cols_ * mi_rows_, MvRef());
    prev_mvs_.assign((size_t)mi_cols_ * mi_rows_, MvRef());
    prev_mvs_valid_ = false;
  }
}

// vp9_setup_past_independence (vp9_entropymode.c:427-470)
void Parser::setup_past_independence() {
  memset(seg_.feature_data, 0, sizeof(seg_.feature_data));
  memset(seg_.feature_mask, 0, sizeof(seg_.feature_mask));
  seg_.abs_delta = 0;
  std::fill(seg_map_last_.begin(), seg_map_last_.end(), 0);
  std::fill(seg_map_cur_.begin(), seg_map_cur_.end(), 0);
  lf_.mode_ref_delta_enabled = 1;
  lf_.mode_ref_delta_update = 1;
  lf_.ref_deltas[INTRA_FRAME] = 1;
  lf_.ref_deltas[LAST_FRAME] = 0;
  lf_.ref_deltas[GOLDEN_FRAME] = -1;
  lf_.ref_deltas[ALTREF_FRAME] = -1;
  lf_.mode_deltas[0] = lf_.mode_deltas[1] = 0;
  // default probabilities
  memcpy(fc_.coef, kDefaultCoefProbs, sizeof(fc_.coef));
  memcpy(fc_.uv_mode, kDefaultIfUvProbs, sizeof(fc_.uv_mode));
  memcpy(fc_.y_mode, kDefaultIfYProbs, sizeof(fc_.y_mode));
  memcpy(fc_.switchable_interp, kDefaultSwitchableInterpProb, sizeof(fc_.switchable_interp));
  memcpy(fc_.partition, kDefaultPartitionProbs, sizeof(fc_.partition));
  memcpy(fc_.intra_inter, kDefaultIntraInterP, sizeof(fc_.intra_inter));
  memcpy(fc_.comp_inter, kDefaultCompInterP, sizeof(fc_.comp_inter));
  memcpy(fc_.comp_ref, kDefaultCompRefP, sizeof(fc_.comp_ref));
  memcpy(fc_.single_ref, kDefaultSingleRefP, sizeof(fc_.single_ref));
  memcpy(fc_.tx32, kDefaultTxProbs32, sizeof(fc_.tx32));
  memcpy(fc_.tx16, kDefaultTxProbs16, sizeof(fc_.tx16));
  memcpy(fc_.tx8, kDefaultTxProbs8, sizeof(fc_.tx8));
  memcpy(fc_.skip, kDefaultSkipProbs, sizeof(fc_.skip));
  memcpy(fc_.inter_mode, kDefaultInterModeProbs, sizeof(fc_.inter_mode));
  memcpy(fc_.mv_joints, kDefaultMvJoints, 3);
  for (int c = 0; c < 2; c++) {
    const uint8_t *s = kDefaultMvComps + c * 33;
    MvCompProbs &m = fc_.mvc[c];
    m.sign = s[0];
    memcpy(m.classes, s + 1, 10);
    m.class0[0] = s[11];
    memcpy(m.bits, s + 12, 10);
    memcpy(m.class0_fp, s + 22, 6);
    memcpy(m.fp, s + 28, 3);
    m.class0_hp = s[31];
    m.hp = s[32];
  }
  fc_.initialized = 1;
  if (frame_type_ == KEY_FRAME || error_res_ || reset_frame_context_ == 3) {
    for (int i = 0; i < 4; i++) frame_contexts_[i] = fc_;
  } else if (reset_frame_context_ == 2) {
    frame_contexts_[frame_context_idx_] = fc_;
  }
  memset(ref_sign_bias_, 0, sizeof(ref_sign_bias_));
  frame_context_idx_ = 0;
}

// read_uncompressed_header (vp9_decodeframe.c:3114-3338)
int Parser::read_uncompressed_header(const uint8_t *data, size_t size, size_t *hdr_bytes,
                                     size_t *first_partition_size) {
  BitBuf rb = { data, size * 8, 0, false };
  last_frame_type_ = frame_type_;
  last_intra_only_ = intra_only_;
  if (rb.lit(2) != 2) return fail(VP9B_ERR_UNSUP, "invalid frame marker");
  profile_ = rb.bit();
  profile_ |= rb.bit() << 1;
  if (profile_ > 2) profile_ += rb.bit();
  if (profile_ > 3) return fail(VP9B_ERR_UNSUP, "unsupported profile");
  show_existing_ = rb.bit();
  *first_partition_size = 0;
  if (show_existing_) {
    const int idx = rb.lit(3);
    const int s = ref_frame_map_[idx];
    if (s < 0) return fail(VP9B_ERR_CORRUPT, "show_existing_frame of an empty slot");
    show_slot_ = s;
    refresh_frame_flags_ = 0;
    lf_.filter_level = 0;
    show_frame_ = 1;
    *hdr_bytes = (rb.pos + 7) >> 3;
    return VP9B_OK;
  }
  frame_type_ = rb.bit();
  show_frame_ = rb.bit();
  error_res_ = rb.bit();
  auto read_sync = [&]() { return rb.lit(8) == 0x49 && rb.lit(8) == 0x83 && rb.lit(8) == 0x42; };
  // read_bitdepth_colorspace_sampling (vp9_decodeframe.c:3054-3096)
  auto read_bitdepth_cs = [&]() -> int {
    if (profile_ >= 2) bit_depth_ = rb.bit() ? 12 : 10; else bit_depth_ = 8;
    const int cs = rb.lit(3);
    if (cs != 7 /* VPX_CS_SRGB */) {
      rb.bit();  // color range
      if (profile_ == 1 || profile_ == 3) {
        ss_x_ = rb.bit();
        ss_y_ = rb.bit();
        if (ss_x_ == 1 && ss_y_ == 1) return fail(VP9B_ERR_UNSUP, "4:2:0 in profile 1/3");
        if (rb.bit()) return fail(VP9B_ERR_UNSUP, "reserved bit set");
      } else {
        ss_x_ = ss_y_ = 1;
      }
    } else {
      if (profile_ == 1 || profile_ == 3) {
        ss_x_ = ss_y_ = 0;
        if (rb.bit()) return fail(VP9B_ERR_UNSUP, "reserved bit set");
      } else {
        return fail(VP9B_ERR_UNSUP, "4:4:4 in profile 0/2");
      }
    }
    return VP9B_OK;
  };
  for (int i = 0; i < 3; i++) { ref_slot_[i] = -1; ref_desc_[i].slot = -1; }
  int w = 0, h = 0;
  if (frame_type_ == KEY_FRAME) {
    if (!read_sync()) return fail(VP9B_ERR_UNSUP, "invalid sync code");
    int e = read_bitdepth_cs();
    if (e) return e;
    refresh_frame_flags_ = 0xff;
    w = rb.lit(16) + 1;
    h = rb.lit(16) + 1;
    if (rb.bit()) rb.lit(32);  // render size
    if (w > kMaxFrameDim || h > kMaxFrameDim) return fail(VP9B_ERR_UNSUP, "frame larger than the decoder limit");
    setup_frame_size(w, h);
  } else {
    intra_only_ = show_frame_ ? 0 : rb.bit();
    reset_frame_context_ = error_res_ ? 0 : rb.lit(2);
    if (intra_only_) {
      if (!read_sync()) return fail(VP9B_ERR_UNSUP, "invalid sync code");
      if (profile_ > 0) {
        int e = read_bitdepth_cs();
        if (e) return e;
      } else {
        ss_x_ = ss_y_ = 1;
        bit_depth_ = 8;
      }
      refresh_frame_flags_ = rb.lit(8);
      w = rb.lit(16) + 1;
      h = rb.lit(16) + 1;
      if (rb.bit()) rb.lit(32);
      if (w > kMaxFrameDim || h > kMaxFrameDim) return fail(VP9B_ERR_UNSUP, "frame larger than the decoder limit");
      setup_frame_size(w, h);
    } else {
      refresh_frame_flags_ = rb.lit(8);
      for (int i = 0; i < 3; i++) {
        const int ref = rb.lit(3);
        ref_slot_[i] = ref_frame_map_[ref];
        ref_sign_bias_[LAST_FRAME + i] = rb.bit();
        if (ref_slot_[i] < 0 || !slots_[ref_slot_[i]].valid)
          return fail(VP9B_ERR_CORRUPT, "reference to an empty frame slot");
      }
      // setup_frame_size_with_refs (vp9_decodeframe.c:1781-1855)
      bool found = false;
      for (int i = 0; i < 3; i++) {
        if (rb.bit()) {
          w = slots_[ref_slot_[i]].width;
          h = slots_[ref_slot_[i]].height;
          found = true;
          break;
        }
      }
      if (!found) {
        w = rb.lit(16) + 1;
        h = rb.lit(16) + 1;
      }
      bool any_valid = false;
      for (int i = 0; i < 3; i++) {
        const SlotInfo &s = slots_[ref_slot_[i]];
        // valid_ref_frame_size (vp9_scale.h:62-66)
        any_valid |= (2 * w >= (int)s.width && 2 * h >= (int)s.height && w <= 16 * (int)s.width &&
                      h <= 16 * (int)s.height);
        if (s.bit_depth != bit_depth_ || s.ss_x != ss_x_ || s.ss_y != ss_y_)
          return fail(VP9B_ERR_CORRUPT, "reference frame has incompatible colour format");
      }
      if (!any_valid) return fail(VP9B_ERR_CORRUPT, "referenced frame has invalid size");
      if (w > kMaxFrameDim || h > kMaxFrameDim) return fail(VP9B_ERR_UNSUP, "frame larger than the decoder limit");
      setup_frame_size(w, h);
      if (rb.bit()) rb.lit(32);  // render size
      allow_hp_ = rb.bit();
      // read_interp_filter (vp9_decodeframe.c:1682-1687)
      static const uint8_t literal_to_filter[4] = { EIGHTTAP_SMOOTH, EIGHTTAP, EIGHTTAP_SHARP,
                                                    BILINEAR };
      interp_filter_ = rb.bit() ? SWITCHABLE : literal_to_filter[rb.lit(2)];
      // vp9_setup_scale_factors_for_frame (vp9_scale.c:46-77)
      for (int i = 0; i < 3; i++) {
        const SlotInfo &s = slots_[ref_slot_[i]];
        vp9b_ref_desc &d = ref_desc_[i];
        d.slot = ref_slot_[i];
        d.width = s.width;
        d.height = s.height;
        const bool valid = (2 * w >= (int)s.width && 2 * h >= (int)s.height &&
                            w <= 16 * (int)s.width && h <= 16 * (int)s.height);
        if (!valid) {
          d.x_scale_fp = d.y_scale_fp = -1;
          d.x_step_q4 = d.y_step_q4 = 16;
        } else {
          d.x_scale_fp = (int)(((int64_t)s.width << 14) / w);
          d.y_scale_fp = (int)(((int64_t)s.height << 14) / h);
          d.x_step_q4 = (int)(((int64_t)16 * d.x_scale_fp) >> 14);
          d.y_step_q4 = (int)(((int64_t)16 * d.y_scale_fp) >> 14);
        }
      }
    }
  }
  if (rb.err) return fail(VP9B_ERR_CORRUPT, "truncated header");
  if (!error_res_) {
    refresh_frame_context_ = rb.bit();
    frame_parallel_ = rb.bit();
  } else {
    refresh_frame_context_ = 0;
    frame_parallel_ = 1;
  }
  frame_context_idx_ = rb.lit(2);
  if (frame_type_ == KEY_FRAME || intra_only_ || error_res_) setup_past_independence();
  // setup_loopfilter (vp9_decodeframe.c:1610-1634)
  lf_.filter_level = rb.lit(6);
  lf_.sharpness_level = rb.lit(3);
  lf_.mode_ref_delta_update = 0;
  lf_.mode_ref_delta_enabled = rb.bit();
  if (lf_.mode_ref_delta_enabled) {
    lf_.mode_ref_delta_update = rb.bit();
    if (lf_.mode_ref_delta_update) {
      for (int i = 0; i < 4; i++)
        if (rb.bit()) lf_.ref_deltas[i] = (int8_t)rb.signed_lit(6);
      for (int i = 0; i < 2; i++)
        if (rb.bit()) lf_.mode_deltas[i] = (int8_t)rb.signed_lit(6);
    }
  }
  // setup_quantization (vp9_decodeframe.c:1640-1653)
  base_qindex_ = rb.lit(8);
  y_dc_delta_q_ = rb.bit() ? rb.signed_lit(4) : 0;
  uv_dc_delta_q_ = rb.bit() ? rb.signed_lit(4) : 0;
  uv_ac_delta_q_ = rb.bit() ? rb.signed_lit(4) : 0;
  lossless_ = base_qindex_ == 0 && y_dc_delta_q_ == 0 && uv_dc_delta_q_ == 0 && uv_ac_delta_q_ == 0;
  // setup_segmentation (vp9_decodeframe.c:1560-1608)
  seg_.update_map = 0;
  seg_.update_data = 0;
  seg_.enabled = rb.bit();
  if (seg_.enabled) {
    seg_.update_map = rb.bit();
    if (seg_.update_map) {
      for (int i = 0; i < 7; i++) seg_.tree_probs[i] = rb.bit() ? rb.lit(8) : 255;
      seg_.temporal_update = rb.bit();
      for (int i = 0; i < 3; i++)
        seg_.pred_probs[i] = seg_.temporal_update ? (rb.bit() ? rb.lit(8) : 255) : 255;
    }
    seg_.update_data = rb.bit();
    if (seg_.update_data) {
      seg_.abs_delta = rb.bit();
      memset(seg_.feature_data, 0, sizeof(seg_.feature_data));
      memset(seg_.feature_mask, 0, sizeof(seg_.feature_mask));
      static const int feat_max[4] = { 255, 63, 3, 0 };
      static const int feat_bits[4] = { 8, 6, 2, 0 };
      static const int feat_signed[4] = { 1, 1, 0, 0 };
      for (int i = 0; i < 8; i++)
        for (int j = 0; j < 4; j++) {
          int v = 0;
          if (rb.bit()) {
            seg_.feature_mask[i] |= 1u << j;
            v = rb.lit(feat_bits[j]);  // decode_unsigned_max
            if (v > feat_max[j]) v = feat_max[j];
            if (feat_signed[j]) v = rb.bit() ? -v : v;
          }
          seg_.feature_data[i][j] = (int16_t)v;
        }
    }
  }
  // setup_segmentation_dequant (vp9_decodeframe.c:1655-1680)
  {
    const int bdi = bit_depth_ == 8 ? 0 : (bit_depth_ == 10 ? 1 : 2);
    const int16_t *dcq = kDcQLookup + bdi * 256, *acq = kAcQLookup + bdi * 256;
    for (int i = 0; i < 8; i++) {
      int q = base_qindex_;
      if (seg_.enabled && (seg_.feature_mask[i] & (1u << SEG_LVL_ALT_Q))) {  // vp9_get_qindex
        const int d = seg_.feature_data[i][SEG_LVL_ALT_Q];
        q = clampi(seg_.abs_delta ? d : base_qindex_ + d, 0, 255);
      }
      y_dequant_[i][0] = dcq[clampi(q + y_dc_delta_q_, 0, 255)];
      y_dequant_[i][1] = acq[clampi(q, 0, 255)];
      uv_dequant_[i][0] = dcq[clampi(q + uv_dc_delta_q_, 0, 255)];
      uv_dequant_[i][1] = acq[clampi(q + uv_ac_delta_q_, 0, 255)];
    }
  }
  // setup_tile_info (vp9_decodeframe.c:1857-1873)
  {
    int min_log2 = 0, max_log2 = 1;
    while ((64 << min_log2) < sb_cols_) ++min_log2;
    while ((sb_cols_ >> max_log2) >= 4) ++max_log2;
    --max_log2;
    int max_ones = max_log2 - min_log2;
    log2_tile_cols_ = min_log2;
    while (max_ones-- > 0 && rb.bit()) log2_tile_cols_++;
    log2_tile_rows_ = rb.bit();
    if (log2_tile_rows_) log2_tile_rows_ += rb.bit();
  }
  const int sz = rb.lit(16);
  if (rb.err) return fail(VP9B_ERR_CORRUPT, "truncated header");
  if (sz == 0) return fail(VP9B_ERR_CORRUPT, "invalid header size");
  *first_partition_size = sz;
  *hdr_bytes = (rb.pos + 7) >> 3;
  return VP9B_OK;
}

namespace {
// vp9_dsubexp.c:17-76
inline int inv_recenter_nonneg(int v, int m) {
  if (v > 2 * m) return v;
  return (v & 1) ? m - ((v + 1) >> 1) : m + (v >> 1);
}
inline void diff_update_prob(BoolReader &r, uint8_t *p) {
  if (r.read(252)) {
    int delp;
    if (!r.read_bit()) delp = r.read_literal(4);
    else if (!r.read_bit()) delp = r.read_literal(4) + 16;
    else if (!r.read_bit()) delp = r.read_literal(5) + 32;
    else {
      const int v = r.read_literal(7);
      delp = (v < 65 ? v : (v << 1) - 65 + r.read_bit()) + 64;
    }
    int v = kInvMapTable[delp];
    int m = *p;
    m--;
    if ((m << 1) <= 255) *p = (uint8_t)(1 + inv_recenter_nonneg(v, m));
    else *p = (uint8_t)(255 - inv_recenter_nonneg(v, 255 - 1 - m));
  }
}
inline void update_mv_probs(BoolReader &r, uint8_t *p, int n) {
  for (int i = 0; i < n; i++)
    if (r.read(252)) p[i] = (uint8_t)((r.read_literal(7) << 1) | 1);
}
}  // namespace

// read_compressed_header (vp9_decodeframe.c:3340-3388)
int Parser::read_compressed_header(const uint8_t *data, size_t size) {
  BoolReader r;
  if (r.init(data, size)) return fail(VP9B_ERR_CORRUPT, "compressed header marker bit set");
  // read_tx_mode
  if (lossless_) {
    tx_mode_ = ONLY_4X4;
  } else {
    tx_mode_ = r.read_literal(2);
    if (tx_mode_ == ALLOW_32X32) tx_mode_ += r.read_bit();
  }
  if (tx_mode_ == TX_MODE_SELECT) {
    for (int i = 0; i < 2; i++) diff_update_prob(r, &fc_.tx8[i][0]);
    for (int i = 0; i < 2; i++)
      for (int j = 0; j < 2; j++) diff_update_prob(r, &fc_.tx16[i][j]);
    for (int i = 0; i < 2; i++)
      for (int j = 0; j < 3; j++) diff_update_prob(r, &fc_.tx32[i][j]);
  }
  // read_coef_probs
  const int max_tx = kTxModeToBiggest[tx_mode_];
  for (int t = 0; t <= max_tx; t++) {
    if (r.read_bit())
      for (int i = 0; i < 2; i++)
        for (int j = 0; j < 2; j++)
          for (int k = 0; k < 6; k++)
            for (int l = 0; l < (k == 0 ? 3 : 6); l++)
              for (int m = 0; m < 3; m++) diff_update_prob(r, &fc_.coef[t][i][j][k][l][m]);
  }
  for (int k = 0; k < 3; k++) diff_update_prob(r, &fc_.skip[k]);
  if (!(frame_type_ == KEY_FRAME || intra_only_)) {
    for (int i = 0; i < 7; i++)
      for (int j = 0; j < 3; j++) diff_update_prob(r, &fc_.inter_mode[i][j]);
    if (interp_filter_ == SWITCHABLE)
      for (int j = 0; j < 4; j++)
        for (int i = 0; i < 2; i++) diff_update_prob(r, &fc_.switchable_interp[j][i]);
    for (int i = 0; i < 4; i++) diff_update_prob(r, &fc_.intra_inter[i]);
    // read_frame_reference_mode (vp9_decodeframe.c:108-117), vp9_pred_common.c:14-40
    bool comp_allowed = false;
    for (int i = 1; i < 3; i++)
      if (ref_sign_bias_[i + 1] != ref_sign_bias_[1]) comp_allowed = true;
    if (comp_allowed)
      reference_mode_ = r.read_bit() ? (r.read_bit() ? REFERENCE_MODE_SELECT : COMPOUND_REFERENCE)
                                     : SINGLE_REFERENCE;
    else
      reference_mode_ = SINGLE_REFERENCE;
    if (reference_mode_ != SINGLE_REFERENCE) {
      if (ref_sign_bias_[LAST_FRAME] == ref_sign_bias_[GOLDEN_FRAME]) {
        comp_fixed_ref_ = ALTREF_FRAME;
        comp_var_ref_[0] = LAST_FRAME;
        comp_var_ref_[1] = GOLDEN_FRAME;
      } else if (ref_sign_bias_[LAST_FRAME] == ref_sign_bias_[ALTREF_FRAME]) {
        comp_fixed_ref_ = GOLDEN_FRAME;
        comp_var_ref_[0] = LAST_FRAME;
        comp_var_ref_[1] = ALTREF_FRAME;
      } else {
        comp_fixed_ref_ = LAST_FRAME;
        comp_var_ref_[0] = GOLDEN_FRAME;
        comp_var_ref_[1] = ALTREF_FRAME;
      }
    }
    if (reference_mode_ == REFERENCE_MODE_SELECT)
      for (int i = 0; i < 5; i++) diff_update_prob(r, &fc_.comp_inter[i]);
    if (reference_mode_ != COMPOUND_REFERENCE)
      for (int i = 0; i < 5; i++) {
        diff_update_prob(r, &fc_.single_ref[i][0]);
        diff_update_prob(r, &fc_.single_ref[i][1]);
      }
    if (reference_mode_ != SINGLE_REFERENCE)
      for (int i = 0; i < 5; i++) diff_update_prob(r, &fc_.comp_ref[i]);
    for (int j = 0; j < 4; j++)
      for (int i = 0; i < 9; i++) diff_update_prob(r, &fc_.y_mode[j][i]);
    for (int j = 0; j < 16; j++)
      for (int i = 0; i < 3; i++) diff_update_prob(r, &fc_.partition[j][i]);
    // read_mv_probs (vp9_decodeframe.c:144-171)
    update_mv_probs(r, fc_.mv_joints, 3);
    for (int i = 0; i < 2; i++) {
      MvCompProbs &c = fc_.mvc[i];
      update_mv_probs(r, &c.sign, 1);
      update_mv_probs(r, c.classes, 10);
      update_mv_probs(r, c.class0, 1);
      update_mv_probs(r, c.bits, 10);
    }
    for (int i = 0; i < 2; i++) {
      MvCompProbs &c = fc_.mvc[i];
      for (int j = 0; j < 2; j++) update_mv_probs(r, c.class0_fp[j], 3);
      update_mv_probs(r, c.fp, 3);
    }
    if (allow_hp_)
      for (int i = 0; i < 2; i++) {
        update_mv_probs(r, &fc_.mvc[i].class0_hp, 1);
        update_mv_probs(r, &fc_.mvc[i].hp, 1);
      }
  }
  if (r.has_error()) return fail(VP9B_ERR_CORRUPT, "compressed header overrun");
  return VP9B_OK;
}

// vp9_loop_filter_frame_init + update_sharpness (vp9_loopfilter.c:212-295)
void Parser::loop_filter_frame_init() {
  const int lvl0 = lf_.filter_level;
  const int scale = 1 << (lvl0 >> 5);
  for (int s = 0; s < 8; s++) {
    int lvl_seg = lvl0;
    if (seg_.enabled && (seg_.feature_mask[s] & (1u << SEG_LVL_ALT_LF))) {
      const int d = seg_.feature_data[s][SEG_LVL_ALT_LF];
      lvl_seg = clampi(seg_.abs_delta ? d : lvl0 + d, 0, 63);
    }
    if (!lf_.mode_ref_delta_enabled) {
      memset(lf_lvl_[s], lvl_seg, sizeof(lf_lvl_[s]));
    } else {
      const int intra_lvl = lvl_seg + lf_.ref_deltas[INTRA_FRAME] * scale;
      lf_lvl_[s][INTRA_FRAME][0] = (uint8_t)clampi(intra_lvl, 0, 63);
      lf_lvl_[s][INTRA_FRAME][1] = 0;  // never read: intra modes map to mode_lf_lut 0
      for (int ref = LAST_FRAME; ref < 4; ++ref)
        for (int m = 0; m < 2; ++m) {
          const int v = lvl_seg + lf_.ref_deltas[ref] * scale + lf_.mode_deltas[m] * scale;
          lf_lvl_[s][ref][m] = (uint8_t)clampi(v, 0, 63);
        }
    }
  }
}

// ------------------------------------------------------------------------------------------
// Per-tile decoding context (role of TileWorkerData + MACROBLOCKD).
struct TileCtx {
  Parser *P;
  TileOut *O;      // what this tile column's worker appends to (merged by Parser::pack)
  BoolReader r;
  int mi_row_start, mi_row_end, mi_col_start, mi_col_end;
  uint8_t left_ctx[3][16];
  uint8_t left_seg_ctx[8];
  bool corrupted = false;
  // current block
  int mi_row, mi_col, bw, bh, bwl, bhl, x_mis, y_mis;
  int mb_to_left_edge, mb_to_right_edge, mb_to_top_edge, mb_to_bottom_edge;
  const BlockInfo *above_mi, *left_mi;
  BlockInfo *mi;
  int bmode_blocks_wl, bmode_blocks_hl;
  int max_blocks_wide, max_blocks_high;  // xd->max_blocks_* (0 when block is inside the frame)
  uint8_t token_cache[1024];

  const BlockInfo *at(int row, int col) const {
    return &P->blocks_[P->grid_[(size_t)row * P->mi_cols_ + col]];
  }
  bool is_inter(const BlockInfo *b) const { return b->ref_frame[0] > INTRA_FRAME; }
  bool has_second(const BlockInfo *b) const { return b->ref_frame[1] > INTRA_FRAME; }
  bool intra_only_frame() const { return P->frame_type_ == KEY_FRAME || P->intra_only_; }

  void decode_tile();
  void decode_partition(int mi_row, int mi_col, int bsize, int n4x4_l2);
  void decode_block(int mi_row, int mi_col, int bsize, int bwl, int bhl);
  void read_intra_frame_mode_info();
  void read_inter_frame_mode_info();
  void read_intra_block_mode_info();
  void read_inter_block_mode_info();
  int read_intra_segment_id();
  int read_inter_segment_id();
  int read_skip(int segment_id);
  int read_tx_size(int allow_select);
  void read_ref_frames(int8_t ref_frame[2]);
  int read_mv_component(const MvCompProbs &c, int usehp);
  void read_mv(Mv *mv, const Mv *ref);
  bool assign_mv(int mode, Mv mv[2], const Mv ref_mv[2], const Mv near_nearest[2], int is_compound);
  int find_mv_refs(int mode, int ref_frame, const int8_t *search, Mv *list, int block);
  void append_sub8x8_mvs_for_idx(const int8_t *search, int b_mode, int block, int ref, Mv *best);
  int get_mode_context(const int8_t *search);
  int decode_block_tokens(int plane, int x, int y, int tx_size, int tx_type, uint32_t *coef_off);
  template <bool HBD, bool COUNTS>
  int decode_block_tokens_t(int plane, int x, int y, int tx_size, int tx_type, uint32_t *coef_off);
  void residual_and_record();
  void build_lf_mask();
  bool seg_active(int seg, int feat) const {
    return P->seg_.enabled && (P->seg_.feature_mask[seg] & (1u << feat));
  }
};

void TileCtx::decode_tile() {
  // decode_tiles tile loop (vp9_decodeframe.c:2390-2430)
  for (int r0 = mi_row_start; r0 < mi_row_end; r0 += 8) {
    memset(left_ctx, 0, sizeof(left_ctx));
    memset(left_seg_ctx, 0, sizeof(left_seg_ctx));
    for (int c0 = mi_col_start; c0 < mi_col_end; c0 += 8) {
      const int sb = (r0 >> 3) * P->sb_cols_ + (c0 >> 3);
      const int nsb = P->sb_cols_ * P->sb_rows_;
      for (int p = 0; p < 3; p++)
        P->out_sb_intra_[p * nsb + sb].start = (uint32_t)O->txb_intra[p].size();   // index inside this worker's list
      decode_partition(r0, c0, BLOCK_64X64, 4);
      for (int p = 0; p < 3; p++)
        P->out_sb_intra_[p * nsb + sb].count =
            (uint32_t)O->txb_intra[p].size() - P->out_sb_intra_[p * nsb + sb].start;
      if (corrupted) return;
    }
  }
}

// decode_partition / read_partition (vp9_decodeframe.c:1362-1453)
void TileCtx::decode_partition(int row, int col, int bsize, int n4x4_l2) {
  const int n8x8_l2 = n4x4_l2 - 1;
  const int num_8x8_wh = 1 << n8x8_l2;
  const int hbs = num_8x8_wh >> 1;
  if (row >= P->mi_rows_ || col >= P->mi_cols_) return;
  const int has_rows = (row + hbs) < P->mi_rows_;
  const int has_cols = (col + hbs) < P->mi_cols_;
  // dec_partition_plane_context
  uint8_t *above_ctx = &P->above_seg_ctx_[col];
  uint8_t *lctx = &left_seg_ctx[row & 7];
  const int above = (*above_ctx >> n8x8_l2) & 1, left = (*lctx >> n8x8_l2) & 1;
  const int ctx = (left * 2 + above) + n8x8_l2 * 4;
  const uint8_t *probs = intra_only_frame() ? &kKfPartitionProbs[ctx * 3] : P->fc_.partition[ctx];
  int p;
  if (has_rows && has_cols) p = r.read_tree(kPartitionTree, probs);
  else if (!has_rows && has_cols) p = r.read(probs[1]) ? PARTITION_SPLIT : PARTITION_HORZ;
  else if (has_rows && !has_cols) p = r.read(probs[2]) ? PARTITION_SPLIT : PARTITION_VERT;
  else p = PARTITION_SPLIT;
  if (P->do_counts_) ++O->counts.partition[ctx][p];
  const int subsize = kSubsize[p * 13 + bsize];
  if (!hbs) {
    bmode_blocks_wl = 1 >> !!(p & PARTITION_VERT);
    bmode_blocks_hl = 1 >> !!(p & PARTITION_HORZ);
    decode_block(row, col, subsize, 1, 1);
  } else {
    switch (p) {
      case PARTITION_NONE: decode_block(row, col, subsize, n4x4_l2, n4x4_l2); break;
      case PARTITION_HORZ:
        decode_block(row, col, subsize, n4x4_l2, n8x8_l2);
        if (has_rows) decode_block(row + hbs, col, subsize, n4x4_l2, n8x8_l2);
        break;
      case PARTITION_VERT:
        decode_block(row, col, subsize, n8x8_l2, n4x4_l2);
        if (has_cols) decode_block(row, col + hbs, subsize, n8x8_l2, n4x4_l2);
        break;
      case PARTITION_SPLIT:
        decode_partition(row, col, subsize, n8x8_l2);
        decode_partition(row, col + hbs, subsize, n8x8_l2);
        decode_partition(row + hbs, col, subsize, n8x8_l2);
        decode_partition(row + hbs, col + hbs, subsize, n8x8_l2);
        break;
    }
  }
  if (corrupted) return;
  // dec_update_partition_context
  if (bsize >= BLOCK_8X8 && (bsize == BLOCK_8X8 || p != PARTITION_SPLIT)) {
    memset(above_ctx, kPartitionContext[subsize * 2 + 0], num_8x8_wh);
    memset(lctx, kPartitionContext[subsize * 2 + 1], num_8x8_wh);
  }
}

// read_intra_segment_id / read_inter_segment_id (vp9_decodemv.c:139-189)
int TileCtx::read_intra_segment_id() {
  Segmentation &seg = P->seg_;
  if (!seg.enabled) return 0;
  const int off = mi_row * P->mi_cols_ + mi_col;
  if (!seg.update_map) {
    for (int y = 0; y < y_mis; y++)
      for (int x = 0; x < x_mis; x++)
        P->seg_map_cur_[off + y * P->mi_cols_ + x] = P->seg_map_last_[off + y * P->mi_cols_ + x];
    return 0;
  }
  const int sid = r.read_tree(kSegmentTree, seg.tree_probs);
  for (int y = 0; y < y_mis; y++)
    for (int x = 0; x < x_mis; x++) P->seg_map_cur_[off + y * P->mi_cols_ + x] = (uint8_t)sid;
  return sid;
}

int TileCtx::read_inter_segment_id() {
  Segmentation &seg = P->seg_;
  if (!seg.enabled) return 0;
  const int off = mi_row * P->mi_cols_ + mi_col;
  int pred = 255;
  for (int y = 0; y < y_mis; y++)
    for (int x = 0; x < x_mis; x++)
      pred = std::min<int>(pred, P->seg_map_last_[off + y * P->mi_cols_ + x]);
  if (!seg.update_map) {
    for (int y = 0; y < y_mis; y++)
      for (int x = 0; x < x_mis; x++)
        P->seg_map_cur_[off + y * P->mi_cols_ + x] = P->seg_map_last_[off + y * P->mi_cols_ + x];
    return pred;
  }
  int sid;
  if (seg.temporal_update) {
    const int a = above_mi ? above_mi->seg_id_predicted : 0;
    const int l = left_mi ? left_mi->seg_id_predicted : 0;
    mi->seg_id_predicted = (uint8_t)r.read(seg.pred_probs[a + l]);
    sid = mi->seg_id_predicted ? pred : r.read_tree(kSegmentTree, seg.tree_probs);
  } else {
    sid = r.read_tree(kSegmentTree, seg.tree_probs);
  }
  for (int y = 0; y < y_mis; y++)
    for (int x = 0; x < x_mis; x++) P->seg_map_cur_[off + y * P->mi_cols_ + x] = (uint8_t)sid;
  return sid;
}

// read_skip (vp9_decodemv.c:191-203), vp9_get_skip_context (vp9_pred_common.h:53-59)
int TileCtx::read_skip(int segment_id) {
  if (seg_active(segment_id, SEG_LVL_SKIP)) return 1;
  const int ctx = (above_mi ? above_mi->skip : 0) + (left_mi ? left_mi->skip : 0);
  const int skip = r.read(P->fc_.skip[ctx]);
  if (P->do_counts_) ++O->counts.skip[ctx][skip];
  return skip;
}

// read_tx_size / read_selected_tx_size (vp9_decodemv.c:66-95), get_tx_size_context
// (vp9_pred_common.h:140-155)
int TileCtx::read_tx_size(int allow_select) {
  const int bsize = mi->sb_type;
  const int max_tx = kMaxTxSize[bsize];
  if (allow_select && P->tx_mode_ == TX_MODE_SELECT && bsize >= BLOCK_8X8) {
    int above_ctx = (above_mi && !above_mi->skip) ? above_mi->tx_size : max_tx;
    int left_ctx_ = (left_mi && !left_mi->skip) ? left_mi->tx_size : max_tx;
    if (!left_mi) left_ctx_ = above_ctx;
    if (!above_mi) above_ctx = left_ctx_;
    const int ctx = (above_ctx + left_ctx_) > max_tx;
    const uint8_t *probs = max_tx == TX_8X8 ? P->fc_.tx8[ctx]
                         : max_tx == TX_16X16 ? P->fc_.tx16[ctx] : P->fc_.tx32[ctx];
    int tx = r.read(probs[0]);
    if (tx != TX_4X4 && max_tx >= TX_16X16) {
      tx += r.read(probs[1]);
      if (tx != TX_8X8 && max_tx >= TX_32X32) tx += r.read(probs[2]);
    }
    if (P->do_counts_) {
      if (max_tx == TX_8X8) ++O->counts.tx8[ctx][tx];
      else if (max_tx == TX_16X16) ++O->counts.tx16[ctx][tx];
      else ++O->counts.tx32[ctx][tx];
    }
    return tx;
  }
  return std::min<int>(max_tx, kTxModeToBiggest[P->tx_mode_]);
}

// read_intra_frame_mode_info (vp9_decodemv.c:205-246), vp9_left/above_block_mode (vp9_blockd.c)
void TileCtx::read_intra_frame_mode_info() {
  mi->segment_id = (uint8_t)read_intra_segment_id();
  mi->skip = (uint8_t)read_skip(mi->segment_id);
  mi->tx_size = (uint8_t)read_tx_size(1);
  mi->ref_frame[0] = INTRA_FRAME;
  mi->ref_frame[1] = REF_NONE;
  mi->interp_filter = SWITCHABLE_FILTERS;
  const int bsize = mi->sb_type;
  auto y_mode_of = [&](const BlockInfo *b, int blk) -> int {
    return b->sb_type < BLOCK_8X8 ? b->bmode[blk] : b->mode;
  };
  auto probs_for = [&](int b) -> const uint8_t * {
    int above, left;
    if (b == 0 || b == 1) above = (!above_mi || is_inter(above_mi)) ? DC_PRED : y_mode_of(above_mi, b + 2);
    else above = mi->bmode[b - 2];
    if (b == 0 || b == 2) left = (!left_mi || is_inter(left_mi)) ? DC_PRED : y_mode_of(left_mi, b + 1);
    else left = mi->bmode[b - 1];
    return &kKfYModeProb[(above * 10 + left) * 9];
  };
  switch (bsize) {
    case BLOCK_4X4:
      for (int i = 0; i < 4; ++i) mi->bmode[i] = (uint8_t)r.read_tree(kIntraModeTree, probs_for(i));
      mi->mode = mi->bmode[3];
      break;
    case BLOCK_4X8:
      mi->bmode[0] = mi->bmode[2] = (uint8_t)r.read_tree(kIntraModeTree, probs_for(0));
      mi->bmode[1] = mi->bmode[3] = mi->mode = (uint8_t)r.read_tree(kIntraModeTree, probs_for(1));
      break;
    case BLOCK_8X4:
      mi->bmode[0] = mi->bmode[1] = (uint8_t)r.read_tree(kIntraModeTree, probs_for(0));
      mi->bmode[2] = mi->bmode[3] = mi->mode = (uint8_t)r.read_tree(kIntraModeTree, probs_for(2));
      break;
    default:
      mi->mode = (uint8_t)r.read_tree(kIntraModeTree, probs_for(0));
      mi->bmode[0] = mi->bmode[1] = mi->bmode[2] = mi->bmode[3] = mi->mode;
  }
  mi->uv_mode = (uint8_t)r.read_tree(kIntraModeTree, &kKfUvModeProb[mi->mode * 9]);
}

// read_intra_block_mode_info (vp9_decodemv.c:354-388)
void TileCtx::read_intra_block_mode_info() {
  const int bsize = mi->sb_type;
  auto read_y = [&](int size_group) -> uint8_t {
    const int m = r.read_tree(kIntraModeTree, P->fc_.y_mode[size_group]);
    if (P->do_counts_) ++O->counts.y_mode[size_group][m];
    return (uint8_t)m;
  };
  switch (bsize) {
    case BLOCK_4X4:
      for (int i = 0; i < 4; ++i) mi->bmode[i] = read_y(0);
      mi->mode = mi->bmode[3];
      break;
    case BLOCK_4X8:
      mi->bmode[0] = mi->bmode[2] = read_y(0);
      mi->bmode[1] = mi->bmode[3] = mi->mode = read_y(0);
      break;
    case BLOCK_8X4:
      mi->bmode[0] = mi->bmode[1] = read_y(0);
      mi->bmode[2] = mi->bmode[3] = mi->mode = read_y(0);
      break;
    default:
      mi->mode = read_y(kSizeGroup[bsize]);
      mi->bmode[0] = mi->bmode[1] = mi->bmode[2] = mi->bmode[3] = mi->mode;
  }
  const int uv = r.read_tree(kIntraModeTree, P->fc_.uv_mode[mi->mode]);
  if (P->do_counts_) ++O->counts.uv_mode[mi->mode][uv];
  mi->uv_mode = (uint8_t)uv;
  mi->interp_filter = SWITCHABLE_FILTERS;
  mi->ref_frame[0] = INTRA_FRAME;
  mi->ref_frame[1] = REF_NONE;
}

// read_ref_frames (vp9_decodemv.c:303-341) with the context derivations of
// vp9_pred_common.c:42-316
void TileCtx::read_ref_frames(int8_t ref_frame[2]) {
  FrameContext &fc = P->fc_;
  if (seg_active(mi->segment_id, SEG_LVL_REF_FRAME)) {
    ref_frame[0] = (int8_t)P->seg_.feature_data[mi->segment_id][SEG_LVL_REF_FRAME];
    ref_frame[1] = REF_NONE;
    return;
  }
  const BlockInfo *a = above_mi, *l = left_mi;
  const bool has_above = a != nullptr, has_left = l != nullptr;
  int mode;
  if (P->reference_mode_ == REFERENCE_MODE_SELECT) {
    // vp9_get_reference_mode_context
    int ctx;
    const int fixed = P->comp_fixed_ref_;
    if (has_above && has_left) {
      if (!has_second(a) && !has_second(l)) ctx = (a->ref_frame[0] == fixed) ^ (l->ref_frame[0] == fixed);
      else if (!has_second(a)) ctx = 2 + (a->ref_frame[0] == fixed || !is_inter(a));
      else if (!has_second(l)) ctx = 2 + (l->ref_frame[0] == fixed || !is_inter(l));
      else ctx = 4;
    } else if (has_above || has_left) {
      const BlockInfo *e = has_above ? a : l;
      ctx = !has_second(e) ? (e->ref_frame[0] == fixed) : 3;
    } else {
      ctx = 1;
    }
    mode = r.read(fc.comp_inter[ctx]);
    if (P->do_counts_) ++O->counts.comp_inter[ctx][mode];
  } else {
    mode = P->reference_mode_;
  }
  if (mode == COMPOUND_REFERENCE) {
    const int fix_ref_idx = P->ref_sign_bias_[P->comp_fixed_ref_];
    const int var_ref_idx = !fix_ref_idx;
    const int var1 = P->comp_var_ref_[1], var0 = P->comp_var_ref_[0], fixed = P->comp_fixed_ref_;
    int ctx;  // vp9_get_pred_context_comp_ref_p
    if (has_above && has_left) {
      const bool ai = !is_inter(a), li = !is_inter(l);
      if (ai && li) {
        ctx = 2;
      } else if (ai || li) {
        const BlockInfo *e = ai ? l : a;
        if (!has_second(e)) ctx = 1 + 2 * (e->ref_frame[0] != var1);
        else ctx = 1 + 2 * (e->ref_frame[var_ref_idx] != var1);
      } else {
        const bool l_sg = !has_second(l), a_sg = !has_second(a);
        const int vrfa = a_sg ? a->ref_frame[0] : a->ref_frame[var_ref_idx];
        const int vrfl = l_sg ? l->ref_frame[0] : l->ref_frame[var_ref_idx];
        if (vrfa == vrfl && var1 == vrfa) {
          ctx = 0;
        } else if (l_sg && a_sg) {
          if ((vrfa == fixed && vrfl == var0) || (vrfl == fixed && vrfa == var0)) ctx = 4;
          else if (vrfa == vrfl) ctx = 3;
          else ctx = 1;
        } else if (l_sg || a_sg) {
          const int vrfc = l_sg ? vrfa : vrfl;
          const int rfs = a_sg ? vrfa : vrfl;
          if (vrfc == var1 && rfs != var1) ctx = 1;
          else if (rfs == var1 && vrfc != var1) ctx = 2;
          else ctx = 4;
        } else if (vrfa == vrfl) {
          ctx = 4;
        } else {
          ctx = 2;
        }
      }
    } else if (has_above || has_left) {
      const BlockInfo *e = has_above ? a : l;
      if (!is_inter(e)) ctx = 2;
      else if (has_second(e)) ctx = 4 * (e->ref_frame[var_ref_idx] != var1);
      else ctx = 3 * (e->ref_frame[0] != var1);
    } else {
      ctx = 2;
    }
    const int bit = r.read(fc.comp_ref[ctx]);
    if (P->do_counts_) ++O->counts.comp_ref[ctx][bit];
    ref_frame[fix_ref_idx] = (int8_t)fixed;
    ref_frame[!fix_ref_idx] = (int8_t)P->comp_var_ref_[bit];
  } else {
    int ctx0;  // vp9_get_pred_context_single_ref_p1
    if (has_above && has_left) {
      const bool ai = !is_inter(a), li = !is_inter(l);
      if (ai && li) {
        ctx0 = 2;
      } else if (ai || li) {
        const BlockInfo *e = ai ? l : a;
        if (!has_second(e)) ctx0 = 4 * (e->ref_frame[0] == LAST_FRAME);
        else ctx0 = 1 + (e->ref_frame[0] == LAST_FRAME || e->ref_frame[1] == LAST_FRAME);
      } else {
        const bool a2 = has_second(a), l2 = has_second(l);
        const int a0 = a->ref_frame[0], a1 = a->ref_frame[1], l0 = l->ref_frame[0], l1 = l->ref_frame[1];
        if (a2 && l2) {
          ctx0 = 1 + (a0 == LAST_FRAME || a1 == LAST_FRAME || l0 == LAST_FRAME || l1 == LAST_FRAME);
        } else if (a2 || l2) {
          const int rfs = !a2 ? a0 : l0;
          const int crf1 = a2 ? a0 : l0, crf2 = a2 ? a1 : l1;
          if (rfs == LAST_FRAME) ctx0 = 3 + (crf1 == LAST_FRAME || crf2 == LAST_FRAME);
          else ctx0 = (crf1 == LAST_FRAME || crf2 == LAST_FRAME);
        } else {
          ctx0 = 2 * (a0 == LAST_FRAME) + 2 * (l0 == LAST_FRAME);
        }
      }
    } else if (has_above || has_left) {
      const BlockInfo *e = has_above ? a : l;
      if (!is_inter(e)) ctx0 = 2;
      else if (!has_second(e)) ctx0 = 4 * (e->ref_frame[0] == LAST_FRAME);
      else ctx0 = 1 + (e->ref_frame[0] == LAST_FRAME || e->ref_frame[1] == LAST_FRAME);
    } else {
      ctx0 = 2;
    }
    const int bit0 = r.read(fc.single_ref[ctx0][0]);
    if (P->do_counts_) ++O->counts.single_ref[ctx0][0][bit0];
    if (bit0) {
      int ctx1;  // vp9_get_pred_context_single_ref_p2
      if (has_above && has_left) {
        const bool ai = !is_inter(a), li = !is_inter(l);
        if (ai && li) {
          ctx1 = 2;
        } else if (ai || li) {
          const BlockInfo *e = ai ? l : a;
          if (!has_second(e)) {
            if (e->ref_frame[0] == LAST_FRAME) ctx1 = 3;
            else ctx1 = 4 * (e->ref_frame[0] == GOLDEN_FRAME);
          } else {
            ctx1 = 1 + 2 * (e->ref_frame[0] == GOLDEN_FRAME || e->ref_frame[1] == GOLDEN_FRAME);
          }
        } else {
          const bool a2 = has_second(a), l2 = has_second(l);
          const int a0 = a->ref_frame[0], a1 = a->ref_frame[1], l0 = l->ref_frame[0], l1 = l->ref_frame[1];
          if (a2 && l2) {
            if (a0 == l0 && a1 == l1)
              ctx1 = 3 * (a0 == GOLDEN_FRAME || a1 == GOLDEN_FRAME || l0 == GOLDEN_FRAME || l1 == GOLDEN_FRAME);
            else
              ctx1 = 2;
          } else if (a2 || l2) {
            const int rfs = !a2 ? a0 : l0;
            const int crf1 = a2 ? a0 : l0, crf2 = a2 ? a1 : l1;
            if (rfs == GOLDEN_FRAME) ctx1 = 3 + (crf1 == GOLDEN_FRAME || crf2 == GOLDEN_FRAME);
            else if (rfs == ALTREF_FRAME) ctx1 = crf1 == GOLDEN_FRAME || crf2 == GOLDEN_FRAME;
            else ctx1 = 1 + 2 * (crf1 == GOLDEN_FRAME || crf2 == GOLDEN_FRAME);
          } else {
            if (a0 == LAST_FRAME && l0 == LAST_FRAME) {
              ctx1 = 3;
            } else if (a0 == LAST_FRAME || l0 == LAST_FRAME) {
              const int e0 = (a0 == LAST_FRAME) ? l0 : a0;
              ctx1 = 4 * (e0 == GOLDEN_FRAME);
            } else {
              ctx1 = 2 * (a0 == GOLDEN_FRAME) + 2 * (l0 == GOLDEN_FRAME);
            }
          }
        }
      } else if (has_above || has_left) {
        const BlockInfo *e = has_above ? a : l;
        if (!is_inter(e) || (e->ref_frame[0] == LAST_FRAME && !has_second(e))) ctx1 = 2;
        else if (!has_second(e)) ctx1 = 4 * (e->ref_frame[0] == GOLDEN_FRAME);
        else ctx1 = 3 * (e->ref_frame[0] == GOLDEN_FRAME || e->ref_frame[1] == GOLDEN_FRAME);
      } else {
        ctx1 = 2;
      }
      const int bit1 = r.read(fc.single_ref[ctx1][1]);
      if (P->do_counts_) ++O->counts.single_ref[ctx1][1][bit1];
      ref_frame[0] = bit1 ? ALTREF_FRAME : GOLDEN_FRAME;
    } else {
      ref_frame[0] = LAST_FRAME;
    }
    ref_frame[1] = REF_NONE;
  }
}

// read_mv_component / read_mv (vp9_decodemv.c:248-292), vp9_inc_mv (vp9_entropymv.c:112-155)
int TileCtx::read_mv_component(const MvCompProbs &c, int usehp) {
  const int sign = r.read(c.sign);
  const int mv_class = r.read_tree(kMvClassTree, c.classes);
  const bool class0 = mv_class == 0;
  int d, mag;
  if (class0) {
    d = r.read(c.class0[0]);
    mag = 0;
  } else {
    const int n = mv_class + 1 - 1;  // mv_class + CLASS0_BITS - 1
    d = 0;
    for (int i = 0; i < n; ++i) d |= r.read(c.bits[i]) << i;
    mag = 2 << (mv_class + 2);
  }
  const int fr = r.read_tree(kMvFpTree, class0 ? c.class0_fp[d] : c.fp);
  const int hp = usehp ? r.read(class0 ? c.class0_hp : c.hp) : 1;
  mag += ((d << 3) | (fr << 1) | hp) + 1;
  return sign ? -mag : mag;
}

static inline void inc_mv_component(int v, MvCompCounts &cc) {
  const int s = v < 0;
  cc.sign[s]++;
  const int z = (s ? -v : v) - 1;
  int c;
  if (z >= 2 * 4096) {
    c = 10;
  } else {
    int t = z >> 3;
    c = 0;
    while (t > 1) { t >>= 1; c++; }
  }
  const int base = c ? 2 << (c + 2) : 0;
  const int o = z - base;
  cc.classes[c]++;
  const int d = o >> 3, f = (o >> 1) & 3, e = o & 1;
  if (c == 0) {
    cc.class0[d]++;
    cc.class0_fp[d][f]++;
    cc.class0_hp[e]++;
  } else {
    const int b = c + 1 - 1;
    for (int i = 0; i < b; ++i) cc.bits[i][(d >> i) & 1]++;
    cc.fp[f]++;
    cc.hp[e]++;
  }
}

void TileCtx::read_mv(Mv *mv, const Mv *ref) {
  const int joint = r.read_tree(kMvJointTree, P->fc_.mv_joints);
  const int use_hp = P->allow_hp_ && (abs(ref->row) < 64 && abs(ref->col) < 64);
  int drow = 0, dcol = 0;
  if (joint == MV_JOINT_HZVNZ || joint == MV_JOINT_HNZVNZ) drow = read_mv_component(P->fc_.mvc[0], use_hp);
  if (joint == MV_JOINT_HNZVZ || joint == MV_JOINT_HNZVNZ) dcol = read_mv_component(P->fc_.mvc[1], use_hp);
  if (P->do_counts_) {
    const int j = (drow == 0) ? (dcol == 0 ? MV_JOINT_ZERO : MV_JOINT_HNZVZ)
                              : (dcol == 0 ? MV_JOINT_HZVNZ : MV_JOINT_HNZVNZ);
    ++O->counts.mv_joints[j];
    if (drow) inc_mv_component(drow, O->counts.mvc[0]);
    if (dcol) inc_mv_component(dcol, O->counts.mvc[1]);
  }
  mv->row = (int16_t)(ref->row + drow);
  mv->col = (int16_t)(ref->col + dcol);
}

// assign_mv (vp9_decodemv.c:402-436)
bool TileCtx::assign_mv(int mode, Mv mv[2], const Mv ref_mv[2], const Mv near_nearest[2],
                        int is_compound) {
  bool ret = true;
  switch (mode) {
    case NEWMV:
      for (int i = 0; i < 1 + is_compound; ++i) {
        read_mv(&mv[i], &ref_mv[i]);
        ret = ret && (mv[i].row > -(1 << 14) && mv[i].row < (1 << 14) - 1 &&
                      mv[i].col > -(1 << 14) && mv[i].col < (1 << 14) - 1);
      }
      break;
    case NEARMV:
    case NEARESTMV:
      mv[0] = near_nearest[0];
      mv[1] = near_nearest[1];
      break;
    case ZEROMV:
      mv[0].row = mv[0].col = mv[1].row = mv[1].col = 0;
      break;
    default: return false;
  }
  return ret;
}

// dec_find_mv_refs (vp9_decodemv.c:482-610) with helpers of vp9_mvref_common.h:213-283
int TileCtx::find_mv_refs(int mode, int ref_frame, const int8_t *search, Mv *list, int block) {
  const int *sign_bias = P->ref_sign_bias_;
  int count = 0;
  bool different_ref_found = false;
  const MvRef *prev = P->use_prev_frame_mvs_ ? &P->prev_mvs_[(size_t)mi_row * P->mi_cols_ + mi_col] : nullptr;
  const bool early_break = (mode != NEARMV);
  list[0].row = list[0].col = list[1].row = list[1].col = 0;
  bool done = false;
  auto add = [&](Mv mv) {  // ADD_MV_REF_LIST_EB
    if (count) {
      if (mv != list[0]) {
        list[count++] = mv;
        done = true;
      }
    } else {
      list[count++] = mv;
      if (early_break) done = true;
    }
  };
  auto inside = [&](int drow, int dcol) {
    return !(mi_row + drow < 0 || mi_col + dcol < mi_col_start || mi_row + drow >= P->mi_rows_ ||
             mi_col + dcol >= mi_col_end);
  };
  int i = 0;
  if (block >= 0) {
    for (i = 0; i < 2 && !done; ++i) {
      const int drow = search[i * 2], dcol = search[i * 2 + 1];
      if (inside(drow, dcol)) {
        const BlockInfo *c = at(mi_row + drow, mi_col + dcol);
        different_ref_found = true;
        auto sub_mv = [&](int which) -> Mv {  // get_sub_block_mv
          if (c->sb_type < BLOCK_8X8) return c->bmv[kIdxNColumnToSubblock[block * 2 + (dcol == 0)]][which];
          return c->mv[which];
        };
        if (c->ref_frame[0] == ref_frame) add(sub_mv(0));
        else if (c->ref_frame[1] == ref_frame) add(sub_mv(1));
      }
    }
  }
  for (; i < 8 && !done; ++i) {
    const int drow = search[i * 2], dcol = search[i * 2 + 1];
    if (inside(drow, dcol)) {
      const BlockInfo *c = at(mi_row + drow, mi_col + dcol);
      different_ref_found = true;
      if (c->ref_frame[0] == ref_frame) add(c->mv[0]);
      else if (c->ref_frame[1] == ref_frame) add(c->mv[1]);
    }
  }
  if (!done && prev) {
    if (prev->ref_frame[0] == ref_frame) add(prev->mv[0]);
    else if (prev->ref_frame[1] == ref_frame) add(prev->mv[1]);
  }
  if (!done && different_ref_found) {
    for (i = 0; i < 8 && !done; ++i) {
      const int drow = search[i * 2], dcol = search[i * 2 + 1];
      if (inside(drow, dcol)) {
        const BlockInfo *c = at(mi_row + drow, mi_col + dcol);
        if (is_inter(c)) {  // IF_DIFF_REF_FRAME_ADD_MV_EB
          auto scaled = [&](int which) -> Mv {
            Mv mv = c->mv[which];
            if (sign_bias[c->ref_frame[which]] != sign_bias[ref_frame]) {
              mv.row = (int16_t)-mv.row;
              mv.col = (int16_t)-mv.col;
            }
            return mv;
          };
          if (c->ref_frame[0] != ref_frame) add(scaled(0));
          if (!done && has_second(c) && c->ref_frame[1] != ref_frame && c->mv[1] != c->mv[0])
            add(scaled(1));
        }
      }
    }
  }
  if (!done && prev) {
    if (prev->ref_frame[0] != ref_frame && prev->ref_frame[0] > INTRA_FRAME) {
      Mv mv = prev->mv[0];
      if (sign_bias[prev->ref_frame[0]] != sign_bias[ref_frame]) {
        mv.row = (int16_t)-mv.row;
        mv.col = (int16_t)-mv.col;
      }
      add(mv);
    }
    if (!done && prev->ref_frame[1] > INTRA_FRAME && prev->ref_frame[1] != ref_frame &&
        prev->mv[1] != prev->mv[0]) {
      Mv mv = prev->mv[1];
      if (sign_bias[prev->ref_frame[1]] != sign_bias[ref_frame]) {
        mv.row = (int16_t)-mv.row;
        mv.col = (int16_t)-mv.col;
      }
      add(mv);
    }
  }
  if (!done) count = (mode == NEARMV) ? 2 : 1;
  // clamp_mv_ref: MV_BORDER = 16 pels in 1/8
  for (i = 0; i < count; ++i) {
    list[i].col = (int16_t)clampi(list[i].col, mb_to_left_edge - 128, mb_to_right_edge + 128);
    list[i].row = (int16_t)clampi(list[i].row, mb_to_top_edge - 128, mb_to_bottom_edge + 128);
  }
  return count;
}

// append_sub8x8_mvs_for_idx (vp9_decodemv.c:612-674)
void TileCtx::append_sub8x8_mvs_for_idx(const int8_t *search, int b_mode, int block, int ref, Mv *best) {
  Mv mv_list[2];
  const Mv zero = { 0, 0 };
  switch (block) {
    case 0: {
      const int n = find_mv_refs(b_mode, mi->ref_frame[ref], search, mv_list, block);
      *best = mv_list[n - 1];
      break;
    }
    case 1:
    case 2:
      if (b_mode == NEARESTMV) {
        *best = mi->bmv[0][ref];
      } else {
        find_mv_refs(b_mode, mi->ref_frame[ref], search, mv_list, block);
        *best = zero;
        for (int n = 0; n < 2; ++n)
          if (mi->bmv[0][ref] != mv_list[n]) {
            *best = mv_list[n];
            break;
          }
      }
      break;
    case 3:
      if (b_mode == NEARESTMV) {
        *best = mi->bmv[2][ref];
      } else {
        *best = zero;
        if (mi->bmv[2][ref] != mi->bmv[1][ref]) {
          *best = mi->bmv[1][ref];
          break;
        }
        if (mi->bmv[2][ref] != mi->bmv[0][ref]) {
          *best = mi->bmv[0][ref];
          break;
        }
        find_mv_refs(b_mode, mi->ref_frame[ref], search, mv_list, block);
        for (int n = 0; n < 2; ++n)
          if (mi->bmv[2][ref] != mv_list[n]) {
            *best = mv_list[n];
            break;
          }
      }
      break;
  }
}

// get_mode_context (vp9_decodemv.c:676-697)
int TileCtx::get_mode_context(const int8_t *search) {
  int counter = 0;
  for (int i = 0; i < 2; ++i) {
    const int drow = search[i * 2], dcol = search[i * 2 + 1];
    if (!(mi_row + drow < 0 || mi_col + dcol < mi_col_start || mi_row + drow >= P->mi_rows_ ||
          mi_col + dcol >= mi_col_end))
      counter += kMode2Counter[at(mi_row + drow, mi_col + dcol)->mode];
  }
  return kCounterToContext[counter];
}

static inline void lower_mv_precision(Mv *mv, int allow_hp) {
  const int use_hp = allow_hp && (abs(mv->row) < 64 && abs(mv->col) < 64);
  if (!use_hp) {
    if (mv->row & 1) mv->row += (mv->row > 0 ? -1 : 1);
    if (mv->col & 1) mv->col += (mv->col > 0 ? -1 : 1);
  }
}

// read_inter_block_mode_info (vp9_decodemv.c:699-798)
void TileCtx::read_inter_block_mode_info() {
  const int bsize = mi->sb_type;
  const int allow_hp = P->allow_hp_;
  Mv best_ref_mvs[2] = { { 0, 0 }, { 0, 0 } };
  const int8_t *search = &kMvRefBlocks[bsize * 16];
  read_ref_frames(mi->ref_frame);
  const int is_compound = has_second(mi);
  const int inter_mode_ctx = get_mode_context(search);
  auto read_inter_mode = [&]() -> int {
    const int m = r.read_tree(kInterModeTree, P->fc_.inter_mode[inter_mode_ctx]);
    if (P->do_counts_) ++O->counts.inter_mode[inter_mode_ctx][m];
    return NEARESTMV + m;
  };
  if (seg_active(mi->segment_id, SEG_LVL_SKIP)) {
    mi->mode = ZEROMV;
    if (bsize < BLOCK_8X8) {
      corrupted = true;
      return;
    }
  } else {
    if (bsize >= BLOCK_8X8) mi->mode = (uint8_t)read_inter_mode();
  }
  if (P->interp_filter_ == SWITCHABLE) {
    // get_pred_context_switchable_interp (vp9_pred_common.h:67-87)
    const int left_type = left_mi ? left_mi->interp_filter : SWITCHABLE_FILTERS;
    const int above_type = above_mi ? above_mi->interp_filter : SWITCHABLE_FILTERS;
    int ctx;
    if (left_type == above_type) ctx = left_type;
    else if (left_type == SWITCHABLE_FILTERS) ctx = above_type;
    else if (above_type == SWITCHABLE_FILTERS) ctx = left_type;
    else ctx = SWITCHABLE_FILTERS;
    const int t = r.read_tree(kSwitchableInterpTree, P->fc_.switchable_interp[ctx]);
    if (P->do_counts_) ++O->counts.switchable_interp[ctx][t];
    mi->interp_filter = (uint8_t)t;
  } else {
    mi->interp_filter = (uint8_t)P->interp_filter_;
  }
  if (bsize < BLOCK_8X8) {
    const int num_4x4_w = 1 << bmode_blocks_wl;
    const int num_4x4_h = 1 << bmode_blocks_hl;
    int b_mode = NEARESTMV;
    bool got_mv_refs_for_new = false;
    Mv best_sub8x8[2] = { { 0, 0 }, { 0, 0 } };
    for (int idy = 0; idy < 2; idy += num_4x4_h) {
      for (int idx = 0; idx < 2; idx += num_4x4_w) {
        const int j = idy * 2 + idx;
        b_mode = read_inter_mode();
        if (b_mode == NEARESTMV || b_mode == NEARMV) {
          for (int ref = 0; ref < 1 + is_compound; ++ref)
            append_sub8x8_mvs_for_idx(search, b_mode, j, ref, &best_sub8x8[ref]);
        } else if (b_mode == NEWMV && !got_mv_refs_for_new) {
          for (int ref = 0; ref < 1 + is_compound; ++ref) {
            Mv tmp[2];
            find_mv_refs(NEWMV, mi->ref_frame[ref], search, tmp, -1);
            lower_mv_precision(&tmp[0], allow_hp);
            best_ref_mvs[ref] = tmp[0];
            got_mv_refs_for_new = true;
          }
        }
        if (!assign_mv(b_mode, mi->bmv[j], best_ref_mvs, best_sub8x8, is_compound)) {
          corrupted = true;
          break;
        }
        if (num_4x4_h == 2) { mi->bmv[j + 2][0] = mi->bmv[j][0]; mi->bmv[j + 2][1] = mi->bmv[j][1]; }
        if (num_4x4_w == 2) { mi->bmv[j + 1][0] = mi->bmv[j][0]; mi->bmv[j + 1][1] = mi->bmv[j][1]; }
      }
    }
    mi->mode = (uint8_t)b_mode;
    mi->mv[0] = mi->bmv[3][0];
    mi->mv[1] = mi->bmv[3][1];
  } else {
    if (mi->mode != ZEROMV) {
      for (int ref = 0; ref < 1 + is_compound; ++ref) {
        Mv tmp[2];
        const int n = find_mv_refs(mi->mode, mi->ref_frame[ref], search, tmp, -1);
        lower_mv_precision(&tmp[n - 1], allow_hp);
        best_ref_mvs[ref] = tmp[n - 1];
      }
    }
    if (!assign_mv(mi->mode, mi->mv, best_ref_mvs, best_ref_mvs, is_compound)) corrupted = true;
    for (int j = 0; j < 4; j++) { mi->bmv[j][0] = mi->mv[0]; mi->bmv[j][1] = mi->mv[1]; }
  }
}

// read_inter_frame_mode_info (vp9_decodemv.c:800-818)
void TileCtx::read_inter_frame_mode_info() {
  mi->segment_id = (uint8_t)read_inter_segment_id();
  mi->skip = (uint8_t)read_skip(mi->segment_id);
  int inter_block;
  if (seg_active(mi->segment_id, SEG_LVL_REF_FRAME)) {
    inter_block = P->seg_.feature_data[mi->segment_id][SEG_LVL_REF_FRAME] != INTRA_FRAME;
  } else {
    // get_intra_inter_context (vp9_pred_common.h:96-110)
    int ctx;
    if (above_mi && left_mi) {
      const bool ai = !is_inter(above_mi), li = !is_inter(left_mi);
      ctx = (li && ai) ? 3 : (li || ai);
    } else if (above_mi || left_mi) {
      ctx = 2 * !is_inter(above_mi ? above_mi : left_mi);
    } else {
      ctx = 0;
    }
    inter_block = r.read(P->fc_.intra_inter[ctx]);
    if (P->do_counts_) ++O->counts.intra_inter[ctx][inter_block];
  }
  mi->tx_size = (uint8_t)read_tx_size(!mi->skip || !inter_block);
  if (inter_block) read_inter_block_mode_info();
  else read_intra_block_mode_info();
}

// decode_coefs + vp9_decode_block_tokens (vp9_detokenize.c:115-333).  Coefficients are written
// to the packed stream in scan order (positions 0..eob-1) already dequantised.
int TileCtx::decode_block_tokens(int plane, int x, int y, int tx_size, int tx_type, uint32_t *coef_off) {
  // four copies of the token loop: sample width and "collect counts" are frame constants
  const bool hbd = P->bit_depth_ > 8;
  if (P->do_counts_)
    return hbd ? decode_block_tokens_t<true, true>(plane, x, y, tx_size, tx_type, coef_off)
               : decode_block_tokens_t<false, true>(plane, x, y, tx_size, tx_type, coef_off);
  return hbd ? decode_block_tokens_t<true, false>(plane, x, y, tx_size, tx_type, coef_off)
             : decode_block_tokens_t<false, false>(plane, x, y, tx_size, tx_type, coef_off);
}
template <bool HBD, bool COUNTS>
int TileCtx::decode_block_tokens_t(int plane, int x, int y, int tx_size, int tx_type, uint32_t *coef_off) {
  const int ss_x = plane ? P->ss_x_ : 0, ss_y = plane ? P->ss_y_ : 0;
  uint8_t *a = &P->above_ctx_[plane][((mi_col * 2) >> ss_x) + x];
  uint8_t *l = &left_ctx[plane][(((mi_row * 2) & 15) >> ss_y) + y];
  const int n = 1 << tx_size;
  // get_ctx_shift: entries past the frame edge stay zero
  int na = n, nl = n;
  if (max_blocks_wide && n + x > max_blocks_wide) na = max_blocks_wide - x;
  if (max_blocks_high && n + y > max_blocks_high) nl = max_blocks_high - y;
  int ctx = 0;
  {
    int sa = 0, sl = 0;
    for (int i = 0; i < n; i++) { sa |= a[i]; sl |= l[i]; }
    ctx = (sa != 0) + (sl != 0);
  }
  const int16_t *scan, *nb;
  switch (tx_size) {
    case TX_4X4:
      scan = tx_type == 0 ? kDefaultScan_4x4 : (tx_type == 1 ? kRowScan_4x4 : (tx_type == 2 ? kColScan_4x4 : kDefaultScan_4x4));
      nb = tx_type == 0 ? kDefaultScan_4x4_nb : (tx_type == 1 ? kRowScan_4x4_nb : (tx_type == 2 ? kColScan_4x4_nb : kDefaultScan_4x4_nb));
      break;
    case TX_8X8:
      scan = tx_type == 0 ? kDefaultScan_8x8 : (tx_type == 1 ? kRowScan_8x8 : (tx_type == 2 ? kColScan_8x8 : kDefaultScan_8x8));
      nb = tx_type == 0 ? kDefaultScan_8x8_nb : (tx_type == 1 ? kRowScan_8x8_nb : (tx_type == 2 ? kColScan_8x8_nb : kDefaultScan_8x8_nb));
      break;
    case TX_16X16:
      scan = tx_type == 0 ? kDefaultScan_16x16 : (tx_type == 1 ? kRowScan_16x16 : (tx_type == 2 ? kColScan_16x16 : kDefaultScan_16x16));
      nb = tx_type == 0 ? kDefaultScan_16x16_nb : (tx_type == 1 ? kRowScan_16x16_nb : (tx_type == 2 ? kColScan_16x16_nb : kDefaultScan_16x16_nb));
      break;
    default:
      scan = kDefaultScan_32x32;
      nb = kDefaultScan_32x32_nb;
  }
  const int type = plane ? 1 : 0;
  const int ref = is_inter(mi);
  const int max_eob = 16 << (tx_size << 1);
  const uint8_t(*coef_probs)[6][3] = P->fc_.coef[tx_size][type][ref];
  uint32_t(*coef_counts)[6][4] = O->counts.coef[tx_size][type][ref];
  uint32_t(*eob_branch)[6] = O->counts.eob_branch[tx_size][type][ref];
  constexpr bool counts = COUNTS;
  const uint8_t *band_translate = tx_size == TX_4X4 ? kCoefBand4x4 : kCoefBand8x8Plus;
  const int dq_shift = (tx_size == TX_32X32);
  const int16_t *dq = plane ? P->uv_dequant_[mi->segment_id] : P->y_dequant_[mi->segment_id];
  const int bd = P->bit_depth_;
  const uint8_t *cat6_prob = bd == 12 ? kCat6ProbHigh12 : (bd == 10 ? kCat6ProbHigh12 + 2 : kCat6Prob);
  const int cat6_bits = bd == 12 ? 18 : (bd == 10 ? 16 : 14);
  constexpr bool hbd = HBD;
  // coefficient stream: grow geometrically, write through raw pointers (no per-block zero fill)
  size_t base;
  if (hbd) {
    base = O->n_coef;
    if (O->coef32.size() < base + max_eob) O->coef32.resize(std::max(O->coef32.size() * 2, base + 65536));
  } else {
    base = O->n_coef;
    if (O->coef16.size() < base + max_eob) O->coef16.resize(std::max(O->coef16.size() * 2, base + 65536));
  }
  int16_t *o16 = hbd ? nullptr : &O->coef16[base];
  int32_t *o32 = hbd ? &O->coef32[base] : nullptr;
  BoolReader r = this->r;  // local copy: value/range/count live in registers for the whole block
  int c = 0;
  int dqv = dq[0];
  int band;
  const uint8_t *prob;
  auto read_coeff = [&](const uint8_t *probs, int nbits) {
    int val = 0;
    for (int i = 0; i < nbits; ++i) val = (val << 1) | r.read_nb(probs[i]);
    return val;
  };
  while (c < max_eob) {
    int val = -1;
    band = band_translate[c];
    prob = coef_probs[band][ctx];
    if (counts) ++eob_branch[band][ctx];
    if (!r.read(prob[0])) {
      if (counts) ++coef_counts[band][ctx][3];  // EOB_MODEL_TOKEN
      break;
    }
    bool ended = false;
    while (!r.read(prob[1])) {
      if (counts) ++coef_counts[band][ctx][0];  // ZERO_TOKEN
      dqv = dq[1];
      token_cache[scan[c]] = 0;
      if (hbd) o32[c] = 0; else o16[c] = 0;
      ++c;
      if (c >= max_eob) { ended = true; break; }
      ctx = (1 + token_cache[nb[2 * c]] + token_cache[nb[2 * c + 1]]) >> 1;
      band = band_translate[c];
      prob = coef_probs[band][ctx];
    }
    if (ended) break;
    int v;
    if (r.read(prob[2])) {
      const uint8_t *p = &kPareto8Full[(prob[2] - 1) * 8];
      if (counts) ++coef_counts[band][ctx][2];  // TWO_TOKEN
      if (r.read(p[0])) {
        if (r.read(p[3])) {
          token_cache[scan[c]] = 5;
          if (r.read(p[5])) {
            if (r.read(p[7])) val = 67 + read_coeff(cat6_prob, cat6_bits);
            else val = 35 + read_coeff(kCat5Prob, 5);
          } else if (r.read(p[6])) {
            val = 19 + read_coeff(kCat4Prob, 4);
          } else {
            val = 11 + read_coeff(kCat3Prob, 3);
          }
        } else {
          token_cache[scan[c]] = 4;
          if (r.read(p[4])) val = 7 + read_coeff(kCat2Prob, 2);
          else val = 5 + read_coeff(kCat1Prob, 1);
        }
        v = (int)(((int64_t)val * dqv) >> dq_shift);
      } else {
        if (r.read(p[1])) {
          token_cache[scan[c]] = 3;
          v = ((3 + r.read(p[2])) * dqv) >> dq_shift;
        } else {
          token_cache[scan[c]] = 2;
          v = (2 * dqv) >> dq_shift;
        }
      }
    } else {
      if (counts) ++coef_counts[band][ctx][1];  // ONE_TOKEN
      token_cache[scan[c]] = 1;
      v = dqv >> dq_shift;
    }
    {
      const int sgn = -r.read_nb(128);
      v = (v ^ sgn) - sgn;
    }
    if (hbd) o32[c] = v; else o16[c] = (int16_t)v;
    ++c;
    ctx = (1 + token_cache[nb[2 * c]] + token_cache[nb[2 * c + 1]]) >> 1;
    dqv = dq[1];
  }
  const int eob = c;
  this->r = r;
  O->n_coef = base + eob;
  *coef_off = (uint32_t)base;
  const uint8_t flag = eob > 0;
  for (int i = 0; i < n; i++) {
    a[i] = i < na ? flag : 0;
    l[i] = i < nl ? flag : 0;
  }
  return eob;
}

// Residual parse + record emission.  Block walk: predict_recon_intra / predict_recon_inter
// (vp9_decodeframe.c:859-917) and upstream decode_block (skip fix-up, vp9_decodeframe.c:1195).
void TileCtx::residual_and_record() {
  const int bsize = mi->sb_type;
  const bool inter = is_inter(mi);
  const bool less8x8 = bsize < BLOCK_8X8;
  int eobtotal = 0;
  for (int plane = 0; plane < 3; ++plane) {
    const int ss_x = plane ? P->ss_x_ : 0, ss_y = plane ? P->ss_y_ : 0;
    const int tx_size = plane ? kUvTxSize[((bsize * 4 + mi->tx_size) * 2 + ss_x) * 2 + ss_y] : mi->tx_size;
    const int n4_w = (bw << 1) >> ss_x, n4_h = (bh << 1) >> ss_y;
    const int step = 1 << tx_size;
    const int mbw = n4_w + (mb_to_right_edge >= 0 ? 0 : mb_to_right_edge >> (5 + ss_x));
    const int mbh = n4_h + (mb_to_bottom_edge >= 0 ? 0 : mb_to_bottom_edge >> (5 + ss_y));
    max_blocks_wide = mb_to_right_edge >= 0 ? 0 : mbw;
    max_blocks_high = mb_to_bottom_edge >= 0 ? 0 : mbh;
    const int px4 = (mi_col * 2) >> ss_x, py4 = (mi_row * 2) >> ss_y;
    const int n4_wl = bwl - ss_x;  // pd->n4_wl
    for (int row = 0; row < mbh; row += step)
      for (int col = 0; col < mbw; col += step) {
        vp9b_txb t;
        t.x4 = (uint16_t)(px4 + col);
        t.y4 = (uint16_t)(py4 + row);
        t.coef_off = 0;
        t.eob = 0;
        int tx_type = 0;
        if (!inter) {
          int mode = plane == 0 ? mi->mode : mi->uv_mode;
          if (less8x8 && plane == 0) mode = mi->bmode[(row << 1) + col];
          if (plane == 0 && !P->lossless_) tx_type = kIntraModeToTxType[mode];
          // vp9_predict_intra_block (vp9_reconintra.c:404-424)
          const int have_top = row || (above_mi != nullptr);
          const int have_left = col || (left_mi != nullptr);
          const int have_right = (col + step) < (1 << n4_wl);
          t.aux = (uint8_t)(mode | (have_top << 4) | (have_left << 5) | (have_right << 6));
        } else {
          t.aux = 0;
        }
        // 32x32 only has DCT; 4x4 lossless is WHT (flagged in the frame header)
        if (tx_size == TX_32X32) tx_type = 0;
        t.tx = (uint8_t)(tx_size | (tx_type << 2) | (plane << 4));
        if (!mi->skip) {
          uint32_t off = 0;
          const int eob = decode_block_tokens(plane, col, row, tx_size, tx_type, &off);
          t.eob = (uint16_t)eob;
          t.coef_off = off;
          eobtotal += eob;
        }
        if (!inter) O->txb_intra[plane].push_back(t);
        else if (t.eob > 0) O->txb_inter[tx_size].push_back(t);
      }
  }
  if (inter && !mi->skip && !less8x8 && eobtotal == 0) mi->skip = 1;  // skip loopfilter
}

// vp9_build_mask (vp9_loopfilter.c:1528-1608)
void TileCtx::build_lf_mask() {
  const int bsize = mi->sb_type;
  const int tx_y = mi->tx_size;
  const int level = P->lf_lvl_[mi->segment_id][mi->ref_frame[0] < 0 ? 0 : mi->ref_frame[0]][kModeLfLut[mi->mode]];
  if (!level) return;
  const int tx_uv = kUvTxSize[((bsize * 4 + tx_y) * 2 + 1) * 2 + 1];
  vp9b_lfm &lfm = P->out_lfm_[(mi_row >> 3) * P->sb_cols_ + (mi_col >> 3)];
  const int row_in_sb = mi_row & 7, col_in_sb = mi_col & 7;
  const int shift_y = col_in_sb + (row_in_sb << 3);
  const int shift_uv = (col_in_sb >> 1) + ((row_in_sb >> 1) << 2);
  // the 4:2:0 chroma masks; 4:4:4 filters chroma with the luma masks, 4:2:2 / 4:4:0 get their own
  // masks from the finished mode info (Parser::build_lfm_non420)
  const int build_uv = (P->ss_x_ == P->ss_y_) && kFirstBlockIn16x16[row_in_sb * 8 + col_in_sb];
  {
    int index = shift_y;
    for (int i = 0; i < bh; i++) {
      memset(&lfm.lfl_y[index], level, bw);
      index += 8;
    }
  }
  lfm.above_y[tx_y] |= kLfAbovePredMask[bsize] << shift_y;
  lfm.left_y[tx_y] |= kLfLeftPredMask[bsize] << shift_y;
  if (build_uv) {
    lfm.above_uv[tx_uv] |= (uint16_t)(kLfAbovePredMaskUv[bsize] << shift_uv);
    lfm.left_uv[tx_uv] |= (uint16_t)(kLfLeftPredMaskUv[bsize] << shift_uv);
  }
  if (mi->skip && is_inter(mi)) return;
  lfm.above_y[tx_y] |= (kLfSizeMask[bsize] & kLfAboveTxMask[tx_y]) << shift_y;
  lfm.left_y[tx_y] |= (kLfSizeMask[bsize] & kLfLeftTxMask[tx_y]) << shift_y;
  if (build_uv) {
    lfm.above_uv[tx_uv] |= (uint16_t)((kLfSizeMaskUv[bsize] & kLfAboveTxMaskUv[tx_uv]) << shift_uv);
    lfm.left_uv[tx_uv] |= (uint16_t)((kLfSizeMaskUv[bsize] & kLfLeftTxMaskUv[tx_uv]) << shift_uv);
  }
  if (tx_y == TX_4X4) lfm.int_4x4_y |= kLfSizeMask[bsize] << shift_y;
  if (build_uv && tx_uv == TX_4X4) lfm.int_4x4_uv |= (uint16_t)((kLfSizeMaskUv[bsize] & 0xffff) << shift_uv);
}

// decode_block (vp9_decodeframe.c:1198-1242, upstream order)
void TileCtx::decode_block(int row, int col, int bsize, int bwl_, int bhl_) {
  mi_row = row;
  mi_col = col;
  bwl = bwl_;
  bhl = bhl_;
  bw = 1 << (bwl - 1);
  bh = 1 << (bhl - 1);
  x_mis = std::min(bw, P->mi_cols_ - col);
  y_mis = std::min(bh, P->mi_rows_ - row);
  // set_mi_row_col (vp9_onyxc_int.h:423-435)
  mb_to_top_edge = -((row * 8) * 8);
  mb_to_bottom_edge = ((P->mi_rows_ - bh - row) * 8) * 8;
  mb_to_left_edge = -((col * 8) * 8);
  mb_to_right_edge = ((P->mi_cols_ - bw - col) * 8) * 8;
  above_mi = row != 0 ? at(row - 1, col) : nullptr;
  left_mi = col > mi_col_start ? at(row, col - 1) : nullptr;
  // the worker's own range of the block array: at most one block per 8x8 cell of its tile columns
  const uint32_t bidx = O->block_next++;
  P->blocks_[bidx] = BlockInfo();
  // NOTE: emplace_back may reallocate; refresh neighbour pointers from indices afterwards
  above_mi = row != 0 ? at(row - 1, col) : nullptr;
  left_mi = col > mi_col_start ? at(row, col - 1) : nullptr;
  mi = &P->blocks_[bidx];
  memset(mi, 0, sizeof(*mi));
  mi->sb_type = (uint8_t)bsize;
  if (bsize >= BLOCK_8X8 && (P->ss_x_ || P->ss_y_)) {
    if (kSsSize[(bsize * 2 + P->ss_x_) * 2 + P->ss_y_] == BLOCK_INVALID) {
      corrupted = true;
      return;
    }
  }
  if (intra_only_frame()) read_intra_frame_mode_info();
  else read_inter_frame_mode_info();
  if (corrupted) return;
  if (mi->skip) {  // dec_reset_skip_context
    for (int p = 0; p < 3; p++) {
      const int ss_x = p ? P->ss_x_ : 0, ss_y = p ? P->ss_y_ : 0;
      memset(&P->above_ctx_[p][(col * 2) >> ss_x], 0, (bw << 1) >> ss_x);
      memset(&left_ctx[p][((row * 2) & 15) >> ss_y], 0, (bh << 1) >> ss_y);
    }
  }
  const bool inter = is_inter(mi);
  // emit the prediction block record
  {
    vp9b_block b;
    memset(&b, 0, sizeof(b));
    b.mi_col = (uint16_t)col;
    b.mi_row = (uint16_t)row;
    b.bsize = (uint8_t)bsize;
    b.ref0 = inter ? (uint8_t)mi->ref_frame[0] : 0;
    b.ref1 = (inter && mi->ref_frame[1] > 0) ? (uint8_t)mi->ref_frame[1] : 0;
    b.filter = inter ? mi->interp_filter : 0;
    if (inter) {
      if (bsize < BLOCK_8X8) {
        vp9b_submv s;
        memset(&s, 0, sizeof(s));
        for (int k = 0; k < 4; k++)
          for (int rf = 0; rf < 2; rf++) {
            s.luma[k][rf][0] = mi->bmv[k][rf].row;
            s.luma[k][rf][1] = mi->bmv[k][rf].col;
          }
        for (int rf = 0; rf < 2; rf++) {  // mi_mv_pred_q4 (vp9_reconinter.c:61-87)
          const int sr = mi->bmv[0][rf].row + mi->bmv[1][rf].row + mi->bmv[2][rf].row + mi->bmv[3][rf].row;
          const int sc = mi->bmv[0][rf].col + mi->bmv[1][rf].col + mi->bmv[2][rf].col + mi->bmv[3][rf].col;
          s.chroma[rf][0] = (int16_t)((sr < 0 ? sr - 2 : sr + 2) / 4);
          s.chroma[rf][1] = (int16_t)((sc < 0 ? sc - 2 : sc + 2) / 4);
        }
        const uint32_t idx = (uint32_t)O->submv.size();   // worker-local; rebased by pack()
        O->submv.push_back(s);
        memcpy(&b.mv[0][0], &idx, 4);
      } else {
        for (int rf = 0; rf < 2; rf++) {
          b.mv[rf][0] = mi->mv[rf].row;
          b.mv[rf][1] = mi->mv[rf].col;
        }
      }
      for (int p = 0; p < 3; p++) {
        const int ss_x = p ? P->ss_x_ : 0, ss_y = p ? P->ss_y_ : 0;
        const uint32_t px = (uint32_t)((x_mis * 8) >> ss_x) * ((y_mis * 8) >> ss_y);
        O->n_inter_px[p] += px;
        if (b.ref1) O->n_compound_px[p] += px;
      }
    }
    O->blocks.push_back(b);
  }
  residual_and_record();
  if (r.has_error()) corrupted = true;
  if (P->lf_.filter_level) build_lf_mask();
  // publish mode info to the grid and the MV buffer (set_offsets + vp9_read_mode_info tail)
  for (int y = 0; y < y_mis; ++y) {
    uint32_t *g = &P->grid_[(size_t)(row + y) * P->mi_cols_ + col];
    MvRef *m = &P->cur_mvs_[(size_t)(row + y) * P->mi_cols_ + col];
    for (int x = 0; x < x_mis; ++x) {
      g[x] = bidx;
      m[x].ref_frame[0] = mi->ref_frame[0];
      m[x].ref_frame[1] = mi->ref_frame[1];
      m[x].mv[0] = mi->mv[0];
      m[x].mv[1] = mi->mv[1];
    }
  }
}

// ------------------------------------------------------------------------------------------
// Backward adaptation (vp9_entropy.c:1051-1103, vp9_entropymode.c:341-405, vp9_entropymv.c:157-196)
namespace {
inline uint8_t get_prob(unsigned num, unsigned den) {
  const int p = (int)(((uint64_t)num * 256 + (den >> 1)) / den);
  return (uint8_t)(p > 255 ? 255 : (p < 1 ? 1 : p));
}
inline uint8_t weighted_prob(int p1, int p2, int factor) {
  return (uint8_t)((p1 * (256 - factor) + p2 * factor + 128) >> 8);
}
inline uint8_t merge_probs(uint8_t pre, unsigned c0, unsigned c1, unsigned count_sat, unsigned max_update) {
  const unsigned den = c0 + c1;
  const uint8_t prob = den == 0 ? 128 : get_prob(c0, den);
  const unsigned count = std::min(den, count_sat);
  const unsigned factor = max_update * count / count_sat;
  return weighted_prob(pre, prob, (int)factor);
}
inline uint8_t mode_mv_merge_probs(uint8_t pre, unsigned c0, unsigned c1) {
  const unsigned den = c0 + c1;
  if (den == 0) return pre;
  const unsigned count = std::min(den, 20u);
  return weighted_prob(pre, get_prob(c0, den), kCountToUpdateFactor[count]);
}
unsigned tree_merge_impl(unsigned i, const int8_t *tree, const uint8_t *pre, const uint32_t *counts, uint8_t *probs) {
  const int l = tree[i];
  const unsigned lc = l <= 0 ? counts[-l] : tree_merge_impl(l, tree, pre, counts, probs);
  const int r = tree[i + 1];
  const unsigned rc = r <= 0 ? counts[-r] : tree_merge_impl(r, tree, pre, counts, probs);
  probs[i >> 1] = mode_mv_merge_probs(pre[i >> 1], lc, rc);
  return lc + rc;
}
inline void tree_merge_probs(const int8_t *tree, const uint8_t *pre, const uint32_t *counts, uint8_t *probs) {
  tree_merge_impl(0, tree, pre, counts, probs);
}
}  // namespace

void Parser::adapt_coef_probs() {
  unsigned update_factor, count_sat = 24;
  if (frame_type_ == KEY_FRAME || intra_only_) update_factor = 112;
  else if (last_frame_type_ == KEY_FRAME) update_factor = 128;
  else update_factor = 112;
  const FrameContext &pre = frame_contexts_[frame_context_idx_];
  for (int t = 0; t < 4; t++)
    for (int i = 0; i < 2; ++i)
      for (int j = 0; j < 2; ++j)
        for (int k = 0; k < 6; ++k)
          for (int l = 0; l < (k == 0 ? 3 : 6); ++l) {
            const uint32_t *c = counts_.coef[t][i][j][k][l];
            const unsigned n0 = c[0], n1 = c[1], n2 = c[2], neob = c[3];
            const unsigned eb = counts_.eob_branch[t][i][j][k][l];
            const unsigned bc[3][2] = { { neob, eb - neob }, { n0, n1 + n2 }, { n1, n2 } };
            for (int m = 0; m < 3; ++m)
              fc_.coef[t][i][j][k][l][m] =
                  merge_probs(pre.coef[t][i][j][k][l][m], bc[m][0], bc[m][1], count_sat, update_factor);
          }
}

void Parser::adapt_mode_probs() {
  const FrameContext &pre = frame_contexts_[frame_context_idx_];
  const FrameCounts &c = counts_;
  for (int i = 0; i < 4; i++) fc_.intra_inter[i] = mode_mv_merge_probs(pre.intra_inter[i], c.intra_inter[i][0], c.intra_inter[i][1]);
  for (int i = 0; i < 5; i++) fc_.comp_inter[i] = mode_mv_merge_probs(pre.comp_inter[i], c.comp_inter[i][0], c.comp_inter[i][1]);
  for (int i = 0; i < 5; i++) fc_.comp_ref[i] = mode_mv_merge_probs(pre.comp_ref[i], c.comp_ref[i][0], c.comp_ref[i][1]);
  for (int i = 0; i < 5; i++)
    for (int j = 0; j < 2; j++)
      fc_.single_ref[i][j] = mode_mv_merge_probs(pre.single_ref[i][j], c.single_ref[i][j][0], c.single_ref[i][j][1]);
  for (int i = 0; i < 7; i++) tree_merge_probs(kInterModeTree, pre.inter_mode[i], c.inter_mode[i], fc_.inter_mode[i]);
  for (int i = 0; i < 4; i++) tree_merge_probs(kIntraModeTree, pre.y_mode[i], c.y_mode[i], fc_.y_mode[i]);
  for (int i = 0; i < 10; ++i) tree_merge_probs(kIntraModeTree, pre.uv_mode[i], c.uv_mode[i], fc_.uv_mode[i]);
  for (int i = 0; i < 16; i++) tree_merge_probs(kPartitionTree, pre.partition[i], c.partition[i], fc_.partition[i]);
  if (interp_filter_ == SWITCHABLE)
    for (int i = 0; i < 4; i++)
      tree_merge_probs(kSwitchableInterpTree, pre.switchable_interp[i], c.switchable_interp[i], fc_.switchable_interp[i]);
  if (tx_mode_ == TX_MODE_SELECT) {
    for (int i = 0; i < 2; ++i) {
      fc_.tx8[i][0] = mode_mv_merge_probs(pre.tx8[i][0], c.tx8[i][0], c.tx8[i][1]);
      const uint32_t *t16 = c.tx16[i];
      fc_.tx16[i][0] = mode_mv_merge_probs(pre.tx16[i][0], t16[0], t16[1] + t16[2]);
      fc_.tx16[i][1] = mode_mv_merge_probs(pre.tx16[i][1], t16[1], t16[2]);
      const uint32_t *t32 = c.tx32[i];
      fc_.tx32[i][0] = mode_mv_merge_probs(pre.tx32[i][0], t32[0], t32[1] + t32[2] + t32[3]);
      fc_.tx32[i][1] = mode_mv_merge_probs(pre.tx32[i][1], t32[1], t32[2] + t32[3]);
      fc_.tx32[i][2] = mode_mv_merge_probs(pre.tx32[i][2], t32[2], t32[3]);
    }
  }
  for (int i = 0; i < 3; ++i) fc_.skip[i] = mode_mv_merge_probs(pre.skip[i], c.skip[i][0], c.skip[i][1]);
}

void Parser::adapt_mv_probs() {
  const FrameContext &pre = frame_contexts_[frame_context_idx_];
  tree_merge_probs(kMvJointTree, pre.mv_joints, counts_.mv_joints, fc_.mv_joints);
  for (int i = 0; i < 2; ++i) {
    MvCompProbs &comp = fc_.mvc[i];
    const MvCompProbs &pc = pre.mvc[i];
    const MvCompCounts &c = counts_.mvc[i];
    comp.sign = mode_mv_merge_probs(pc.sign, c.sign[0], c.sign[1]);
    tree_merge_probs(kMvClassTree, pc.classes, c.classes, comp.classes);
    tree_merge_probs(kMvClass0Tree, pc.class0, c.class0, comp.class0);
    for (int j = 0; j < 10; ++j) comp.bits[j] = mode_mv_merge_probs(pc.bits[j], c.bits[j][0], c.bits[j][1]);
    for (int j = 0; j < 2; ++j) tree_merge_probs(kMvFpTree, pc.class0_fp[j], c.class0_fp[j], comp.class0_fp[j]);
    tree_merge_probs(kMvFpTree, pc.fp, c.fp, comp.fp);
    if (allow_hp_) {
      comp.class0_hp = mode_mv_merge_probs(pc.class0_hp, c.class0_hp[0], c.class0_hp[1]);
      comp.hp = mode_mv_merge_probs(pc.hp, c.hp[0], c.hp[1]);
    }
  }
}

// vp9_adjust_mask (vp9_loopfilter.c:766-878) applied to every superblock before upload.
// Chroma masks of a 4:2:2 / 4:4:0 frame: the mask construction of vp9_filter_block_plane_non420
// (vp9_loopfilter.c:1080-1239) for every superblock, from the finished mode info; the frame-edge rules
// of the two filter passes (left-most column, top row, last odd row) are folded in.
void Parser::build_lfm_non420() {
  const int ss_x = ss_x_, ss_y = ss_y_, row_step = 1 << ss_y, col_step = 1 << ss_x;
  vp9b_lfm_uv z;
  memset(&z, 0, sizeof(z));
  out_lfm_uv_.assign((size_t)sb_cols_ * sb_rows_, z);
  for (int sbr = 0; sbr < sb_rows_; sbr++)
    for (int sbc = 0; sbc < sb_cols_; sbc++) {
      vp9b_lfm_uv &m = out_lfm_uv_[(size_t)sbr * sb_cols_ + sbc];
      const int mi_row = sbr * 8, mi_col = sbc * 8;
      for (int r = 0; r < 8 && mi_row + r < mi_rows_; r += row_step) {
        for (int c = 0; c < 8 && mi_col + c < mi_cols_; c += col_step) {
          const BlockInfo &mi = blocks_[grid_[(size_t)(mi_row + r) * mi_cols_ + mi_col + c]];
          const int sb_type = mi.sb_type;
          const bool skip_this = mi.skip && mi.ref_frame[0] > INTRA_FRAME;
          // left / top edge of the unit is a block edge -> not skipped
          const bool edge_left = kNum4x4Wide[sb_type] > 1 ? !(c & (kNum8x8Wide[sb_type] - 1)) : true;
          const bool edge_above = kNum4x4High[sb_type] > 1 ? !(r & (kNum8x8High[sb_type] - 1)) : true;
          const bool skip_c = skip_this && !edge_left, skip_r = skip_this && !edge_above;
          const int tx = kUvTxSize[((sb_type * 4 + mi.tx_size) * 2 + ss_x) * 2 + ss_y];
          const bool border_c = ss_x && mi_col + c == mi_cols_ - 1, border_r = ss_y && mi_row + r == mi_rows_ - 1;
          const int level = lf_lvl_[mi.segment_id][mi.ref_frame[0] < 0 ? 0 : mi.ref_frame[0]][kModeLfLut[mi.mode]];
          const uint8_t bit = (uint8_t)(1u << (c >> ss_x));
          m.lfl[(r << 3) + (c >> ss_x)] = (uint8_t)level;
          if (!level) continue;
          if (tx == TX_32X32 || tx == TX_16X16) {
            const int al = tx == TX_32X32 ? 3 : 1;
            if (!skip_c && ((c >> ss_x) & al) == 0) (border_c ? m.v8 : m.v16)[r] |= bit;
            if (!skip_r && ((r >> ss_y) & al) == 0) (border_r ? m.h8 : m.h16)[r] |= bit;
          } else {
            if (!skip_c) ((tx == TX_8X8 || ((c >> ss_x) & 3) == 0) ? m.v8 : m.v4)[r] |= bit;   // 8x8 filter on 32x32 boundaries
            if (!skip_r) ((tx == TX_8X8 || ((r >> ss_y) & 3) == 0) ? m.h8 : m.h4)[r] |= bit;
            if (!skip_this && tx < TX_8X8 && !border_c) m.v4i[r] |= bit;
          }
        }
        // horizontal pass: its inner edges are the same mask minus the last odd row; no top edge in frame row 0
        m.h4i[r] = (ss_y && mi_row + r == mi_rows_ - 1) ? 0 : m.v4i[r];
        if (mi_row + r == 0) m.h16[r] = m.h8[r] = m.h4[r] = 0;
        if (mi_col == 0) { m.v16[r] &= 0xfe; m.v8[r] &= 0xfe; m.v4[r] &= 0xfe; }   // left-most column of the frame
      }
    }
}

void Parser::finish_lfm() {
  const uint64_t left_border = 0x1111111111111111ULL, above_border = 0x000000ff000000ffULL;
  const uint16_t left_border_uv = 0x1111, above_border_uv = 0x000f;
  for (int sbr = 0; sbr < sb_rows_; sbr++)
    for (int sbc = 0; sbc < sb_cols_; sbc++) {
      vp9b_lfm &m = out_lfm_[sbr * sb_cols_ + sbc];
      const int mi_row = sbr * 8, mi_col = sbc * 8;
      m.left_y[2] |= m.left_y[3];
      m.above_y[2] |= m.above_y[3];
      m.left_uv[2] |= m.left_uv[3];
      m.above_uv[2] |= m.above_uv[3];
      m.left_y[1] |= m.left_y[0] & left_border;
      m.left_y[0] &= ~left_border;
      m.above_y[1] |= m.above_y[0] & above_border;
      m.above_y[0] &= ~above_border;
      m.left_uv[1] |= m.left_uv[0] & left_border_uv;
      m.left_uv[0] &= ~left_border_uv;
      m.above_uv[1] |= m.above_uv[0] & above_border_uv;
      m.above_uv[0] &= ~above_border_uv;
      if (mi_row + 8 > mi_rows_) {
        const uint64_t rows = mi_rows_ - mi_row;
        const uint64_t mask_y = (((uint64_t)1 << (rows << 3)) - 1);
        const uint16_t mask_uv = (uint16_t)(((uint16_t)1 << (((rows + 1) >> 1) << 2)) - 1);
        for (int i = 0; i < 3; i++) {
          m.left_y[i] &= mask_y;
          m.above_y[i] &= mask_y;
          m.left_uv[i] &= mask_uv;
          m.above_uv[i] &= mask_uv;
        }
        m.int_4x4_y &= mask_y;
        m.int_4x4_uv &= mask_uv;
        if (rows == 1) {
          m.above_uv[1] |= m.above_uv[2];
          m.above_uv[2] = 0;
        }
        if (rows == 5) {
          m.above_uv[1] |= m.above_uv[2] & 0xff00;
          m.above_uv[2] &= ~(m.above_uv[2] & 0xff00);
        }
      }
      if (mi_col + 8 > mi_cols_) {
        const uint64_t columns = mi_cols_ - mi_col;
        const uint64_t mask_y = (((1 << columns) - 1)) * 0x0101010101010101ULL;
        const uint16_t mask_uv = (uint16_t)(((1 << ((columns + 1) >> 1)) - 1) * 0x1111);
        const uint16_t mask_uv_int = (uint16_t)(((1 << (columns >> 1)) - 1) * 0x1111);
        for (int i = 0; i < 3; i++) {
          m.left_y[i] &= mask_y;
          m.above_y[i] &= mask_y;
          m.left_uv[i] &= mask_uv;
          m.above_uv[i] &= mask_uv;
        }
        m.int_4x4_y &= mask_y;
        m.int_4x4_uv &= mask_uv_int;
        if (columns == 1) {
          m.left_uv[1] |= m.left_uv[2];
          m.left_uv[2] = 0;
        }
        if (columns == 5) {
          m.left_uv[1] |= (m.left_uv[2] & 0xcccc);
          m.left_uv[2] &= ~(m.left_uv[2] & 0xcccc);
        }
      }
      if (mi_col == 0) {
        for (int i = 0; i < 3; i++) {
          m.left_y[i] &= 0xfefefefefefefefeULL;
          m.left_uv[i] &= 0xeeee;
        }
      }
      m.left_y[3] = m.above_y[3] = 0;
      m.left_uv[3] = m.above_uv[3] = 0;
    }
}

// Assemble the packed buffer: header + 16-byte aligned sections.
void Parser::pack() {
  vp9b_frame_hdr h;
  memset(&h, 0, sizeof(h));
  h.magic = VP9B_MAGIC;
  h.width = width_;
  h.height = height_;
  h.mi_cols = mi_cols_;
  h.mi_rows = mi_rows_;
  h.sb_cols = sb_cols_;
  h.sb_rows = sb_rows_;
  h.bit_depth = (uint8_t)bit_depth_;
  h.ss_x = (uint8_t)ss_x_;
  h.ss_y = (uint8_t)ss_y_;
  h.lossless = (uint8_t)lossless_;
  h.intra_only = (frame_type_ == KEY_FRAME || intra_only_);
  h.show_frame = (uint8_t)show_frame_;
  h.lf_level = (uint8_t)lf_.filter_level;
  h.lf_sharpness = (uint8_t)lf_.sharpness_level;
  h.coef_bytes = bit_depth_ > 8 ? 4 : 2;
  h.dst_slot = new_slot_;
  for (int i = 0; i < 3; i++) h.ref[i] = ref_desc_[i];
  h.frame_index = frame_index_;
  // update_sharpness (vp9_loopfilter.c:212-231)
  for (int lvl = 0; lvl < 64; lvl++) {
    const int sh = lf_.sharpness_level;
    int lim = lvl >> ((sh > 0) + (sh > 4));
    if (sh > 0 && lim > 9 - sh) lim = 9 - sh;
    if (lim < 1) lim = 1;
    h.lf_lim[lvl] = (uint8_t)lim;
    h.lf_mblim[lvl] = (uint8_t)(2 * (lvl + 2) + lim);
  }
  // Sections are concatenations of the tile-column workers' lists; worker-local indices (sub-8x8 MV
  // table, coefficient offsets, per-superblock intra ranges) are rebased while copying.
  const int nw = n_workers_;
  size_t n_blocks = 0, n_submv = 0, n_inter[4] = { 0, 0, 0, 0 }, n_intra[3] = { 0, 0, 0 }, n_coef = 0;
  std::vector<size_t> submv_base(nw), coef_base(nw), intra_base(3 * nw);
  for (int w = 0; w < nw; w++) {
    const TileOut &o = *outs_[w];
    submv_base[w] = n_submv;
    coef_base[w] = n_coef;
    n_blocks += o.blocks.size();
    n_submv += o.submv.size();
    n_coef += o.n_coef;
    for (int i = 0; i < 4; i++) n_inter[i] += o.txb_inter[i].size();
    for (int p = 0; p < 3; p++) {
      intra_base[p * nw + w] = n_intra[p];
      n_intra[p] += o.txb_intra[p].size();
    }
    for (int p = 0; p < 3; p++) {
      h.n_inter_px[p] += o.n_inter_px[p];
      h.n_compound_px[p] += o.n_compound_px[p];
    }
  }
  const size_t coef_elem = bit_depth_ > 8 ? 4 : 2;
  size_t count[VP9B_SEC_COUNT], bytes[VP9B_SEC_COUNT];
  memset(count, 0, sizeof(count));
  memset(bytes, 0, sizeof(bytes));
  auto set = [&](int id, size_t elem, size_t n) {
    count[id] = n;
    bytes[id] = elem * n;
  };
  set(VP9B_SEC_BLOCKS, sizeof(vp9b_block), n_blocks);
  set(VP9B_SEC_SUBMV, sizeof(vp9b_submv), n_submv);
  set(VP9B_SEC_TXB_INTER, sizeof(vp9b_txb), n_inter[0] + n_inter[1] + n_inter[2] + n_inter[3]);
  for (int i = 0; i < 4; i++) h.n_txb_inter_size[i] = (uint32_t)n_inter[i];
  for (int p = 0; p < 3; p++) set(VP9B_SEC_TXB_INTRA_Y + p, sizeof(vp9b_txb), n_intra[p]);
  set(VP9B_SEC_SB_INTRA, sizeof(vp9b_sbrange), out_sb_intra_.size());
  set(VP9B_SEC_COEF, coef_elem, n_coef);
  if (lf_.filter_level) set(VP9B_SEC_LFM, sizeof(vp9b_lfm), out_lfm_.size());
  if (!out_lfm_uv_.empty()) set(VP9B_SEC_LFM_UV, sizeof(vp9b_lfm_uv), out_lfm_uv_.size());
  size_t off = (sizeof(h) + 15) & ~(size_t)15;
  for (int i = 0; i < VP9B_SEC_COUNT; i++) {
    h.sec[i].offset = (uint32_t)off;
    h.sec[i].count = (uint32_t)count[i];
    off += (bytes[i] + 15) & ~(size_t)15;
  }
  h.total_bytes = (uint32_t)off;
  // straight into the caller's (page-locked) upload buffer when there is one: no second copy
  uint8_t *dst = out_alloc_ ? (uint8_t *)out_alloc_(out_priv_, off) : nullptr;
  if (dst) {
    ext_out_ = dst;
    ext_bytes_ = off;
  } else {
    packed_.resize(off);
    dst = packed_.data();
  }
  memcpy(dst, &h, sizeof(h));
  auto copy_txb = [](vp9b_txb *d, const std::vector<vp9b_txb> &v, size_t cbase) {
    memcpy(d, v.data(), v.size() * sizeof(vp9b_txb));
    if (cbase)
      for (size_t i = 0; i < v.size(); i++) d[i].coef_off += (uint32_t)cbase;
    return d + v.size();
  };
  {
    vp9b_block *db = (vp9b_block *)(dst + h.sec[VP9B_SEC_BLOCKS].offset);
    vp9b_submv *ds = (vp9b_submv *)(dst + h.sec[VP9B_SEC_SUBMV].offset);
    uint8_t *dc = dst + h.sec[VP9B_SEC_COEF].offset;
    vp9b_txb *di[3];
    for (int p = 0; p < 3; p++) di[p] = (vp9b_txb *)(dst + h.sec[VP9B_SEC_TXB_INTRA_Y + p].offset);
    for (int w = 0; w < nw; w++) {
      const TileOut &o = *outs_[w];
      memcpy(db, o.blocks.data(), o.blocks.size() * sizeof(vp9b_block));
      if (submv_base[w])
        for (size_t i = 0; i < o.blocks.size(); i++)
          if (db[i].ref0 && db[i].bsize < 3) {
            uint32_t idx;
            memcpy(&idx, &db[i].mv[0][0], 4);
            idx += (uint32_t)submv_base[w];
            memcpy(&db[i].mv[0][0], &idx, 4);
          }
      db += o.blocks.size();
      memcpy(ds, o.submv.data(), o.submv.size() * sizeof(vp9b_submv));
      ds += o.submv.size();
      memcpy(dc, bit_depth_ > 8 ? (const void *)o.coef32.data() : (const void *)o.coef16.data(), o.n_coef * coef_elem);
      dc += o.n_coef * coef_elem;
      for (int p = 0; p < 3; p++) di[p] = copy_txb(di[p], o.txb_intra[p], coef_base[w]);
    }
    vp9b_txb *dt = (vp9b_txb *)(dst + h.sec[VP9B_SEC_TXB_INTER].offset);
    for (int i = 0; i < 4; i++)          // grouped by transform size
      for (int w = 0; w < nw; w++) dt = copy_txb(dt, outs_[w]->txb_inter[i], coef_base[w]);
    // per-superblock intra ranges: rebase to the owning worker's position in the merged lists
    vp9b_sbrange *dr = (vp9b_sbrange *)(dst + h.sec[VP9B_SEC_SB_INTRA].offset);
    memcpy(dr, out_sb_intra_.data(), out_sb_intra_.size() * sizeof(vp9b_sbrange));
    if (nw > 1) {
      const int nsb = sb_cols_ * sb_rows_;
      for (int p = 0; p < 3; p++)
        for (int sbr = 0; sbr < sb_rows_; sbr++)
          for (int sbc = 0; sbc < sb_cols_; sbc++)
            dr[p * nsb + sbr * sb_cols_ + sbc].start += (uint32_t)intra_base[p * nw + sb_worker_[sbc]];
    }
    if (lf_.filter_level) memcpy(dst + h.sec[VP9B_SEC_LFM].offset, out_lfm_.data(), out_lfm_.size() * sizeof(vp9b_lfm));
    if (!out_lfm_uv_.empty())
      memcpy(dst + h.sec[VP9B_SEC_LFM_UV].offset, out_lfm_uv_.data(), out_lfm_uv_.size() * sizeof(vp9b_lfm_uv));
  }
}

// first mi column / row of tile idx (get_tile_offset, vp9_tile_common.c:20-27)
static inline int tile_col_offset(int idx, int mis, int log2) {
  const int sb = (mis + 7) >> 3;
  const int off = ((idx * sb) >> log2) << 3;
  return std::min(off, mis);
}

// All tiles of a group of tile columns, column by column, top to bottom (role of tile_worker_hook,
// vp9_decodeframe.c:2663-2752).  The block array range of a tile column starts at
// mi_rows * (first mi column): at most one BlockInfo per 8x8 cell, so ranges never overlap.
void Parser::run_tile_column(TileOut *o, const TileJob *jobs, int n_jobs) {
  for (int i = 0; i < n_jobs; i++) {
    const TileJob &j = jobs[i];
    TileCtx t;
    t.P = this;
    t.O = o;
    t.mi_row_start = tile_col_offset(j.tr, mi_rows_, log2_tile_rows_);
    t.mi_row_end = tile_col_offset(j.tr + 1, mi_rows_, log2_tile_rows_);
    t.mi_col_start = tile_col_offset(j.tc, mi_cols_, log2_tile_cols_);
    t.mi_col_end = tile_col_offset(j.tc + 1, mi_cols_, log2_tile_cols_);
    if (j.tr == 0) o->block_next = (uint32_t)((size_t)mi_rows_ * t.mi_col_start);
    if (t.r.init(j.data, j.size)) { o->bad_marker = 1; return; }
    t.decode_tile();
    if (t.corrupted) { o->corrupted = true; return; }
  }
}

// vp9_receive_compressed_data + vp9_decode_frame (vp9_decoder.c:407-489, vp9_decodeframe.c:3476-3587)
int Parser::parse(const uint8_t *data, size_t size) {
  adapt_worker_.wait();   // the previous frame's backward adaptation
  packed_.clear();
  ext_out_ = nullptr;
  ext_bytes_ = 0;
  show_slot_ = -1;
  err_.clear();
  if (!data || size == 0) return fail(VP9B_ERR_INVALID, "empty frame");
  size_t hdr_bytes = 0, first_partition_size = 0;
  new_slot_ = find_free_slot();
  if (new_slot_ < 0) return fail(VP9B_ERR_MEM, "no free frame slot");
  int e = read_uncompressed_header(data, size, &hdr_bytes, &first_partition_size);
  if (e) return e;
  if (show_existing_) {
    // shown directly; no decode, no state update besides "last shown" bookkeeping
    last_shown_slot_ = show_slot_;
    last_width_ = width_;
    last_height_ = height_;
    return VP9B_OK;
  }
  if (hdr_bytes + first_partition_size > size) return fail(VP9B_ERR_CORRUPT, "truncated packet or corrupt header length");
  use_prev_frame_mvs_ = !error_res_ && width_ == last_width_ && height_ == last_height_ &&
                        !last_intra_only_ && last_show_frame_ && prev_mvs_valid_;
  fc_ = frame_contexts_[frame_context_idx_];
  if (!fc_.initialized) return fail(VP9B_ERR_CORRUPT, "uninitialised entropy context");
  e = read_compressed_header(data + hdr_bytes, first_partition_size);
  if (e) return e;
  do_counts_ = !frame_parallel_;
  if (do_counts_) memset(&counts_, 0, sizeof(counts_));
  if (lf_.filter_level) loop_filter_frame_init();
  // per-frame buffers
  const int aligned_cols = sb_cols_ * 8;
  for (int p = 0; p < 3; p++) above_ctx_[p].assign((size_t)aligned_cols * 2 + 16, 0);
  above_seg_ctx_.assign(aligned_cols + 8, 0);
  if (blocks_.size() < (size_t)mi_cols_ * mi_rows_ + 64) blocks_.resize((size_t)mi_cols_ * mi_rows_ + 64);
  out_sb_intra_.assign((size_t)3 * sb_cols_ * sb_rows_, vp9b_sbrange{ 0, 0 });
  out_lfm_.clear();
  if (lf_.filter_level) {
    vp9b_lfm z;
    memset(&z, 0, sizeof(z));
    out_lfm_.assign((size_t)sb_cols_ * sb_rows_, z);
  }
  // tiles (get_tile_buffers + decode_tiles, vp9_decodeframe.c:1877-1925, 2303-2430): the sizes first,
  // then one worker per group of tile columns -- tile columns share nothing but the previous frame's
  // motion vectors (decode_tiles_mt, :2919-3047); a column's tiles are parsed top to bottom.
  const uint8_t *p = data + hdr_bytes + first_partition_size;
  const uint8_t *end = data + size;
  const int tile_rows = 1 << log2_tile_rows_, tile_cols = 1 << log2_tile_cols_;
  std::vector<TileJob> jobs((size_t)tile_rows * tile_cols);     // [tc][tr]
  for (int tr = 0; tr < tile_rows; tr++) {
    for (int tc = 0; tc < tile_cols; tc++) {
      const bool last = (tr == tile_rows - 1) && (tc == tile_cols - 1);
      size_t tsz;
      if (!last) {
        if (end - p < 4) return fail(VP9B_ERR_CORRUPT, "truncated packet or corrupt tile length");
        tsz = ((size_t)p[0] << 24) | ((size_t)p[1] << 16) | ((size_t)p[2] << 8) | p[3];
        p += 4;
        if (tsz > (size_t)(end - p)) return fail(VP9B_ERR_CORRUPT, "truncated packet or corrupt tile size");
      } else {
        tsz = end - p;
      }
      if (tsz == 0) return fail(VP9B_ERR_CORRUPT, "empty tile");
      jobs[(size_t)tc * tile_rows + tr] = TileJob{ p, tsz, tr, tc };
      p += tsz;
    }
  }
  n_workers_ = std::min(threads_, tile_cols);
  while ((int)outs_.size() < n_workers_) outs_.emplace_back(new TileOut());
  while ((int)tile_workers_.size() < n_workers_ - 1) tile_workers_.emplace_back(new SideWorker());
  sb_worker_.assign(sb_cols_, 0);
  std::vector<int> first(n_workers_ + 1);
  for (int w = 0; w <= n_workers_; w++) first[w] = w * tile_cols / n_workers_;   // contiguous tile columns per worker
  for (int w = 0; w < n_workers_; w++) {
    TileOut &o = *outs_[w];
    o.blocks.clear();
    o.submv.clear();
    for (int i = 0; i < 4; i++) o.txb_inter[i].clear();
    for (int q = 0; q < 3; q++) o.txb_intra[q].clear();
    o.n_coef = 0;
    if (do_counts_) memset(&o.counts, 0, sizeof(o.counts));
    memset(o.n_inter_px, 0, sizeof(o.n_inter_px));
    memset(o.n_compound_px, 0, sizeof(o.n_compound_px));
    o.corrupted = false;
    o.bad_marker = 0;
    for (int tc = first[w]; tc < first[w + 1]; tc++) {
      const int c0 = tile_col_offset(tc, mi_cols_, log2_tile_cols_) >> 3, c1 = (tile_col_offset(tc + 1, mi_cols_, log2_tile_cols_) + 7) >> 3;
      for (int c = c0; c < c1 && c < sb_cols_; c++) sb_worker_[c] = (uint8_t)w;
    }
  }
  for (int w = n_workers_ - 1; w >= 0; w--) {
    TileOut *o = outs_[w].get();
    const TileJob *j = &jobs[(size_t)first[w] * tile_rows];
    const int nj = (first[w + 1] - first[w]) * tile_rows;
    if (w > 0) tile_workers_[w - 1]->run([this, o, j, nj] { run_tile_column(o, j, nj); });
    else run_tile_column(o, j, nj);      // the calling thread takes the first group itself
  }
  for (int w = 1; w < n_workers_; w++) tile_workers_[w - 1]->wait();
  for (int w = 0; w < n_workers_; w++) {
    if (outs_[w]->bad_marker) return fail(VP9B_ERR_CORRUPT, "tile marker bit set");
    if (outs_[w]->corrupted) return fail(VP9B_ERR_CORRUPT, "corrupt tile data");
  }
  if (do_counts_) {      // symbol counts of the frame = sum over the workers
    counts_ = outs_[0]->counts;
    for (int w = 1; w < n_workers_; w++) {
      uint32_t *d = (uint32_t *)&counts_;
      const uint32_t *a = (const uint32_t *)&outs_[w]->counts;
      for (size_t i = 0; i < sizeof(FrameCounts) / sizeof(uint32_t); i++) d[i] += a[i];
    }
  }
  // Backward adaptation (vp9_decodeframe.c:3570-3586) only feeds the NEXT frame's entropy context: it
  // runs on a side thread while this thread finishes the loop-filter masks, assembles the packed
  // buffer and hands the frame to the GPU; parse() of the next frame joins it before it touches
  // the contexts.  (fc_, counts_ and frame_contexts_ are not used by anything below.)
  {
    const bool adapt = !error_res_ && !frame_parallel_;
    const bool inter = !(frame_type_ == KEY_FRAME || intra_only_);
    auto job = [this, adapt, inter] {
      if (adapt) {
        adapt_coef_probs();
        if (inter) {
          adapt_mode_probs();
          adapt_mv_probs();
        }
      }
      if (refresh_frame_context_) frame_contexts_[frame_context_idx_] = fc_;
    };
    static const bool inline_adapt = getenv("VP9B_SYNC_ADAPT") != nullptr;
    if ((adapt || refresh_frame_context_) && !inline_adapt) adapt_worker_.run(job);
    else job();
  }
  if (lf_.filter_level) finish_lfm();
  out_lfm_uv_.clear();
  if (lf_.filter_level && ss_x_ != ss_y_) build_lfm_non420();
  pack();
  // swap_frame_buffers (vp9_decoder.c:334-367) + tail of vp9_receive_compressed_data
  SlotInfo &s = slots_[new_slot_];
  s.width = width_;
  s.height = height_;
  s.bit_depth = bit_depth_;
  s.ss_x = ss_x_;
  s.ss_y = ss_y_;
  s.valid = true;
  for (int i = 0; i < 8; i++)
    if (refresh_frame_flags_ & (1 << i)) ref_frame_map_[i] = new_slot_;
  last_show_frame_ = show_frame_;
  std::swap(cur_mvs_, prev_mvs_);
  prev_mvs_valid_ = true;
  if (seg_.enabled) std::swap(seg_map_cur_, seg_map_last_);
  if (show_frame_) {
    show_slot_ = new_slot_;
    last_shown_slot_ = new_slot_;
  }
  last_width_ = width_;
  last_height_ = height_;
  frame_index_++;
  return VP9B_OK;
}

}  // namespace vp9b
