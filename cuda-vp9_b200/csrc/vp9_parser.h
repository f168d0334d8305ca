// Host-side VP9 bitstream parser: compressed frame -> packed per-frame buffer (vp9_b200.h).
// Takes the role the libvpx host keeps in the reference design: bool decoder
// (vpx_dsp/bitreader.*), headers (vp9_decodeframe.c:3114-3388), mode/MV parse
// (vp9_decodemv.c:811), detokenize + dequant (vp9_detokenize.c:115-333), loop-filter mask
// build (vp9_loopfilter.c:1528-1608, 766-878) and backward adaptation
// (vp9_entropy.c:1051, vp9_entropymode.c, vp9_entropymv.c:161).  Written from scratch.
#pragma once
#include <stddef.h>
#include <stdint.h>

#include <condition_variable>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>
#include <string>
#include <vector>

#include "vp9_b200.h"

namespace vp9b {

struct Mv {
  int16_t row, col;
};
inline bool operator==(const Mv &a, const Mv &b) { return a.row == b.row && a.col == b.col; }
inline bool operator!=(const Mv &a, const Mv &b) { return !(a == b); }

struct BlockInfo {  // one per decoded prediction block
  uint8_t sb_type, mode, uv_mode, tx_size;
  uint8_t skip, segment_id, seg_id_predicted, interp_filter;
  int8_t ref_frame[2];
  Mv mv[2];
  Mv bmv[4][2];
  uint8_t bmode[4];
};

struct MvRef {  // per 8x8, saved for the next frame's temporal MV candidate
  int8_t ref_frame[2];
  Mv mv[2];
};

struct MvCompProbs {
  uint8_t sign, classes[10], class0[1], bits[10], class0_fp[2][3], fp[3], class0_hp, hp;
};

struct FrameContext {
  uint8_t y_mode[4][9], uv_mode[10][9], partition[16][3];
  uint8_t coef[4][2][2][6][6][3];
  uint8_t switchable_interp[4][2], inter_mode[7][3], intra_inter[4], comp_inter[5];
  uint8_t single_ref[5][2], comp_ref[5];
  uint8_t tx32[2][3], tx16[2][2], tx8[2][1];
  uint8_t skip[3];
  uint8_t mv_joints[3];
  MvCompProbs mvc[2];
  int initialized;
};

struct MvCompCounts {
  uint32_t sign[2], classes[11], class0[2], bits[10][2], class0_fp[2][4], fp[4], class0_hp[2],
      hp[2];
};

struct FrameCounts {
  uint32_t y_mode[4][10], uv_mode[10][10], partition[16][4];
  uint32_t coef[4][2][2][6][6][4], eob_branch[4][2][2][6][6];
  uint32_t switchable_interp[4][3], inter_mode[7][4], intra_inter[4][2], comp_inter[5][2];
  uint32_t single_ref[5][2][2], comp_ref[5][2];
  uint32_t tx32[2][4], tx16[2][3], tx8[2][2];
  uint32_t skip[3][2];
  uint32_t mv_joints[4];
  MvCompCounts mvc[2];
};

struct Segmentation {
  uint8_t enabled, update_map, update_data, abs_delta, temporal_update;
  uint8_t tree_probs[7], pred_probs[3];
  int16_t feature_data[8][4];
  uint32_t feature_mask[8];
};

struct LoopFilterParams {
  int filter_level, sharpness_level;
  uint8_t mode_ref_delta_enabled, mode_ref_delta_update;
  int8_t ref_deltas[4], mode_deltas[2];
};

struct SlotInfo {
  uint32_t width = 0, height = 0;
  int bit_depth = 8, ss_x = 1, ss_y = 1;
  bool valid = false;
};

class BoolReader;
struct TileCtx;

// One helper thread per parser, started on first use: runs the previous frame's backward adaptation
// while the parsing thread packs and submits the frame (creating a thread per frame costs more than
// the adaptation itself).
class SideWorker {
 public:
  ~SideWorker() {
    {
      std::lock_guard<std::mutex> l(m_);
      quit_ = true;
    }
    cv_.notify_all();
    if (th_.joinable()) th_.join();
  }
  void run(std::function<void()> f) {
    {
      std::lock_guard<std::mutex> l(m_);
      job_ = std::move(f);
      has_ = true;
      done_ = false;
    }
    if (!th_.joinable()) th_ = std::thread([this] { loop(); });
    cv_.notify_all();
  }
  void wait() {
    std::unique_lock<std::mutex> l(m_);
    cv_.wait(l, [this] { return done_; });
  }

 private:
  void loop() {
    std::unique_lock<std::mutex> l(m_);
    for (;;) {
      cv_.wait(l, [this] { return has_ || quit_; });
      if (quit_) return;
      std::function<void()> f = std::move(job_);
      has_ = false;
      l.unlock();
      f();
      l.lock();
      done_ = true;
      cv_.notify_all();
    }
  }
  std::thread th_;
  std::mutex m_;
  std::condition_variable cv_;
  std::function<void()> job_;
  bool has_ = false, done_ = true, quit_ = false;
};

// What one tile-column worker produces for a frame (tile columns are independent:
// vp9_decodeframe.c:2919-3047 decode_tiles_mt / :2663-2752 tile_worker_hook): record lists with
// worker-local indices, its share of the coefficient stream, its symbol counts.  Parser::pack()
// concatenates the workers' lists and rebases the indices; the counts are summed before adaptation.
struct TileOut {
  std::vector<vp9b_block> blocks;
  std::vector<vp9b_submv> submv;
  std::vector<vp9b_txb> txb_inter[4];   // by transform size (blocks of one size share a warp on the GPU)
  std::vector<vp9b_txb> txb_intra[3];
  std::vector<int16_t> coef16;
  std::vector<int32_t> coef32;
  size_t n_coef = 0;                    // coefficients written this frame (the vectors only grow)
  FrameCounts counts;
  uint32_t n_inter_px[3], n_compound_px[3];
  uint32_t block_next = 0;              // next free entry of this worker's range of Parser::blocks_
  bool corrupted = false;
  int bad_marker = 0;
};

class Parser {
 public:
  Parser();
  ~Parser();
  // returns VP9B_OK or error; packed()/packed_bytes() valid until next call
  int parse(const uint8_t *data, size_t size);
  const uint8_t *packed() const { return ext_out_ ? ext_out_ : (packed_.empty() ? nullptr : packed_.data()); }
  size_t packed_bytes() const { return ext_out_ ? ext_bytes_ : packed_.size(); }
  // Where pack() assembles the packed buffer: alloc(priv, bytes) returns the destination (e.g. the
  // page-locked upload buffer of a reconstruction context); nullptr / no allocator: a parser-owned buffer.
  void set_output_allocator(void *(*alloc)(void *priv, size_t bytes), void *priv) { out_alloc_ = alloc; out_priv_ = priv; }
  // Tile-column workers per frame (1 = everything on the calling thread).  More than one pays off when
  // there are fewer streams than host cores; the helpers are threads owned by this parser.
  void set_threads(int n) { threads_ = n < 1 ? 1 : (n > 64 ? 64 : n); }
  int show_slot() const { return show_slot_; }
  const SlotInfo &slot(int i) const { return slots_[i]; }
  const std::string &error() const { return err_; }

 private:
  friend struct TileCtx;
  int fail(int code, const char *msg);
  int read_uncompressed_header(const uint8_t *data, size_t size, size_t *hdr_bytes,
                               size_t *first_partition_size);
  int read_compressed_header(const uint8_t *data, size_t size);
  void setup_past_independence();
  void setup_frame_size(int w, int h);
  int find_free_slot();
  void loop_filter_frame_init();
  void adapt_coef_probs();
  void adapt_mode_probs();
  void adapt_mv_probs();
  void finish_lfm();
  void build_lfm_non420();
  void pack();
  struct TileJob { const uint8_t *data; size_t size; int tr, tc; };
  void run_tile_column(TileOut *o, const TileJob *jobs, int n_jobs);

  std::string err_;
  SideWorker adapt_worker_;          // backward adaptation of the previous frame, joined by the next parse()
  void *(*out_alloc_)(void *, size_t) = nullptr;
  void *out_priv_ = nullptr;
  uint8_t *ext_out_ = nullptr;
  size_t ext_bytes_ = 0;
  // ---- persistent decoder state
  SlotInfo slots_[VP9B_NUM_SLOTS];
  int ref_frame_map_[8];
  FrameContext frame_contexts_[4];
  FrameContext fc_;
  FrameCounts counts_;
  Segmentation seg_;
  LoopFilterParams lf_;
  int last_width_ = 0, last_height_ = 0, last_show_frame_ = 0;
  int last_frame_type_ = 0, last_intra_only_ = 0;
  int frame_type_ = 0, intra_only_ = 0, show_frame_ = 0;
  int profile_ = 0, bit_depth_ = 8, ss_x_ = 1, ss_y_ = 1;
  int width_ = 0, height_ = 0, mi_cols_ = 0, mi_rows_ = 0, sb_cols_ = 0, sb_rows_ = 0;
  int last_shown_slot_ = -1;
  uint32_t frame_index_ = 0;
  bool prev_mvs_valid_ = false;
  std::vector<MvRef> cur_mvs_, prev_mvs_;
  std::vector<uint8_t> seg_map_cur_, seg_map_last_;
  // ---- per frame
  int show_existing_ = 0, error_res_ = 0, reset_frame_context_ = 0;
  int refresh_frame_flags_ = 0, refresh_frame_context_ = 0, frame_parallel_ = 0;
  int frame_context_idx_ = 0;
  int new_slot_ = -1, show_slot_ = -1;
  int ref_slot_[3];
  int ref_sign_bias_[4];
  int allow_hp_ = 0, interp_filter_ = 0;
  int base_qindex_ = 0, y_dc_delta_q_ = 0, uv_dc_delta_q_ = 0, uv_ac_delta_q_ = 0, lossless_ = 0;
  int log2_tile_cols_ = 0, log2_tile_rows_ = 0;
  int tx_mode_ = 0, reference_mode_ = 0;
  int comp_fixed_ref_ = 0, comp_var_ref_[2] = {0, 0};
  int use_prev_frame_mvs_ = 0;
  bool do_counts_ = false;
  int16_t y_dequant_[8][2], uv_dequant_[8][2];
  uint8_t lf_lvl_[8][4][2];
  vp9b_ref_desc ref_desc_[3];
  // mode info
  std::vector<BlockInfo> blocks_;
  std::vector<uint32_t> grid_;  // mi -> index into blocks_
  // entropy contexts
  std::vector<uint8_t> above_ctx_[3];
  std::vector<uint8_t> above_seg_ctx_;
  // ---- per frame output
  int threads_ = 1, n_workers_ = 1;
  std::vector<std::unique_ptr<TileOut>> outs_;          // one per tile-column worker
  std::vector<std::unique_ptr<SideWorker>> tile_workers_;
  std::vector<uint8_t> sb_worker_;                       // superblock column -> worker that parsed it
  std::vector<vp9b_sbrange> out_sb_intra_;  // [3][sb_rows*sb_cols], start = index in the owning worker's list
  std::vector<vp9b_lfm> out_lfm_;
  std::vector<vp9b_lfm_uv> out_lfm_uv_;      // 4:2:2 / 4:4:0 frames only
  std::vector<uint8_t> packed_;
};

}  // namespace vp9b
