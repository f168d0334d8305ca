/*
 * vp9_b200.h -- C ABI of the B200-native VP9 decode-reconstruction engine.
 *
 * This is the drop-in boundary that takes the role of the reference's
 *     extern "C" wrap_cuda_inter_prediction / wrap_cuda_intra_prediction
 *     (vpx-master/cuda_extern_wrap.cpp:5-17, declared vp9_decodeframe.c:2299-2302,
 *      called vp9_decodeframe.c:2546,2564)
 * but with plain pointers and sizes only -- no libvpx types cross it.  Three layers:
 *
 *   1. vp9gpu_*   reconstruction context: device frame pool + "one packed buffer per frame"
 *                 (SURVEY.md section 8(b)).  A libvpx host calls this once per frame from
 *                 decode_tiles() after parsing all tiles (INTEGRATION.md shows the stub).
 *   2. vp9b_parser_*   host bitstream parser that produces the packed buffer (role of
 *                 vp9_read_mode_info / vp9_decode_block_tokens / vp9_build_mask,
 *                 vp9_decodemv.c:811, vp9_detokenize.c:274, vp9_loopfilter.c:1528).
 *   3. vp9b_decoder_*  whole decoder = parser + gpu ctx; mirrors vpx_codec_decode /
 *                 vpx_codec_get_frame (vpx/src/vpx_decoder.c:104,  vp9_dx_iface.c:331,399).
 *
 * All functions return 0 on success and a negative VP9B_ERR_* otherwise; nothing aborts.
 */
#ifndef VP9_B200_H_
#define VP9_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VP9B_MAGIC 0x39425056u /* "VPB9" */
#define VP9B_NUM_SLOTS 12      /* device frame pool: 8 refs + current + held output + spare */

enum {
  VP9B_OK = 0,
  VP9B_ERR_INVALID = -1,     /* bad argument / corrupt packed buffer            */
  VP9B_ERR_UNSUP = -2,       /* unsupported bitstream feature (e.g. 4:2:2 LPF)  */
  VP9B_ERR_CORRUPT = -3,     /* corrupt bitstream                                */
  VP9B_ERR_CUDA = -4,        /* CUDA runtime error (see *_last_error)           */
  VP9B_ERR_MEM = -5,
  VP9B_ERR_NOGPU = -6        /* library built without / run without a CUDA device */
};

/* ------------------------------------------------------------------ packed frame buffer --- */
/* Section ids of the packed per-frame buffer ("one packed SoA buffer", BASELINE north_star). */
enum {
  VP9B_SEC_BLOCKS = 0,   /* vp9b_block[]      every prediction block in decode order          */
  VP9B_SEC_SUBMV = 1,    /* vp9b_submv[]      MVs of sub-8x8 inter blocks                     */
  VP9B_SEC_TXB_INTER = 2,/* vp9b_txb[]        inter residual transform blocks, eob > 0 only,
                                              grouped by transform size (hdr.n_txb_inter_size) */
  VP9B_SEC_TXB_INTRA_Y = 3, /* vp9b_txb[]     intra transform blocks per plane, decode order  */
  VP9B_SEC_TXB_INTRA_U = 4,
  VP9B_SEC_TXB_INTRA_V = 5,
  VP9B_SEC_SB_INTRA = 6, /* vp9b_sbrange[3][sb_rows*sb_cols]  per-plane range into TXB_INTRA_* */
  VP9B_SEC_COEF = 7,     /* int16 (8-bit) or int32 (HBD) dequantised coefs, scan order, eob each */
  VP9B_SEC_LFM = 8,      /* vp9b_lfm[sb_rows*sb_cols]  final (adjusted) loop-filter masks     */
  VP9B_SEC_LFM_UV = 9,   /* vp9b_lfm_uv[sb_rows*sb_cols]  chroma masks of 4:2:2 / 4:4:0 frames only */
  VP9B_SEC_COUNT = 12
};

typedef struct vp9b_section {
  uint32_t offset; /* bytes from start of the packed buffer, 16-byte aligned */
  uint32_t count;  /* number of records (COEF: number of coefficients)       */
} vp9b_section;

typedef struct vp9b_ref_desc { /* role of struct scale_factors + RefBuffer (vp9_scale.h:25-38) */
  int32_t slot;                /* device frame slot, -1 if unused               */
  uint32_t width, height;      /* reference frame crop size (luma)              */
  int32_t x_scale_fp, y_scale_fp; /* 14-bit fixed point (vp9_scale.c:37-43)     */
  int32_t x_step_q4, y_step_q4;   /* 16 == unscaled                              */
} vp9b_ref_desc;

typedef struct vp9b_frame_hdr {
  uint32_t magic;
  uint32_t total_bytes;
  uint32_t width, height;        /* frame crop size in luma samples */
  uint32_t mi_cols, mi_rows;     /* 8x8 units; aligned size = 8*mi   */
  uint32_t sb_cols, sb_rows;     /* 64x64 units                      */
  uint8_t bit_depth;             /* 8, 10, 12                        */
  uint8_t ss_x, ss_y;
  uint8_t lossless;              /* all transforms are 4x4 WHT       */
  uint8_t intra_only;            /* key frame or intra-only frame: no inter blocks */
  uint8_t show_frame;
  uint8_t lf_level;              /* 0 => loop filter skipped          */
  uint8_t lf_sharpness;
  uint8_t coef_bytes;            /* 2 or 4                            */
  uint8_t reserved[3];
  int32_t dst_slot;              /* device frame slot reconstructed into */
  vp9b_ref_desc ref[3];          /* LAST, GOLDEN, ALTREF              */
  vp9b_section sec[VP9B_SEC_COUNT];
  uint8_t lf_lim[64];            /* per level: inside limit  (vp9_loopfilter.c:212-231) */
  uint8_t lf_mblim[64];          /* per level: edge limit; hev_thr = level >> 4 (:249)  */
  uint32_t frame_index;          /* decode-order counter, informational */
  uint32_t n_inter_px[3];        /* predicted samples per plane (roofline accounting)   */
  uint32_t n_compound_px[3];
  uint32_t n_txb_inter_size[4];  /* TXB_INTER is ordered by transform size: counts of 4x4, 8x8, 16x16, 32x32 */
  uint32_t reserved2[4];
} vp9b_frame_hdr;

typedef struct vp9b_block { /* 16 bytes. role of MODE_INFO fields in vp9_blockd.h:74-97 */
  uint16_t mi_col, mi_row;
  uint8_t bsize;      /* BLOCK_4X4..BLOCK_64X64 (0..12) */
  uint8_t ref0;       /* 0 intra, 1 LAST, 2 GOLDEN, 3 ALTREF */
  uint8_t ref1;       /* 0 none, else second reference (compound) */
  uint8_t filter;     /* 0 regular, 1 smooth, 2 sharp, 3 bilinear (vp9_filter.c:79-82) */
  int16_t mv[2][2];   /* [ref][row,col] in 1/8 pel; for bsize < 8X8 inter blocks
                         mv[0] holds the uint32 index into VP9B_SEC_SUBMV instead */
} vp9b_block;

typedef struct vp9b_submv { /* 40 bytes */
  int16_t luma[4][2][2];   /* [block 0..3][ref][row,col]  (bmi[], vp9_decodemv.c:761-768) */
  int16_t chroma[2][2];    /* [ref][row,col]: 4:2:0 average, vp9_reconinter.c:61-87,112-124 */
} vp9b_submv;

typedef struct vp9b_txb { /* 12 bytes */
  uint32_t coef_off;  /* index of first coefficient in VP9B_SEC_COEF            */
  uint16_t x4, y4;    /* position inside the plane in units of 4 samples        */
  uint16_t eob;       /* number of coefficients stored (scan order); 0 = none   */
  uint8_t tx;         /* tx_size (bits 0-1) | tx_type (2-3) | plane (4-5)       */
  uint8_t aux;        /* intra: mode (0-3) | have_top (4) | have_left (5) | have_right (6) */
} vp9b_txb;

typedef struct vp9b_sbrange {
  uint32_t start, count;
} vp9b_sbrange;

typedef struct vp9b_lfm { /* 160 bytes. role of LOOP_FILTER_MASK, vp9_loopfilter.h:60-68 */
  uint64_t left_y[4];
  uint64_t above_y[4];
  uint64_t int_4x4_y;
  uint16_t left_uv[4];
  uint16_t above_uv[4];
  uint16_t int_4x4_uv;
  uint16_t pad[3];
  uint8_t lfl_y[64];
} vp9b_lfm;

/* Chroma loop-filter masks of a superblock when ss_x != ss_y (4:2:2, 4:4:0): what
 * vp9_filter_block_plane_non420 (vp9_loopfilter.c:1080-1239) derives from the mode info, frame-edge
 * rules already applied.  Index r = mi row inside the superblock (only every (1 << ss_y)-th entry is
 * used), bit = mi column >> ss_x.  v*: vertical edges of the 8-sample-row band, h*: its top edge
 * (and inner edge), by filter size 16 / 8 / 4 / inner 4x4.  (In those frames the uv fields of
 * vp9b_lfm are zero.) */
typedef struct vp9b_lfm_uv { /* 128 bytes */
  uint8_t v16[8], v8[8], v4[8], v4i[8];
  uint8_t h16[8], h8[8], h4[8], h4i[8];
  uint8_t lfl[64]; /* filter level at [(r << 3) + (c >> ss_x)] */
} vp9b_lfm_uv;

/* ------------------------------------------------------------------ 1. reconstruction ctx --- */
typedef struct vp9gpu_ctx vp9gpu_ctx;

/* device < 0: current device.  Returns NULL on failure (no CUDA device).  Side effect shared by every
 * vp9gpu_* / vp9b_decoder_* call: the calling thread's current CUDA device is set to the context's
 * device (cudaSetDevice) and left there. */
vp9gpu_ctx *vp9gpu_create(int device);
void vp9gpu_destroy(vp9gpu_ctx *ctx);
const char *vp9gpu_last_error(vp9gpu_ctx *ctx);

/* Reconstruct one frame from a packed buffer in HOST memory (copied H2D asynchronously from an
 * internal pinned ring) -- enqueues upload + all kernels on the ctx stream and returns. */
int vp9gpu_frame_submit(vp9gpu_ctx *ctx, const void *packed, size_t bytes);
/* The same without the staging copy: vp9gpu_upload_buffer() hands out the next page-locked upload
 * buffer of the ctx (at least `bytes` large; valid until the matching submit), the caller builds the
 * packed frame in it, vp9gpu_frame_submit_staged() uploads and reconstructs it.  One buffer can be
 * outstanding per ctx; asking again before submitting returns the same buffer (grown if needed,
 * contents not preserved). */
void *vp9gpu_upload_buffer(vp9gpu_ctx *ctx, size_t bytes);
int vp9gpu_frame_submit_staged(vp9gpu_ctx *ctx, const void *packed, size_t bytes);
/* Same, but the packed buffer already lives in DEVICE memory (recon-only replay, bench). */
int vp9gpu_frame_submit_device(vp9gpu_ctx *ctx, const void *d_packed, const vp9b_frame_hdr *hdr);
/* The same for n device-resident frames in decode order with one call (replay / transcoding
 * front ends that queue a whole GOP): stops at the first error. */
int vp9gpu_frame_submit_device_n(vp9gpu_ctx *ctx, const void *const *d_packed,
                                 const vp9b_frame_hdr *const *hdr, int n);
/* One frame of each of n DIFFERENT contexts (same device) through one kernel launch per stage: the
 * hardware has 32 work queues, so serving more than 32 concurrent streams per GPU means batching
 * several streams per launch.  Up to VP9B_MAX_BATCH frames of the same sample type (8-bit / high bit
 * depth) and chroma format share the launches (others are submitted one by one); they run on the
 * stream of ctxs[0], the other contexts' streams are joined before and after.  Frame sizes may differ. */
#define VP9B_MAX_BATCH 32
int vp9gpu_frame_submit_device_multi(vp9gpu_ctx *const *ctxs, const void *const *d_packed,
                                     const vp9b_frame_hdr *const *hdr, int n);
/* Upload a packed buffer to a ctx-owned device allocation for later submit_device calls. */
int vp9gpu_upload_packed(vp9gpu_ctx *ctx, const void *packed, size_t bytes, void **d_packed);
int vp9gpu_free_packed(vp9gpu_ctx *ctx, void *d_packed);
/* Wait for everything enqueued on the ctx stream. */
int vp9gpu_sync(vp9gpu_ctx *ctx);
/* D2H copy of the cropped planes of a slot (8-bit: 1 byte/sample, HBD: 2, little endian).
 * Strides in bytes.  Synchronous.  (role of vp9_get_raw_frame, vp9_decoder.c:494) */
int vp9gpu_get_frame(vp9gpu_ctx *ctx, int slot, uint8_t *y, uint8_t *u, uint8_t *v,
                     int y_stride, int uv_stride);
/* Same without the final synchronisation: the copy is ordered after the reconstruction of the
 * slot on the ctx stream; call vp9gpu_sync() (or reuse the ctx) before reading the buffers.
 * Use pinned destination buffers (vp9gpu_alloc_host) so that the copy overlaps later frames. */
int vp9gpu_get_frame_async(vp9gpu_ctx *ctx, int slot, uint8_t *y, uint8_t *u, uint8_t *v, int y_stride,
                           int uv_stride);
/* High-bit-depth output at 8 bits: samples are shifted down by (bit_depth - 8) on the GPU
 * (vpx_img_downshift, vpxdec --output-bit-depth=8, vpxdec.c:950-978) and copied as 1 byte per
 * sample, i.e. half the D2H bytes.  Asynchronous like vp9gpu_get_frame_async; for an 8-bit slot it
 * is identical to it. */
int vp9gpu_get_frame_8bit_async(vp9gpu_ctx *ctx, int slot, uint8_t *y, uint8_t *u, uint8_t *v, int y_stride,
                                int uv_stride);
/* Output scaled to a display size (vpxdec --scale: vpx_image_scale -> libyuv I420Scale[_16] with
 * kFilterBox, vpxdec.c:910-949), done on the GPU before the copy; dst_w x dst_h luma samples, 4:2:0
 * only (as in vpxdec).  Bit-exact with libyuv's C paths for every change within a factor of two that
 * libyuv handles with its bilinear code (any resize of a VP9 stream scaled back to its first size);
 * the ratios it gives to box kernels (exactly 3/4, 1/2, 3/8, below 1/2, height-only) return
 * VP9B_ERR_UNSUP.  Asynchronous like vp9gpu_get_frame_async; same size = plain copy. */
int vp9gpu_get_frame_scaled_async(vp9gpu_ctx *ctx, int slot, uint32_t dst_w, uint32_t dst_h, uint8_t *y, uint8_t *u,
                                  uint8_t *v, int y_stride, int uv_stride);
/* Tickets: vp9gpu_mark() records a point on the ctx stream (e.g. right after a get_frame_async),
 * vp9gpu_wait() blocks until that point has been reached.  Lets a host thread wait for the frame
 * it handed out N frames ago instead of draining the stream.  Up to 64 tickets may be pending. */
int vp9gpu_mark(vp9gpu_ctx *ctx, uint64_t *ticket);
int vp9gpu_wait(vp9gpu_ctx *ctx, uint64_t ticket);
/* Page-locked host memory for output frames / packed buffers. */
void *vp9gpu_alloc_host(size_t bytes);
void vp9gpu_free_host(void *p);
/* H2D copy into a slot (VP8_SET_REFERENCE, vp9_decoder.c:258-331, and tests). */
int vp9gpu_set_frame(vp9gpu_ctx *ctx, int slot, uint32_t width, uint32_t height, int bit_depth,
                     int ss_x, int ss_y, const uint8_t *y, const uint8_t *u, const uint8_t *v,
                     int y_stride, int uv_stride);
/* Record-level validation of packed buffers handed over in host memory (vp9gpu_frame_submit,
 * vp9gpu_upload_packed): every block / transform / range record is checked against the frame
 * geometry and the section sizes before anything reaches a kernel.  On by default -- a host-side
 * packer is a foreign producer; vp9b_decoder_* turns it off for buffers of its own parser. */
int vp9gpu_set_validation(vp9gpu_ctx *ctx, int enable);
/* Debug/test hook: stop after a stage (0 = all; 1 inter pred; 2 + inter residual; 3 + intra;
 * 4 + loop filter). */
int vp9gpu_set_stop_stage(vp9gpu_ctx *ctx, int stage);
/* CUDA stream handle (cudaStream_t) used by the ctx, for event timing by the caller. */
void *vp9gpu_stream(vp9gpu_ctx *ctx);
/* Number of kernel launches issued since creation. */
uint64_t vp9gpu_launch_count(vp9gpu_ctx *ctx);
/* Accumulated device milliseconds per stage since last reset (needs profiling enabled on the
 * context whose stream the launches run on -- ctxs[0] of a batch): 0 expand, 1 inter pred,
 * 2 inter residual, 3 intra residual (inverse transforms of intra blocks), 4 intra wavefront,
 * 5 loop filter. */
int vp9gpu_set_profiling(vp9gpu_ctx *ctx, int enable);
int vp9gpu_get_stage_ms(vp9gpu_ctx *ctx, float ms[8], int reset);

/* ------------------------------------------------------------------ 1b. frame packer ------- */
/* Builds the packed buffer record by record, for a host decoder that keeps its own bitstream parser
 * (libvpx: called from decode_block(); integration/libvpx/vp9_b200_shim.inc is that binding).  Role
 * of createBuffers / createBuffersTr (vpx-master/inter_cuda_kernel.cu:897-1028,
 * intra_cuda_kernel.cu:1252-1304) without any libvpx type crossing the boundary. */
typedef struct vp9b_packer vp9b_packer;

vp9b_packer *vp9b_packer_create(void);
void vp9b_packer_destroy(vp9b_packer *p);
const char *vp9b_packer_last_error(vp9b_packer *p);
/* Start a frame.  Taken from *hdr: width, height, bit_depth, ss_x, ss_y, lossless, intra_only,
 * show_frame, lf_level, lf_sharpness, dst_slot, ref[3], lf_lim, lf_mblim, frame_index.  Geometry
 * (mi / sb counts), coef_bytes, the section table and the counters are filled in by the packer. */
int vp9b_packer_begin(vp9b_packer *p, const vp9b_frame_hdr *hdr);
/* One prediction block, in decode order.  sub_mv: the four 4x4 motion vectors [block][ref][row,col]
 * of an inter block below 8x8 (bmi[], vp9_blockd.h:74-97), else NULL; b->mv is ignored for those. */
int vp9b_packer_add_block(vp9b_packer *p, const vp9b_block *b, const int16_t sub_mv[4][2][2]);
/* One transform block.  x4, y4: position in the plane in units of 4 samples.  inter != 0: residual
 * of an inter block (dropped when eob == 0); inter == 0: intra block (always recorded) with its
 * prediction mode (PREDICTION_MODE 0..9) and the edge availability vp9_predict_intra_block derives
 * (vp9_reconintra.c:404-424).  dqcoeff: dequantised coefficients at their raster positions, as
 * vp9_decode_block_tokens leaves them (vp9_detokenize.c:217-248), coef_bytes 2 (int16) or 4
 * (int32); scan: the scan table used for the block (>= eob entries). */
int vp9b_packer_add_txb(vp9b_packer *p, int plane, int inter, int x4, int y4, int tx_size, int tx_type, int mode,
                        int have_top, int have_left, int have_right, const void *dqcoeff, int coef_bytes,
                        const int16_t *scan, int eob);
/* Final loop-filter mask of one superblock (after vp9_adjust_mask, vp9_loopfilter.c:766-878). */
int vp9b_packer_set_lfm(vp9b_packer *p, int sb_row, int sb_col, const vp9b_lfm *m);
/* 4:2:2 / 4:4:0 frames only: the chroma masks of one superblock (what vp9_filter_block_plane_non420
 * derives from the mode info, vp9_loopfilter.c:1080-1178, with its frame-edge rules applied); the uv
 * fields of the vp9b_lfm passed to vp9b_packer_set_lfm must then be zero.  Ignored for other frames. */
int vp9b_packer_set_lfm_uv(vp9b_packer *p, int sb_row, int sb_col, const vp9b_lfm_uv *m);
/* Close the frame: *packed (packer-owned, valid until the next begin) is ready for
 * vp9gpu_frame_submit. */
int vp9b_packer_finish(vp9b_packer *p, const void **packed, size_t *bytes);

/* ------------------------------------------------------------------ 2. host parser --------- */
typedef struct vp9b_parser vp9b_parser;

vp9b_parser *vp9b_parser_create(void);
void vp9b_parser_destroy(vp9b_parser *p);
const char *vp9b_parser_last_error(vp9b_parser *p);
/* Parse ONE frame (not a superframe) of compressed data.  On success *packed points to a
 * parser-owned buffer (valid until the next call) holding vp9b_frame_hdr + sections, or NULL
 * with *bytes == 0 for a show_existing_frame.  *show_slot is the frame slot to output
 * (-1: nothing to show).  (role of vp9_receive_compressed_data, vp9_decoder.c:407) */
int vp9b_parser_parse(vp9b_parser *p, const uint8_t *data, size_t size, const void **packed,
                      size_t *bytes, int *show_slot);
/* Tile-column workers per frame (vp9_decodeframe.c:2919-3047 decode_tiles_mt): 1 (default, or the
 * environment variable VP9B_PARSE_THREADS) parses every tile on the calling thread; n > 1 spreads the
 * tile columns over n threads owned by the parser.  Worth it when there are fewer streams than host
 * cores; the packed buffer is equivalent either way (records may be ordered differently). */
int vp9b_parser_set_threads(vp9b_parser *p, int threads);
/* Where the packed buffer is assembled: alloc(priv, bytes) returns the destination, e.g. the
 * page-locked upload buffer of a reconstruction context (vp9gpu_upload_buffer) so that the frame is
 * packed straight into the memory the H2D copy reads.  alloc == NULL (default) or a NULL return:
 * a parser-owned buffer. */
int vp9b_parser_set_output_allocator(vp9b_parser *p, void *(*alloc)(void *priv, size_t bytes), void *priv);
/* Geometry of a slot as the parser last set it up. */
int vp9b_parser_slot_info(vp9b_parser *p, int slot, uint32_t *width, uint32_t *height,
                          int *bit_depth, int *ss_x, int *ss_y);
/* Split a superframe (vp9_parse_superframe_index, vp9_dx_iface.c:354): fills up to 8 sizes,
 * returns the number of frames (>=1) or a negative error. */
int vp9b_parse_superframe(const uint8_t *data, size_t size, uint32_t sizes[8]);

/* ------------------------------------------------------------------ 3. whole decoder ------- */
typedef struct vp9b_decoder vp9b_decoder;

typedef struct vp9b_image {
  uint32_t width, height;
  int bit_depth, ss_x, ss_y;
  int slot;
} vp9b_image;

vp9b_decoder *vp9b_decoder_create(int device);
void vp9b_decoder_destroy(vp9b_decoder *d);
const char *vp9b_decoder_last_error(vp9b_decoder *d);
/* Decode one chunk (IVF frame payload; superframes handled).  (vpx_codec_decode) */
int vp9b_decoder_decode(vp9b_decoder *d, const uint8_t *data, size_t size);
/* After decode: 1 if a frame is ready to show (fills *img), 0 if none.  (vpx_codec_get_frame) */
int vp9b_decoder_get_frame(vp9b_decoder *d, vp9b_image *img);
/* Copy the cropped planes of the shown frame to host memory (strides in bytes). */
int vp9b_decoder_read_frame(vp9b_decoder *d, const vp9b_image *img, uint8_t *y, uint8_t *u,
                            uint8_t *v, int y_stride, int uv_stride);
/* Asynchronous variant (see vp9gpu_get_frame_async) and the matching wait. */
int vp9b_decoder_read_frame_async(vp9b_decoder *d, const vp9b_image *img, uint8_t *y, uint8_t *u, uint8_t *v,
                                  int y_stride, int uv_stride);
/* The shown frame scaled to dst_w x dst_h (vpxdec --scale; see vp9gpu_get_frame_scaled_async);
 * synchronous. */
int vp9b_decoder_read_frame_scaled(vp9b_decoder *d, const vp9b_image *img, uint32_t dst_w, uint32_t dst_h, uint8_t *y,
                                   uint8_t *u, uint8_t *v, int y_stride, int uv_stride);
int vp9b_decoder_sync(vp9b_decoder *d);
/* Ticket of everything enqueued so far / wait for a ticket (see vp9gpu_mark, vp9gpu_wait). */
int vp9b_decoder_mark(vp9b_decoder *d, uint64_t *ticket);
int vp9b_decoder_wait(vp9b_decoder *d, uint64_t ticket);
vp9gpu_ctx *vp9b_decoder_gpu(vp9b_decoder *d);
vp9b_parser *vp9b_decoder_parser(vp9b_decoder *d);

/* External output buffers (role of vpx_codec_set_frame_buffer_functions, vpx/src/vpx_decoder.c:174,
 * vpx/vpx_frame_buffer.h:35-80).  Reference frames live on the GPU, so what the application lends
 * are the buffers the SHOWN frames are delivered in -- ideally page-locked (vp9gpu_alloc_host), so
 * that the device-to-host copy overlaps the following frames.  get() is called once per shown frame
 * with the minimum size, release() when the application returns the frame.  Like libvpx: both
 * callbacks are required, and they cannot be installed once decoding has started. */
typedef struct vp9b_frame_buffer {
  uint8_t *data; /* set by get(): at least min_size bytes */
  size_t size;
  void *priv;    /* for the application's bookkeeping; handed back to release() */
} vp9b_frame_buffer;
typedef int (*vp9b_get_frame_buffer_cb_fn_t)(void *priv, size_t min_size, vp9b_frame_buffer *fb);
typedef int (*vp9b_release_frame_buffer_cb_fn_t)(void *priv, vp9b_frame_buffer *fb);
int vp9b_decoder_set_frame_buffer_functions(vp9b_decoder *d, vp9b_get_frame_buffer_cb_fn_t get,
                                            vp9b_release_frame_buffer_cb_fn_t release, void *priv);
typedef struct vp9b_out_frame {
  vp9b_image img;
  uint8_t *planes[3]; /* inside fb.data, rows of stride[] bytes (64-byte multiples) */
  int stride[3];
  vp9b_frame_buffer fb;
  uint64_t ticket;    /* the samples are in host memory once vp9b_decoder_wait(d, ticket) returns */
} vp9b_out_frame;
/* After decode: 1 if a frame is shown -- a buffer is taken from get(), the copy of the frame into it
 * is enqueued and *out is filled; 0 if there is nothing to show.  (vpx_codec_get_frame with external
 * frame buffers.)  The frame is returned with vp9b_decoder_release_frame (-> release()). */
int vp9b_decoder_get_frame_ext(vp9b_decoder *d, vp9b_out_frame *out);
int vp9b_decoder_release_frame(vp9b_decoder *d, vp9b_out_frame *out);

#ifdef __cplusplus
}
#endif

/* The packed format is shared with foreign producers (a libvpx-side packer, the ctypes mirror in
 * cuda-vp9_b200/capi.py): pin the record sizes at compile time. */
#if defined(__cplusplus)
#define VP9B_STATIC_ASSERT(c, m) static_assert(c, m)
#elif defined(__STDC_VERSION__) && __STDC_VERSION__ >= 201112L
#define VP9B_STATIC_ASSERT(c, m) _Static_assert(c, m)
#else
#define VP9B_STATIC_ASSERT(c, m)
#endif
VP9B_STATIC_ASSERT(sizeof(vp9b_block) == 16, "vp9b_block is 16 bytes");
VP9B_STATIC_ASSERT(sizeof(vp9b_submv) == 40, "vp9b_submv is 40 bytes");
VP9B_STATIC_ASSERT(sizeof(vp9b_txb) == 12, "vp9b_txb is 12 bytes");
VP9B_STATIC_ASSERT(sizeof(vp9b_sbrange) == 8, "vp9b_sbrange is 8 bytes");
VP9B_STATIC_ASSERT(sizeof(vp9b_lfm) == 160, "vp9b_lfm is 160 bytes");
VP9B_STATIC_ASSERT(sizeof(vp9b_lfm_uv) == 128, "vp9b_lfm_uv is 128 bytes");
VP9B_STATIC_ASSERT(sizeof(vp9b_frame_hdr) == 416, "vp9b_frame_hdr is 416 bytes");
#endif /* VP9_B200_H_ */
