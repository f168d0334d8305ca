// Whole decoder = host parser + GPU reconstruction context (include/vp9_b200.h section 3).
// Mirrors decoder_decode / decode_one / decoder_get_frame (vp9_dx_iface.c:294-423): a chunk may
// be a superframe; each frame is parsed on the host, its packed buffer is submitted to the GPU
// stream, and the last frame flagged show_frame becomes the output of the chunk.
#include <string.h>
#include <nvtx3/nvToolsExt.h>

#include <new>
#include <string>

#include "vp9_b200.h"

struct vp9b_decoder {
  vp9b_parser *parser = nullptr;
  vp9gpu_ctx *gpu = nullptr;
  std::string err;
  int show_slot = -1;
  bool started = false;   // a chunk has been decoded: the frame-buffer callbacks are fixed from here on
  vp9b_get_frame_buffer_cb_fn_t fb_get = nullptr;
  vp9b_release_frame_buffer_cb_fn_t fb_release = nullptr;
  void *fb_priv = nullptr;
};

extern "C" {

vp9b_decoder *vp9b_decoder_create(int device) {
  vp9b_decoder *d = new (std::nothrow) vp9b_decoder();
  if (!d) return nullptr;
  d->parser = vp9b_parser_create();
  d->gpu = vp9gpu_create(device);
  if (!d->parser || !d->gpu) {  // no CPU fallback: without a CUDA device there is no decoder
    vp9b_parser_destroy(d->parser);
    vp9gpu_destroy(d->gpu);
    delete d;
    return nullptr;
  }
  vp9gpu_set_validation(d->gpu, 0);   // packed buffers come from vp9b_parser_* of this library
  // the parser assembles every packed frame directly in the context's page-locked upload ring
  vp9b_parser_set_output_allocator(d->parser, [](void *priv, size_t bytes) { return vp9gpu_upload_buffer((vp9gpu_ctx *)priv, bytes); }, d->gpu);
  return d;
}

void vp9b_decoder_destroy(vp9b_decoder *d) {
  if (!d) return;
  vp9gpu_destroy(d->gpu);
  vp9b_parser_destroy(d->parser);
  delete d;
}

const char *vp9b_decoder_last_error(vp9b_decoder *d) { return d ? d->err.c_str() : "null decoder"; }
vp9gpu_ctx *vp9b_decoder_gpu(vp9b_decoder *d) { return d ? d->gpu : nullptr; }
vp9b_parser *vp9b_decoder_parser(vp9b_decoder *d) { return d ? d->parser : nullptr; }

int vp9b_decoder_decode(vp9b_decoder *d, const uint8_t *data, size_t size) {
  if (!d || !data || !size) return VP9B_ERR_INVALID;
  d->show_slot = -1;
  d->started = true;
  uint32_t sizes[8];
  const int n = vp9b_parse_superframe(data, size, sizes);
  if (n < 0) {
    d->err = "corrupt superframe index";
    return n;
  }
  size_t off = 0;
  for (int i = 0; i < n; i++) {
    const void *packed = nullptr;
    size_t bytes = 0;
    int show = -1;
    nvtxRangePushA("vp9b: host parse (bool decoder, modes, tokens, masks)");
    int e = vp9b_parser_parse(d->parser, data + off, sizes[i], &packed, &bytes, &show);
    nvtxRangePop();
    if (e) {
      d->err = vp9b_parser_last_error(d->parser);
      return e;
    }
    if (packed) {
      e = vp9gpu_frame_submit(d->gpu, packed, bytes);
      if (e) {
        d->err = vp9gpu_last_error(d->gpu);
        return e;
      }
    }
    if (show >= 0) d->show_slot = show;
    off += sizes[i];
  }
  return VP9B_OK;
}

int vp9b_decoder_get_frame(vp9b_decoder *d, vp9b_image *img) {
  if (!d || !img) return VP9B_ERR_INVALID;
  if (d->show_slot < 0) return 0;
  img->slot = d->show_slot;
  int e = vp9b_parser_slot_info(d->parser, d->show_slot, &img->width, &img->height, &img->bit_depth, &img->ss_x,
                                &img->ss_y);
  if (e) return e;
  return 1;
}

int vp9b_decoder_read_frame(vp9b_decoder *d, const vp9b_image *img, uint8_t *y, uint8_t *u, uint8_t *v, int y_stride,
                            int uv_stride) {
  if (!d || !img) return VP9B_ERR_INVALID;
  const int e = vp9gpu_get_frame(d->gpu, img->slot, y, u, v, y_stride, uv_stride);
  if (e) d->err = vp9gpu_last_error(d->gpu);
  return e;
}

int vp9b_decoder_read_frame_async(vp9b_decoder *d, const vp9b_image *img, uint8_t *y, uint8_t *u, uint8_t *v,
                                  int y_stride, int uv_stride) {
  if (!d || !img) return VP9B_ERR_INVALID;
  const int e = vp9gpu_get_frame_async(d->gpu, img->slot, y, u, v, y_stride, uv_stride);
  if (e) d->err = vp9gpu_last_error(d->gpu);
  return e;
}

// vpxdec --scale (vpxdec.c:910-949): the shown frame scaled to a display size on the GPU
int vp9b_decoder_read_frame_scaled(vp9b_decoder *d, const vp9b_image *img, uint32_t dst_w, uint32_t dst_h, uint8_t *y,
                                   uint8_t *u, uint8_t *v, int y_stride, int uv_stride) {
  if (!d || !img) return VP9B_ERR_INVALID;
  int e = vp9gpu_get_frame_scaled_async(d->gpu, img->slot, dst_w, dst_h, y, u, v, y_stride, uv_stride);
  if (!e) e = vp9gpu_sync(d->gpu);
  if (e) d->err = vp9gpu_last_error(d->gpu);
  return e;
}

int vp9b_decoder_sync(vp9b_decoder *d) {
  if (!d) return VP9B_ERR_INVALID;
  const int e = vp9gpu_sync(d->gpu);
  if (e) d->err = vp9gpu_last_error(d->gpu);
  return e;
}

int vp9b_decoder_mark(vp9b_decoder *d, uint64_t *ticket) {
  if (!d) return VP9B_ERR_INVALID;
  const int e = vp9gpu_mark(d->gpu, ticket);
  if (e) d->err = vp9gpu_last_error(d->gpu);
  return e;
}

int vp9b_decoder_wait(vp9b_decoder *d, uint64_t ticket) {
  if (!d) return VP9B_ERR_INVALID;
  const int e = vp9gpu_wait(d->gpu, ticket);
  if (e) d->err = vp9gpu_last_error(d->gpu);
  return e;
}

// vpx_codec_set_frame_buffer_functions (vpx/src/vpx_decoder.c:174-191): both callbacks or an error,
// and not after the first frame (vp9_dx_iface.c:decoder_set_fb_fn: "ctx->pbi == NULL" only).
int vp9b_decoder_set_frame_buffer_functions(vp9b_decoder *d, vp9b_get_frame_buffer_cb_fn_t get,
                                            vp9b_release_frame_buffer_cb_fn_t release, void *priv) {
  if (!d || !get || !release) return VP9B_ERR_INVALID;
  if (d->started) {
    d->err = "frame buffer functions must be set before the first frame is decoded";
    return VP9B_ERR_INVALID;
  }
  d->fb_get = get;
  d->fb_release = release;
  d->fb_priv = priv;
  return VP9B_OK;
}

int vp9b_decoder_get_frame_ext(vp9b_decoder *d, vp9b_out_frame *out) {
  if (!d || !out) return VP9B_ERR_INVALID;
  if (!d->fb_get) {
    d->err = "no frame buffer functions installed";
    return VP9B_ERR_INVALID;
  }
  memset(out, 0, sizeof(*out));
  const int r = vp9b_decoder_get_frame(d, &out->img);
  if (r <= 0) return r;
  const vp9b_image &im = out->img;
  const int bps = im.bit_depth > 8 ? 2 : 1;
  size_t off[3], total = 0;
  for (int p = 0; p < 3; p++) {
    const uint32_t w = p ? (im.width + im.ss_x) >> im.ss_x : im.width, h = p ? (im.height + im.ss_y) >> im.ss_y : im.height;
    out->stride[p] = (int)((w * bps + 63) & ~63u);
    off[p] = total;
    total += (size_t)out->stride[p] * h;
  }
  if (d->fb_get(d->fb_priv, total, &out->fb) || !out->fb.data || out->fb.size < total) {
    d->err = "get_frame_buffer callback failed or returned a buffer that is too small";
    return VP9B_ERR_MEM;
  }
  for (int p = 0; p < 3; p++) out->planes[p] = out->fb.data + off[p];
  int e = vp9gpu_get_frame_async(d->gpu, im.slot, out->planes[0], out->planes[1], out->planes[2], out->stride[0], out->stride[1]);
  if (!e) e = vp9gpu_mark(d->gpu, &out->ticket);
  if (e) {
    d->err = vp9gpu_last_error(d->gpu);
    d->fb_release(d->fb_priv, &out->fb);
    return e;
  }
  return 1;
}

int vp9b_decoder_release_frame(vp9b_decoder *d, vp9b_out_frame *out) {
  if (!d || !out || !d->fb_release || !out->fb.data) return VP9B_ERR_INVALID;
  // the copy into the buffer must have finished before the application can reuse it
  if (out->ticket) vp9gpu_wait(d->gpu, out->ticket);
  const int e = d->fb_release(d->fb_priv, &out->fb);
  memset(&out->fb, 0, sizeof(out->fb));
  return e ? VP9B_ERR_INVALID : VP9B_OK;
}

}  // extern "C"
