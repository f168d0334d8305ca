/* TEST INFRASTRUCTURE -- the vp9gpu_* entry points the libvpx shim binds, implemented on the CPU
 * oracle (oracle/vp9_recon_ref.c).  Lets the reference tree's vpxdec run its "gpu" mode without a
 * GPU, so that the packing done by integration/libvpx/vp9_b200_shim.inc (libvpx's own parser ->
 * vp9b_packer_*) is checked here against the golden MD5s.  Never part of the product. */
#include "../../oracle/vp9_recon_ref.c"

struct vp9gpu_ctx { vp9ref_ctx *ref; const char *err; };

vp9gpu_ctx *vp9gpu_create(int device) {
  (void)device;
  vp9gpu_ctx *c = (vp9gpu_ctx *)calloc(1, sizeof(*c));
  c->ref = vp9ref_create();
  c->err = "";
  return c;
}
void vp9gpu_destroy(vp9gpu_ctx *c) { if (c) { vp9ref_destroy(c->ref); free(c); } }
const char *vp9gpu_last_error(vp9gpu_ctx *c) { return c ? c->err : "null"; }
int vp9gpu_frame_submit(vp9gpu_ctx *c, const void *packed, size_t bytes) {
  const int e = vp9ref_recon_frame(c->ref, packed, bytes, 0);
  if (e) c->err = "oracle rejected the packed frame";
  return e;
}
int vp9gpu_get_frame(vp9gpu_ctx *c, int slot, uint8_t *y, uint8_t *u, uint8_t *v, int ys, int uvs) {
  return vp9ref_get_frame(c->ref, slot, y, u, v, ys, uvs);
}
