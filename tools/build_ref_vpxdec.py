#!/usr/bin/env python3
"""Build the reference tree's own vpxdec with the B200 binding (and the restored CPU path).

Nothing of the reference is copied into this repository: the script works on a scratch copy of
<reference>/libvpx (default /root/reference/libvpx), edits it there and builds out of tree.

  1. vp9/decoder/vp9_decodeframe.c: the fork's whole-frame reconstruction passes and its broken
     multi-thread paths are dropped (by function name), integration/libvpx/vp9_b200_shim.inc is
     included in their place and vp9_decode_frame() calls b200_decode_tiles().  Headers, tile and
     partition parsing -- upstream code the fork left alone -- stay.
  2. vp9/decoder/vp9_decoder.c: vp9_get_raw_frame() fetches the frame from the GPU before handing
     it out, vp9_decoder_remove() releases the engine.
  3. The fork's CUDA-isms in otherwise portable files are neutralised (cuda_runtime.h include,
     __host__ __device__, buffers_struct.h).
  4. configure --target=generic-gnu (the tree has no x86 SIMD sources) and make vpxdec.

Output: oracle/_ref/vpxdec_ref (+ a copy of libvp9b200.so next to it so that the gpu mode finds
it).  VP9B_SHIM=cpu runs the reference C decoder, VP9B_SHIM=gpu (default) the B200 engine.
"""
import argparse
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# functions of the fork's vp9_decodeframe.c that are removed (fork-specific or broken by the fork)
DROP = {
    "inverse_transform_block_inter", "inverse_transform_block_intra", "block_sum",
    "parse_intra_block_row_mt", "predict_and_reconstruct_intra_block_row_mt", "parse_inter_block_row_mt",
    "reconstruct_inter_block_row_mt", "build_mc_border", "high_build_mc_border", "extend_and_predict",
    "dec_build_inter_predictors", "dec_build_inter_predictors_sb", "set_offsets_recon", "predict_recon_inter",
    "predict_recon_intra", "detoken_block", "intra_decode", "intra_predict_and_reconstruct", "inter_residual_sum",
    "inter_decode", "decode_block", "recon_block", "parse_block", "decode_partition", "process_partition",
    "map_write", "map_read", "lpf_map_write_check", "vp9_tile_done", "recon_tile_row",
    "parse_tile_row", "row_decode_worker_hook", "initBuf", "cloneBuf", "gotoPtrBuf", "freeBuf", "decode_tiles",
    "set_rows_after_error", "tile_worker_hook", "compare_tile_buffers", "init_mt", "decode_tiles_row_wise_mt",
    "decode_tiles_mt", "X_Fuel",
}


def split_top_level(src):
    """Yield (kind, text): 'func' chunks are whole function definitions (from the first line of the
    declarator to the closing brace), 'other' is everything else.  Good enough for libvpx's style:
    function bodies start with '{' at the end of a line at brace depth 0 and definitions start in
    column 0."""
    out, i, n = [], 0, len(src)
    depth, start_other = 0, 0
    # find candidates: a line starting in column 0 with an identifier, containing '(' and ending in '{'
    pat = re.compile(r"^(?:[A-Za-z_][^\n;{}#]*?)\b([A-Za-z_][A-Za-z0-9_]*)\s*\((?:[^;{}]*?)\)\s*\{[ \t]*$", re.M | re.S)
    pos = 0
    while True:
        m = pat.search(src, pos)
        if not m:
            break
        # the match must start at a line start at depth 0 and not inside a preprocessor line / comment
        head = src[m.start():m.end()]
        if head.lstrip().startswith(("typedef", "struct", "enum", "union", "#", "//", "/*", "else", "if", "for", "while", "switch", "do")):
            pos = m.start() + 1
            continue
        body_open = m.end() - 1 - (len(head) - len(head.rstrip())) if False else src.rfind("{", m.start(), m.end())
        d, j = 0, body_open
        while j < n:
            c = src[j]
            if c == "{":
                d += 1
            elif c == "}":
                d -= 1
                if d == 0:
                    break
            elif c == '"':
                j = src.index('"', j + 1)
            elif c == "'" :
                j = src.index("'", j + 1)
            elif src.startswith("//", j):
                j = src.index("\n", j)
            elif src.startswith("/*", j):
                j = src.index("*/", j) + 1
            j += 1
        out.append(("other", src[start_other:m.start()]))
        out.append(("func:" + m.group(1), src[m.start():j + 1]))
        start_other = j + 1
        pos = j + 1
    out.append(("other", src[start_other:]))
    return out


def patch_decodeframe(path):
    src = open(path).read()
    # the fork's non-static prototype of the static set_offsets(), and its wrap_cuda_* externs
    src = re.sub(r"\nMODE_INFO \*set_offsets\([^;]*?\);\n", "\n", src, count=1, flags=re.S)
    src = re.sub(r"\nextern int wrap_cuda_(?:inter|intra)_prediction\([^;]*?\);", "\n", src, flags=re.S)
    chunks = split_top_level(src)
    names = [k[5:] for k, _ in chunks if k.startswith("func:")]
    missing = [f for f in ("set_offsets", "read_partition", "dec_partition_plane_context", "dec_update_partition_context",
                           "get_tile_buffers", "setup_token_decoder", "dec_reset_skip_context", "vp9_decode_frame",
                           "read_uncompressed_header") if f not in names]
    if missing:
        raise SystemExit("vp9_decodeframe.c: expected functions not found: %s" % missing)
    kept, dropped = [], []
    for kind, text in chunks:
        if kind.startswith("func:") and kind[5:] in DROP:
            dropped.append(kind[5:])
            continue
        if kind == "func:vp9_decode_frame":
            # single-thread tile path -> the shim; the multi-thread dispatch above it goes away
            a = text.index("if (pbi->max_threads > 1 && tile_rows == 1")
            b = text.index("X_Fuel(pbi);", a)
            b = text.index("}", b) + 1
            text = text[:a] + "*p_data_end = b200_decode_tiles(pbi, data + first_partition_size, data_end);\n" + text[b:]
            text = '#include "vp9_b200_shim.inc"\n\n' + text
        kept.append(text)
    out = "".join(kept)
    # typedefs of function pointer types used only by dropped code are harmless; drop orphan #if blocks
    # that wrapped nothing but dropped functions (e.g. "#if CONFIG_VP9_HIGHBITDEPTH\n#endif")
    for _ in range(4):
        out = re.sub(r"#if[^\n]*\n(?:\s*//[^\n]*\n|\s*\n)*#else[^\n]*\n(?:\s*//[^\n]*\n|\s*\n)*#endif[^\n]*\n", "\n", out)
        out = re.sub(r"#if[^\n]*\n(?:\s*//[^\n]*\n|\s*\n)*#endif[^\n]*\n", "\n", out)
    open(path, "w").write(out)
    return dropped


def patch_decoder(path):
    src = open(path).read()
    # vp9_get_raw_frame: fetch the frame to show from the GPU first
    m = re.search(r"int vp9_get_raw_frame\([^)]*\)\s*\{", src)
    if not m:
        raise SystemExit("vp9_decoder.c: vp9_get_raw_frame not found")
    body_end = src.index("\n}\n", m.end())
    body = src[m.end():body_end]
    marker = "if (!cm->show_frame) return ret;"
    k = body.index(marker) + len(marker)
    hook = ("\n  {\n    RefCntBuffer *const b200_rc = (RefCntBuffer *)((char *)cm->frame_to_show - offsetof(RefCntBuffer, buf));\n"
            "    vp9_b200_fetch_frame(pbi, (int)(b200_rc - cm->buffer_pool->frame_bufs));\n  }\n")
    body = body[:k] + hook + body[k:]
    src = src[:m.end()] + body + src[body_end:]
    m = re.search(r"void vp9_decoder_remove\(VP9Decoder \*pbi\)\s*\{\n", src)
    if not m:
        raise SystemExit("vp9_decoder.c: vp9_decoder_remove not found")
    src = src[:m.end()] + "  if (pbi) vp9_b200_release(pbi);\n" + src[m.end():]
    src = src.replace('#include "vp9/decoder/vp9_decoder.h"',
                      '#include "vp9/decoder/vp9_decoder.h"\n#include <stddef.h>\nvoid vp9_b200_fetch_frame(struct VP9Decoder *pbi, int idx);\n'
                      'void vp9_b200_release(struct VP9Decoder *pbi);', 1)
    open(path, "w").write(src)


def sub_file(path, pattern, repl, count=0):
    s = open(path).read()
    s2, n = re.subn(pattern, repl, s, count=count)
    if n:
        open(path, "w").write(s2)
    return n


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    ap.add_argument("--work", default="/tmp/vp9b_refbuild")
    ap.add_argument("--out", default=os.path.join(ROOT, "oracle", "_ref"))
    ap.add_argument("--jobs", type=int, default=os.cpu_count() or 4)
    ap.add_argument("--highbitdepth", action="store_true", help="--enable-vp9-highbitdepth (the fork's HBD transforms do not build)")
    args = ap.parse_args()
    ref = os.path.join(args.reference, "libvpx")
    if not os.path.isdir(ref):
        raise SystemExit("reference tree %s not present (this recipe only runs where /root/reference exists)" % ref)
    work = args.work
    shutil.rmtree(work, ignore_errors=True)
    os.makedirs(work)
    src = os.path.join(work, "libvpx")
    shutil.copytree(ref, src, ignore=shutil.ignore_patterns("build_debug", "third_party/googletest"))
    subprocess.check_call(["chmod", "-R", "u+w", src])
    for f in ["configure"] + [os.path.join("build", "make", x) for x in os.listdir(os.path.join(src, "build", "make"))]:
        os.chmod(os.path.join(src, f), 0o755)
    # --- scratch-copy edits
    dropped = patch_decodeframe(os.path.join(src, "vp9", "decoder", "vp9_decodeframe.c"))
    patch_decoder(os.path.join(src, "vp9", "decoder", "vp9_decoder.c"))
    sub_file(os.path.join(src, "vp9", "decoder", "vp9_decodeframe.h"), r'#include "buffers_struct.h"\n', "")
    sub_file(os.path.join(src, "vpx_dsp", "vpx_convolve.h"), r'#include "cuda_runtime.h"\n', "")
    sub_file(os.path.join(src, "vp9", "common", "vp9_scale.c"), r"__host__ __device__ ", "")
    shutil.copy(os.path.join(ROOT, "integration", "libvpx", "vp9_b200_shim.inc"), os.path.join(src, "vp9", "decoder"))
    shutil.copy(os.path.join(ROOT, "include", "vp9_b200.h"), os.path.join(src, "vp9", "decoder"))
    print("vp9_decodeframe.c: dropped %d fork functions" % len(dropped))
    # --- configure + make
    bld = os.path.join(work, "build")
    os.makedirs(bld)
    cfg = ["../libvpx/configure", "--target=generic-gnu", "--disable-vp8", "--disable-vp9-encoder", "--disable-unit-tests",
           "--disable-docs", "--disable-webm-io", "--disable-libyuv", "--enable-pic", "--extra-cflags=-O3 -w"]
    if args.highbitdepth:
        cfg.append("--enable-vp9-highbitdepth")
    subprocess.check_call(cfg, cwd=bld, stdout=subprocess.DEVNULL)
    env = dict(os.environ)
    r = subprocess.run(["make", "-j%d" % args.jobs, "LDFLAGS=-ldl"], cwd=bld, env=env,
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode:
        sys.stdout.write(r.stdout[-6000:])
        raise SystemExit("make vpxdec failed")
    os.makedirs(args.out, exist_ok=True)
    shutil.copy(os.path.join(bld, "vpxdec"), os.path.join(args.out, "vpxdec_ref"))
    lib = os.path.join(ROOT, "cuda-vp9_b200", "libvp9b200.so")
    print("built %s" % os.path.join(args.out, "vpxdec_ref"))
    print("gpu mode loads %s (VP9B_LIB) at run time" % lib)


if __name__ == "__main__":
    main()
