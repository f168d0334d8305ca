#!/usr/bin/env python3
"""Generate the synthetic test clips under tests/golden/clips/ and their per-frame MD5 lists
(tests/golden/md5/<clip>.md5, same format as the reference's Sony.md5 / netflix.md5) with the
libvpx v1.15.2 encoder + decoder that ships in this image.

The clips stand in for the reference's Bravia.1440.ivf / FoodMarket2.ivf, which are absent from
the reference checkout (.MISSING_LARGE_BLOBS:1-2).  Usage:
    python tools/make_clips.py            # all small parity clips
    python tools/make_clips.py bench      # larger bench clips (1440p / 4K)
"""
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle import synth, vpx_oracle as vo  # noqa: E402

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")

# name -> encode parameters.  Each clip targets features of the hot path.
PARITY = {
    # odd frame sizes (edge emulation in MC, intra edge extension, lpf mask trimming), real-time
    # mode so that the loop filter level is non-zero on key frames too
    "odd_rt": dict(w=200, h=136, n=8, deadline=1, cpu_used=6, bitrate=200),
    "odd2": dict(w=130, h=98, n=6, cpu_used=2, cq_level=35),
    # 4 tile columns x 2 tile rows (intra have_left is tile relative)
    "tiles_rt": dict(w=1280, h=384, n=4, deadline=1, cpu_used=7, tile_cols_log2=2, tile_rows_log2=1, bitrate=700),
    # two-pass, hidden alt-ref frame (superframe), compound prediction
    "comp": dict(w=352, h=288, n=40, cpu_used=1, lag=16, auto_alt_ref=1, two_pass=True, bitrate=400, end_usage=0,
                 arnr_maxframes=7, arnr_strength=5, detail=0.02, temporal_noise=4.0, noise_corner=False),
    # same kind of stream with the loop-filter level forced on (bitstream edit, oracle/synth.py)
    "comp_lf": dict(w=352, h=288, n=40, cpu_used=1, lag=16, auto_alt_ref=1, two_pass=True, bitrate=400, end_usage=0,
                    arnr_maxframes=7, arnr_strength=5, detail=0.02, temporal_noise=4.0, noise_corner=False,
                    force_lf=True),
    # sub-8x8 partitions (4x4/4x8/8x4 inter + intra)
    "sub8": dict(w=352, h=288, n=6, cpu_used=0, bitrate=2000, end_usage=0, min_q=4, max_q=40),
    # profile 2, 10 and 12 bit (highbd convolve / idct / lpf)
    "hbd10": dict(w=320, h=240, n=6, bit_depth=10, cpu_used=3, cq_level=30),
    "hbd12": dict(w=208, h=160, n=5, bit_depth=12, cpu_used=4, cq_level=40, deadline=1),
    # lossless: 4x4 WHT path
    "lossless": dict(w=176, h=144, n=3, lossless=1, cpu_used=4),
    # segmentation (per-segment q and lf level)
    "aq": dict(w=352, h=288, n=8, aq_mode=3, cpu_used=5, deadline=1, bitrate=300),
    "errres": dict(w=352, h=288, n=6, error_resilient=1, cpu_used=5, deadline=1, bitrate=300),
    "fpar": dict(w=352, h=288, n=6, frame_parallel=1, cpu_used=3, cq_level=30),
    # reference scaling: the coded size changes mid-stream (VP8E_SET_SCALEMODE), so inter
    # prediction reads references of a different size (x_step_q4 != 16)
    "scaled": dict(w=352, h=288, n=12, cpu_used=3, cq_level=30, scale_at={3: (3, 3), 6: (1, 2), 9: (0, 0)}),
    # profile 1 / 3: 4:4:4 sampling (vp9_filter_block_plane_non420, full-size chroma MC / intra)
    "p444": dict(w=176, h=144, n=5, chroma444=True, deadline=1, cpu_used=5, bitrate=300),
    "p444_hbd": dict(w=144, h=112, n=4, chroma444=True, bit_depth=10, deadline=1, cpu_used=5, bitrate=300),
    # profile 1 / 3 with 4:2:2 and 4:4:0 sampling: vp9_filter_block_plane_non420 for chroma, non-square
    # chroma blocks in MC / intra, q2 averages of sub-8x8 motion vectors (odd sizes, sub-8x8 partitions)
    "p422": dict(w=200, h=152, n=6, chroma="422", cpu_used=1, bitrate=900, end_usage=0, min_q=4, max_q=40, force_lf=True),
    "p440": dict(w=184, h=136, n=6, chroma="440", cpu_used=1, bitrate=900, end_usage=0, min_q=4, max_q=40, force_lf=True),
    "p422_hbd": dict(w=176, h=144, n=5, chroma="422", bit_depth=10, deadline=1, cpu_used=5, bitrate=300),
    "p440_rt": dict(w=352, h=288, n=8, chroma="440", deadline=1, cpu_used=5, bitrate=300, tile_cols_log2=1),
    # a forced key frame rewritten as a hidden intra-only frame, then shown with
    # show_existing_frame (bitstream edits in oracle/synth.py).  frame_parallel: without backward
    # adaptation the frames after the edit do not depend on last_frame_type == KEY_FRAME
    # (vp9_entropy.c:1063-1071), so the edited stream stays consistent with what was encoded.
    "intra_only": dict(w=352, h=288, n=8, deadline=1, cpu_used=5, bitrate=300, frame_parallel=1, force_kf_at={4},
                       intra_only_at={4},
                       show_existing_at={4: 0}),
    # show_existing_frame of other reference slots between ordinary frames
    "show_exist": dict(w=352, h=288, n=7, deadline=1, cpu_used=5, bitrate=300, show_existing_at={2: 1, 5: 2}),
    # scaled references at high bit depth (per-sample path of the u16 kernel)
    "scaled_hbd": dict(w=256, h=176, n=8, bit_depth=10, cpu_used=5, deadline=1, bitrate=250,
                       scale_at={2: (3, 1), 4: (1, 2), 6: (0, 0)}, temporal_noise=2.0),
    "scaled_rt": dict(w=416, h=240, n=10, cpu_used=6, deadline=1, bitrate=300,
                      scale_at={2: (3, 1), 5: (2, 3), 8: (0, 0)}, temporal_noise=2.0),
}

# Bench clips: stand-ins for BASELINE.json configs[1] (1440p) and configs[2]/[3] (4K, 4 tile
# columns, compound).  Real-time two-pass is not available, so good quality / cpu_used 5.
BENCH = {
    "bench_1440p": dict(w=2560, h=1440, n=24, cpu_used=2, lag=16, auto_alt_ref=1, two_pass=True, bitrate=6000,
                        end_usage=0, tile_cols_log2=2, arnr_maxframes=7, arnr_strength=5, detail=0.02,
                        temporal_noise=3.5, noise_corner=False, threads=8, force_lf=True),
    # BASELINE.json configs[4]: 1080p profile 2 (10-bit) -- highbd convolve / idct / lpf
    "bench_hbd1080": dict(w=1920, h=1080, n=16, bit_depth=10, cpu_used=3, lag=16, auto_alt_ref=1, two_pass=True,
                          bitrate=5000, end_usage=0, tile_cols_log2=2, arnr_maxframes=7, arnr_strength=5, detail=0.02,
                          temporal_noise=3.5, noise_corner=False, threads=8, force_lf=True),
    # BASELINE.json configs[3]: 4K, 4 tile columns, compound + scaled references.  libvpx ignores
    # VP8E_SET_SCALEMODE whenever frames are lagged (no alt-ref, hence no compound prediction, in the
    # encodes that can change the coded size), so the clip is two segments of the same content
    # spliced at a key frame: a two-pass alt-ref segment (hidden ARF + compound) followed by a
    # one-pass segment whose coded size goes 4K -> 4/5 -> 3/5 -> 4K, every step predicted from
    # references of the previous size (x_step_q4 != 16 on up to three references).
    "bench_4k": dict(segments=[
        dict(w=3840, h=2160, n=14, cpu_used=2, lag=16, auto_alt_ref=1, two_pass=True, bitrate=12000,
             end_usage=0, tile_cols_log2=2, arnr_maxframes=7, arnr_strength=5, detail=0.02,
             temporal_noise=3.5, noise_corner=False, threads=8, force_lf=True),
        dict(w=3840, h=2160, n=8, seed=77, cpu_used=3, cq_level=52, min_q=40, tile_cols_log2=2, detail=0.02, temporal_noise=3.5,
             noise_corner=False, threads=8, force_lf=True, scale_at={2: (1, 1), 4: (2, 2), 6: (0, 0)}),
    ]),
    # BASELINE.json configs[4], second half: lossless stream (4x4 WHT for every block) at bench size
    "bench_lossless": dict(w=1920, h=1080, n=6, lossless=1, cpu_used=4, tile_cols_log2=2, detail=0.02,
                           temporal_noise=2.0, noise_corner=False, threads=8),
}


def make(name, params, outdir):
    kw = dict(params)
    path = os.path.join(outdir, "clips", name + ".ivf")
    t = time.time()
    if "segments" in kw:      # independent encodes spliced at their key frames
        chunks = []
        for seg in kw["segments"]:
            seg = dict(seg)
            w, h, n = seg.pop("w"), seg.pop("h"), seg.pop("n")
            chunks += synth.encode_clip(path, w, h, n, **seg)
        vo.ivf_write(path, chunks, w, h)
    else:
        w, h, n = kw.pop("w"), kw.pop("h"), kw.pop("n")
        synth.encode_clip(path, w, h, n, **kw)
    md5 = vo.decode_ivf_md5(path)
    with open(os.path.join(outdir, "md5", name + ".md5"), "w") as f:
        for i, m in enumerate(md5):
            f.write("%s  img-%dx%d-%04d.i420\n" % (m, w, h, i + 1))
    print("%-12s %4dx%-4d %3d frames %8d bytes  %.1fs" % (name, w, h, len(md5), os.path.getsize(path), time.time() - t))


def main():
    os.makedirs(os.path.join(ROOT, "clips"), exist_ok=True)
    os.makedirs(os.path.join(ROOT, "md5"), exist_ok=True)
    args = sys.argv[1:]
    table = dict(PARITY)
    if args and args[0] == "bench":
        table = dict(BENCH)
        args = args[1:]
    for name in (args or list(table)):
        make(name, table[name], ROOT)


if __name__ == "__main__":
    main()
