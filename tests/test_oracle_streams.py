"""Pins the oracle (and the packing contract of the host parser): whole-stream per-frame MD5 of
product parser -> packed buffer -> oracle/vp9_recon_ref.c must equal (a) the committed libvpx
v1.15.2 MD5 lists and (b) a live libvpx decode when the library is present.  This is the
reference's own parity method (vpxdec --md5 vs Sony.md5 / netflix.md5, test/test_vector_test.cc:83-100)."""
import pytest

from conftest import clip_names, clip_path, golden_md5

import cuda_vp9_b200 as vb
from oracle import ref_decoder, vpx_oracle


@pytest.mark.parametrize("name", clip_names())
def test_cpu_decode_matches_golden_md5(name):
    _, _, chunks = vb.ivf_read(clip_path(name))
    got = ref_decoder.decode_chunks(chunks, want="md5")
    assert got == golden_md5(name)


@pytest.mark.parametrize("name", ["tiles_rt", "bench_1440p", "bench_hbd1080", "intra_only", "aq", "scaled_rt"])
@pytest.mark.parametrize("threads", [2, 4, 7])
def test_tile_parallel_parse_matches_golden_md5(name, threads):
    """Tile columns parsed by worker threads (vp9b_parser_set_threads): the merged packed buffer must
    reconstruct to the same frames (per-worker record lists are concatenated and rebased)."""
    _, _, chunks = vb.ivf_read(clip_path(name))
    n = 6 if name.startswith("bench_") else len(chunks)
    assert ref_decoder.decode_chunks(chunks[:n], want="md5", parse_threads=threads) == golden_md5(name)[:n]


@pytest.mark.skipif(not vpx_oracle.available(), reason="libvpx .so not present")
@pytest.mark.parametrize("name", ["odd_rt", "comp", "hbd10", "scaled"])
def test_golden_md5_matches_live_libvpx(name):
    assert vpx_oracle.decode_ivf_md5(clip_path(name)) == golden_md5(name)


@pytest.mark.skipif(not vpx_oracle.available(), reason="libvpx .so not present")
@pytest.mark.parametrize("name", ["odd_rt", "aq", "tiles_rt"])
def test_pre_loopfilter_stage_matches_libvpx(name):
    """Stage isolation: with VP9_SET_SKIP_LOOP_FILTER the reference output must equal the oracle
    stopped after intra reconstruction (stage 3)."""
    import numpy as np
    _, _, chunks = vb.ivf_read(clip_path(name))
    d = vpx_oracle.Decoder(skip_loop_filter=True)
    exp = []
    for c in chunks:
        exp += d.decode(c, want="planes")
    got = ref_decoder.decode_chunks(chunks, want="planes", stop_stage=3)
    assert len(exp) == len(got)
    for a, b in zip(exp, got):
        for pa, pb in zip(a, b):
            assert np.array_equal(pa, pb)


def test_features_are_exercised():
    """The clips must actually contain what they claim ('always verify')."""
    import ctypes as C
    stats = {}
    for name in clip_names():
        _, _, chunks = vb.ivf_read(clip_path(name))
        p = vb.HostParser()
        s = dict(comp=0, submv=0, scaled=0, intra_in_inter=0, lf=0, hidden=0, tiles=0, hbd=0, lossless=0)
        for ch in chunks:
            for fr in vb.parse_superframe(ch):
                packed, show = p.parse(fr)
                if packed is None:
                    continue
                h = vb.FrameHdr.from_buffer_copy(packed[: C.sizeof(vb.FrameHdr)])
                s["comp"] += sum(h.n_compound_px)
                s["submv"] += h.sec[1].count
                s["lf"] += h.lf_level > 0
                s["hidden"] += not h.show_frame
                s["hbd"] += h.bit_depth > 8
                s["lossless"] += h.lossless
                if not h.intra_only:
                    s["intra_in_inter"] += h.sec[3].count
                    s["scaled"] += any(r.x_step_q4 != 16 or r.y_step_q4 != 16 for r in h.ref)
        stats[name] = s
    assert stats["comp"]["comp"] > 0 and stats["comp"]["hidden"] > 0
    assert stats["sub8"]["submv"] > 0
    assert stats["scaled"]["scaled"] > 0 and stats["scaled_rt"]["scaled"] > 0
    assert stats["hbd10"]["hbd"] > 0 and stats["hbd12"]["hbd"] > 0
    assert stats["lossless"]["lossless"] > 0
    assert stats["odd_rt"]["lf"] > 0 and stats["aq"]["lf"] > 0
    assert sum(s["intra_in_inter"] for s in stats.values()) > 0
