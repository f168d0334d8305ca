"""GPU parity (runs on the B200 box): every frame of every clip decoded through the C ABI must be
bit-exact with the libvpx MD5 lists, and every reconstruction stage must match the CPU oracle."""
import ctypes as C

import numpy as np
import pytest

from conftest import clip_names, clip_path, golden_md5

import cuda_vp9_b200 as vb
from oracle import ref_decoder

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", clip_names())
def test_gpu_decode_md5(name):
    got = vb.decode_ivf_md5(clip_path(name), device=0)
    assert got == golden_md5(name)


def _stage_frames(name, stage):
    """Decode with recon stopped after `stage` on both the GPU and the oracle.  After each frame
    the slot is overwritten with the full-recon oracle result so that later frames see correct
    references on both sides."""
    _, _, chunks = vb.ivf_read(clip_path(name))
    parser = vb.HostParser()
    gpu = vb.GpuRecon(0)
    ref_full = ref_decoder.RefRecon()
    ref_stage = ref_decoder.RefRecon()
    out = []
    for ch in chunks:
        for fr in vb.parse_superframe(ch):
            packed, show = parser.parse(fr)
            if packed is None:
                continue
            h = vb.FrameHdr.from_buffer_copy(packed[: C.sizeof(vb.FrameHdr)])
            geom = (h.width, h.height, h.bit_depth, h.ss_x, h.ss_y)
            # same starting content on both sides (partial stages leave some samples untouched)
            dt = np.uint16 if h.bit_depth > 8 else np.uint8
            shapes = [(h.height, h.width)] + [((h.height + h.ss_y) >> h.ss_y, (h.width + h.ss_x) >> h.ss_x)] * 2
            blank = [np.full(s, 7, dtype=dt) for s in shapes]
            gpu.set_frame(h.dst_slot, blank, h.bit_depth, h.ss_x, h.ss_y)
            ref_stage.set_frame(h.dst_slot, blank, h.bit_depth, h.ss_x, h.ss_y)
            gpu.set_stop_stage(stage)
            gpu.submit(packed)
            g = gpu.get_frame(h.dst_slot, *geom)
            ref_stage.submit(packed, stage)
            r = ref_stage.get_frame(h.dst_slot, *geom)
            out.append((h.frame_index, g, r))
            # bring all three states to the true frame
            ref_full.submit(packed, 0)
            full = ref_full.get_frame(h.dst_slot, *geom)
            ref_stage.submit(packed, 0)
            gpu.set_stop_stage(0)
            gpu.submit(packed)
            assert all(np.array_equal(a, b) for a, b in zip(gpu.get_frame(h.dst_slot, *geom), full)), \
                "full recon mismatch at frame %d" % h.frame_index
    return out


@pytest.mark.parametrize("stage", [1, 2, 3])
@pytest.mark.parametrize("name", ["odd_rt", "comp", "sub8", "hbd10", "scaled", "lossless", "tiles_rt"])
def test_gpu_stage_matches_oracle(name, stage):
    for idx, g, r in _stage_frames(name, stage):
        for p in range(3):
            if not np.array_equal(g[p], r[p]):
                d = np.argwhere(g[p] != r[p])
                pytest.fail("%s stage %d frame %d plane %d: %d samples differ, first at (row,col)=%s gpu=%d ref=%d" %
                            (name, stage, idx, p, len(d), tuple(d[0]), g[p][tuple(d[0])], r[p][tuple(d[0])]))


def test_gpu_launches_counted():
    d = vb.Decoder(0)
    _, _, chunks = vb.ivf_read(clip_path("odd_rt"))
    for ch in chunks:
        d.decode(ch)
        d.get_frame()
    assert d.gpu_launch_count() > len(chunks)


@pytest.mark.gpu
def test_gpu_rejects_corrupt_packed_buffers():
    """vp9gpu_frame_submit checks what it can on the host and returns an error instead of
    launching on a buffer whose sections or geometry are inconsistent."""
    _, _, chunks = vb.ivf_read(clip_path("odd_rt"))
    parser = vb.HostParser()
    packed, _ = parser.parse(vb.parse_superframe(chunks[0])[0])
    hdr_size = C.sizeof(vb.FrameHdr)
    gpu = vb.GpuRecon(0)

    def mutated(fn):
        h = vb.FrameHdr.from_buffer_copy(packed[:hdr_size])
        fn(h)
        return bytes(h) + packed[hdr_size:]

    def bad_magic(h): h.magic = 0
    def bad_section(h): h.sec[3].offset = h.total_bytes   # TXB_INTRA_Y is not empty on a key frame
    def bad_count(h): h.sec[7].count = 0x7fffffff
    def bad_geometry(h): h.sb_cols += 1
    def bad_slot(h): h.dst_slot = 99
    def bad_depth(h): h.bit_depth = 9

    for fn in (bad_magic, bad_section, bad_count, bad_geometry, bad_slot, bad_depth):
        with pytest.raises(vb.VP9BError):
            gpu.submit(mutated(fn))
    with pytest.raises(vb.VP9BError):
        gpu.submit(packed[:hdr_size - 8])
    gpu.submit(packed)          # the context is still usable afterwards
    gpu.sync()
    gpu.close()
    parser.close()


@pytest.mark.gpu
def test_gpu_concurrent_streams_stay_bit_exact():
    """Independent decoder contexts on concurrent host threads / CUDA streams (how the engine is
    deployed): every stream must still match its golden MD5 list."""
    import threading
    names = ["comp_lf", "tiles_rt", "hbd10", "scaled_rt", "sub8", "p444", "aq", "odd_rt", "intra_only", "lossless",
             "comp_lf", "tiles_rt"]
    results = {}

    def work(i, name):
        try:
            got = []
            for _ in range(3):               # several passes: contexts overlap in every phase
                got = vb.decode_ivf_md5(clip_path(name))
            results[i] = (name, got == golden_md5(name))
        except Exception as e:               # noqa: BLE001
            results[i] = (name, repr(e))

    th = [threading.Thread(target=work, args=(i, n)) for i, n in enumerate(names)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    bad = {i: r for i, r in results.items() if r[1] is not True}
    assert not bad and len(results) == len(names), bad


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["hbd10", "hbd12", "p444_hbd", "odd_rt"])
def test_gpu_8bit_output_equals_libvpx_downshift(name):
    """vp9gpu_get_frame_8bit_async == vpx_img_downshift of libvpx's own output (vpxdec
    --output-bit-depth=8, vpxdec.c:950-978; tools_common.c highbd_img_downshift: sample >> shift).  The
    reference frames come from libvpx v1.15.2, not from this decoder."""
    from oracle import vpx_oracle as vo
    _, _, chunks = vb.ivf_read(clip_path(name))
    parser, gpu, ref = vb.HostParser(), vb.GpuRecon(0), vo.Decoder(threads=1)
    shown = 0
    for ch in chunks:
        want = ref.decode(ch, want="planes")
        got = []
        for fr in vb.parse_superframe(ch):
            packed, show = parser.parse(fr)
            if packed is not None:
                gpu.submit(packed)
            if show < 0:
                continue
            w, h, bd, sx, sy = parser.slot_info(show)
            got = [(gpu.get_frame_8bit(show, w, h, sx, sy), bd)]     # the chunk's last shown frame is its output
        assert len(got) == len(want)
        for (eight, bd), full in zip(got, want):
            for a, b in zip(full, eight):
                assert b.dtype == np.uint8 and np.array_equal((np.asarray(a).astype(np.uint16) >> (bd - 8)).astype(np.uint8), b)
            shown += 1
    assert shown > 0
    gpu.close()
    parser.close()
    ref.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["odd_rt", "hbd10", "scaled_rt", "show_exist"])
def test_gpu_external_frame_buffers(name):
    """vp9b_decoder_set_frame_buffer_functions: every shown frame is delivered in an application
    (page-locked) buffer, bit-exact, and every buffer comes back through release()."""
    import ctypes as C
    w, h, chunks = vb.ivf_read(clip_path(name))
    L = vb.lib_gpu()
    pool = {}            # tag -> (address, size, in use)
    stats = {"get": 0, "release": 0}

    def get(min_size):
        stats["get"] += 1
        for tag, (addr, size, used) in pool.items():
            if not used and size >= min_size:
                pool[tag] = (addr, size, True)
                return addr, size, tag
        tag = len(pool) + 1
        size = max(min_size, 1 << 16)
        addr = L.vp9gpu_alloc_host(size)
        assert addr
        pool[tag] = (addr, size, True)
        return addr, size, tag

    def release(tag):
        stats["release"] += 1
        addr, size, used = pool[tag]
        assert used
        pool[tag] = (addr, size, False)

    d = vb.Decoder(0)
    d.set_frame_buffer_functions(get, release)
    md5, held = [], []
    for ch in chunks:
        d.decode(ch)
        r = d.get_frame_ext()
        if r is None:
            continue
        held.append(r)
        if len(held) == 3:       # keep a few frames in flight, as a renderer would
            out, planes = held.pop(0)
            d.wait(out.ticket)
            md5.append(vb.md5_planes(planes))
            d.release_frame(out)
    for out, planes in held:
        d.wait(out.ticket)
        md5.append(vb.md5_planes(planes))
        d.release_frame(out)
    assert md5 == golden_md5(name)
    assert stats["get"] == stats["release"] == len(md5) and len(pool) <= 4
    with pytest.raises(vb.VP9BError):      # like libvpx: not after decoding has started
        d.set_frame_buffer_functions(get, release)
    d.close()
    for addr, _, _ in pool.values():
        L.vp9gpu_free_host(addr)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["odd_rt", "hbd10", "p444"])
def test_c_client_prints_the_golden_md5_list(name):
    """examples/vpxdec_md5.c: a plain C program on the C ABI reproduces `vpxdec --md5`."""
    import os
    import subprocess
    from conftest import ROOT
    exe = os.path.join(ROOT, "examples", "vpxdec_md5")
    if not os.path.exists(exe):      # normally built by __graft_entry__.build(); it is one gcc call
        subprocess.check_call(["gcc", "-O2", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "vpxdec_md5.c"),
                               "-L" + os.path.join(ROOT, "cuda-vp9_b200"), "-lvp9b200",
                               "-Wl,-rpath,$ORIGIN/../cuda-vp9_b200", "-o", exe])
    r = subprocess.run([exe, clip_path(name)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    assert [l.split()[0] for l in r.stdout.splitlines()] == golden_md5(name)


def _decode_lockstep(names):
    """Decode several clips with one frame of each per vp9gpu_frame_submit_device_multi call."""
    import hashlib
    streams = []
    for nme in names:
        _, _, chunks = vb.ivf_read(clip_path(nme))
        frames = [fr for ch in chunks for fr in vb.parse_superframe(ch)]
        streams.append(dict(name=nme, frames=frames, parser=vb.HostParser(), gpu=vb.GpuRecon(0), md5=[], pos=0))
    while any(s["pos"] < len(s["frames"]) for s in streams):
        members, shows = [], []
        for s in streams:
            if s["pos"] >= len(s["frames"]):
                continue
            packed, show = s["parser"].parse(s["frames"][s["pos"]])
            s["pos"] += 1
            if packed is not None:
                members.append((s, s["gpu"].upload(packed)))
            shows.append((s, show))
        if members:
            pack = vb.GpuRecon.make_multi([s["gpu"] for s, _ in members], [up for _, up in members])
            vb.GpuRecon.submit_multi(pack)
        for s, show in shows:
            if show >= 0:
                w, h, bd, sx, sy = s["parser"].slot_info(show)
                planes = s["gpu"].get_frame(show, w, h, bd, sx, sy)
                s["md5"].append(vb.md5_planes(planes))
        for s, (d, _) in members:
            s["gpu"].free(d)
    out = {s["name"] + str(i): s["md5"] for i, s in enumerate(streams)}
    for s in streams:
        s["gpu"].close()
        s["parser"].close()
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("names", [["comp_lf", "odd_rt", "tiles_rt", "aq"], ["sub8", "scaled_rt", "intra_only"],
                                   ["hbd10", "hbd12", "scaled_hbd"], ["p444", "odd_rt", "hbd10", "comp_lf"]],
                         ids=["u8_420_x4", "u8_420_x3", "hbd_x3", "mixed_types"])
def test_gpu_batched_contexts_stay_bit_exact(names):
    """Several contexts per kernel launch (frames of different sizes, lengths and stage sets)."""
    got = _decode_lockstep(names)
    for i, nme in enumerate(names):
        assert got[nme + str(i)] == golden_md5(nme), nme


@pytest.mark.gpu
def test_decoder_survives_corrupt_chunks():
    """A corrupt chunk is reported as an error (vpx_codec_decode returns VPX_CODEC_CORRUPT_FRAME,
    vpxdec.c:896-901) and the decoder keeps working when the stream restarts at a key frame."""
    _, _, chunks = vb.ivf_read(clip_path("odd_rt"))
    d = vb.Decoder(0)
    for bad in (b"\x00" * 40, chunks[1][: len(chunks[1]) // 3], bytes(reversed(chunks[0]))):
        with pytest.raises(vb.VP9BError):
            d.decode(bad)
    got = []
    for ch in chunks:                      # restart at the key frame
        d.decode(ch)
        planes = d.get_frame()
        if planes is not None:
            got.append(vb.md5_planes(planes))
    assert got == golden_md5("odd_rt")
    d.close()
