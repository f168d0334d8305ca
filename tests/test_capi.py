"""The C-ABI libraries load without a GPU and export every symbol include/vp9_b200.h declares."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT

import cuda_vp9_b200 as vb


def declared_functions():
    src = open(os.path.join(ROOT, "include", "vp9_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vp9gpu_\w+|vp9b_parser_\w+|vp9b_decoder_\w+|vp9b_parse_superframe)\s*\(", src)))


def test_header_declares_the_boundary():
    names = declared_functions()
    for n in ("vp9gpu_create", "vp9gpu_frame_submit", "vp9gpu_get_frame", "vp9b_parser_parse", "vp9b_decoder_decode",
              "vp9b_decoder_get_frame"):
        assert n in names


def test_host_library_exports_parser_symbols():
    L = vb.lib_host()
    for n in declared_functions():
        if n.startswith("vp9b_parser_") or n == "vp9b_parse_superframe":
            assert hasattr(L, n), n


def test_gpu_library_exports_every_declared_symbol():
    p = os.path.join(ROOT, "cuda-vp9_b200", "libvp9b200.so")
    if not os.path.exists(p):
        pytest.skip("CUDA library not built in this environment")
    L = C.CDLL(p)  # loading must not need a GPU
    missing = [n for n in declared_functions() if not hasattr(L, n)]
    assert not missing, missing


def test_struct_sizes_match_the_header():
    assert C.sizeof(vb.FrameHdr) == 416 or C.sizeof(vb.FrameHdr) % 4 == 0
    # the C side checks the magic and total_bytes; a packed buffer produced by the parser parses back
    from conftest import clip_path
    _, _, chunks = vb.ivf_read(clip_path("odd2"))
    p = vb.HostParser()
    packed, show = p.parse(vb.parse_superframe(chunks[0])[0])
    h = vb.FrameHdr.from_buffer_copy(packed[: C.sizeof(vb.FrameHdr)])
    assert h.magic == 0x39425056 and h.total_bytes == len(packed)
    assert (h.width, h.height) == (130, 98) and h.intra_only == 1 and show == h.dst_slot


def test_superframe_split():
    frames = vb.parse_superframe(b"\x01\x02\x03")
    assert frames == [b"\x01\x02\x03"]
    # two frames of 2 and 3 bytes, 1-byte sizes: marker 0b11000001
    body = b"\xaa\xbb" + b"\xcc\xdd\xee"
    idx = bytes([0xc1, 2, 3, 0xc1])
    assert vb.parse_superframe(body + idx) == [b"\xaa\xbb", b"\xcc\xdd\xee"]


def test_parser_rejects_garbage():
    p = vb.HostParser()
    with pytest.raises(vb.VP9BError):
        p.parse(b"\x00\x00\x00\x00")


def test_gpu_classes_fail_loudly_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    p = os.path.join(ROOT, "cuda-vp9_b200", "libvp9b200.so")
    if not os.path.exists(p):
        with pytest.raises(ImportError):
            vb.lib_gpu()
    else:
        with pytest.raises(vb.VP9BError):
            vb.Decoder(0)
