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


def test_ctypes_mirror_has_the_layout_of_the_header(tmp_path):
    """sizeof / offsetof of vp9b_frame_hdr as the C compiler sees them == the hand-written ctypes mirror."""
    import subprocess
    fields = ["magic", "total_bytes", "width", "height", "mi_cols", "mi_rows", "sb_cols", "sb_rows", "bit_depth", "ss_x",
              "ss_y", "lossless", "intra_only", "show_frame", "lf_level", "lf_sharpness", "coef_bytes", "dst_slot", "ref", "sec",
              "lf_lim", "lf_mblim", "frame_index", "n_inter_px", "n_compound_px", "n_txb_inter_size"]
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "vp9_b200.h"\nint main(void) {\n' +
                   'printf("%zu %zu %zu %zu\\n", sizeof(vp9b_frame_hdr), sizeof(vp9b_image), sizeof(vp9b_out_frame), sizeof(vp9b_frame_buffer));\n' +
                   "".join('printf("%s %%zu\\n", offsetof(vp9b_frame_hdr, %s));\n' % (f, f) for f in fields) + "return 0; }\n")
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-std=c11", "-I" + os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    out = subprocess.check_output([str(exe)], text=True).splitlines()
    sizes = [int(x) for x in out[0].split()]
    assert sizes == [C.sizeof(vb.FrameHdr), C.sizeof(vb.capi.Image), C.sizeof(vb.capi.OutFrame), C.sizeof(vb.capi.FrameBuffer)]
    assert C.sizeof(vb.FrameHdr) == 416
    for line in out[1:]:
        name, off = line.split()
        assert getattr(vb.FrameHdr, name).offset == int(off), name


def test_struct_sizes_match_the_header():
    assert C.sizeof(vb.FrameHdr) == 416
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
