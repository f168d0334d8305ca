"""cuda-vp9_b200 -- B200-native VP9 decode-reconstruction engine (host-side Python mirror).

The product is the C-ABI library (include/vp9_b200.h): `libvp9b200.so` (CUDA kernels + parser +
decoder) and `libvp9b200_host.so` (parser only, no CUDA dependency).  This package is a thin
ctypes mirror of that ABI whose classes read like the reference's public API:

    Decoder.decode(chunk) / Decoder.get_frame()   ~ vpx_codec_decode / vpx_codec_get_frame
                                                    (libvpx/vpxdec.c:847,888)
    decode_ivf_md5(path)                          ~ vpxdec --md5 (vpxdec.c:1036-1042)

There is no CPU fallback for reconstruction: if the CUDA library or a GPU is missing the
reconstruction classes raise.
"""
from .capi import (HostParser, GpuRecon, Decoder, FrameHdr, lib_host, lib_gpu, VP9BError,  # noqa: F401
                   parse_superframe, NUM_SLOTS, MAX_BATCH, PinnedFrame)
from .ivf import ivf_read, ivf_write, md5_planes  # noqa: F401
from .vpxdec import decode_ivf_md5, decode_ivf  # noqa: F401
