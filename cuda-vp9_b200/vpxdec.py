"""vpxdec-like helpers over the GPU decoder (libvpx/vpxdec.c main_loop :517, --md5 :1036-1042)."""
from .capi import Decoder
from .ivf import ivf_read, md5_planes


def decode_ivf(path, device=-1, want="planes"):
    """Decode an IVF file on the GPU; returns a list with one entry per shown frame
    (planes, md5 string, or None for want='none')."""
    _, _, chunks = ivf_read(path)
    d = Decoder(device)
    out = []
    for ch in chunks:
        d.decode(ch)
        if want == "none":
            img = d.get_frame(read=False)
            if img is not None:
                out.append(None)
            continue
        pl = d.get_frame()
        if pl is not None:
            out.append(md5_planes(pl) if want == "md5" else pl)
    d.close()
    return out


def decode_ivf_md5(path, device=-1):
    return decode_ivf(path, device, want="md5")
