"""IVF container + per-frame MD5, as vpxdec does it (libvpx/ivfdec.c:40-110, vpxdec.c:285-302)."""
import hashlib
import struct

import numpy as np


def ivf_read(path):
    with open(path, "rb") as f:
        b = f.read()
    if b[:4] != b"DKIF":
        raise ValueError("not an IVF file: %s" % path)
    hl = struct.unpack_from("<H", b, 6)[0]
    w, h = struct.unpack_from("<HH", b, 12)
    pos = hl
    frames = []
    while pos + 12 <= len(b):
        sz = struct.unpack_from("<I", b, pos)[0]
        frames.append(b[pos + 12: pos + 12 + sz])
        pos += 12 + sz
    return w, h, frames


def ivf_write(path, frames, w, h, fps=30):
    with open(path, "wb") as f:
        f.write(struct.pack("<4sHH4sHHIII4x", b"DKIF", 0, 32, b"VP90", w, h, fps, 1, len(frames)))
        for i, d in enumerate(frames):
            f.write(struct.pack("<IQ", len(d), i))
            f.write(d)


def md5_planes(planes):
    """MD5 over cropped Y, U, V rows; 2 bytes/sample little endian for high bit depth."""
    m = hashlib.md5()
    for a in planes:
        m.update(np.ascontiguousarray(a).tobytes())
    return m.hexdigest()
