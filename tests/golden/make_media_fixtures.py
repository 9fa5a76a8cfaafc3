"""Derive small golden fixtures from the reference's published images.

Run in the build container only (it reads /root/reference/media, which does not
exist on the GPU box):

    python tests/golden/make_media_fixtures.py

Output: tests/golden/media_fixtures.npz holding, for each of the three PNGs the
reference's README shows (README.md:72-77 gives the command that made each),
  <name>_blocks : block-mean of the 16-bit RGB image (float32, [H/b, W/b, 3])
  <name>_px00   : the exact 16-bit RGB value of pixel (0,0) — the known answer of
                  the NaN path
  <name>_lit    : fraction of pixels with any non-zero channel
The PNGs are 16-bit RGB (colour type 2), which Pillow truncates to 8 bits, so
they are decoded here with zlib directly.
"""
import os
import struct
import zlib

import numpy as np

MEDIA = "/root/reference/media"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "media_fixtures.npz")
IMAGES = {  # name -> (file, block)
    "poisson_saturne": ("poisson-saturne.png", 8),
    "solar_sail": ("solar-sail.png", 10),
    "solar_sail_220": ("solar-sail-220deg.png", 10),
}


def decode_png16(path):
    data = open(path, "rb").read()
    assert data[:8] == b"\x89PNG\r\n\x1a\n"
    pos, idat, ihdr = 8, [], None
    while pos < len(data):
        (ln,), typ = struct.unpack(">I", data[pos:pos + 4]), data[pos + 4:pos + 8]
        body = data[pos + 8:pos + 8 + ln]
        if typ == b"IHDR":
            ihdr = struct.unpack(">IIBBBBB", body)
        elif typ == b"IDAT":
            idat.append(body)
        pos += 12 + ln
    w, h, depth, ctype, _, _, interlace = ihdr
    assert (depth, ctype, interlace) == (16, 2, 0), ihdr
    bpp = 6
    raw = np.frombuffer(zlib.decompress(b"".join(idat)), dtype=np.uint8).reshape(h, 1 + w * bpp)
    out = np.zeros((h, w * bpp), dtype=np.uint8)
    prev = np.zeros(w * bpp, dtype=np.int32)
    for y in range(h):
        ft, line = int(raw[y, 0]), raw[y, 1:].astype(np.int32)
        if ft == 0:
            cur = line
        elif ft == 2:
            cur = (line + prev) & 255
        elif ft == 1:
            cur = line.copy()
            for c in range(bpp):  # per byte lane, running sum mod 256
                cur[c::bpp] = np.cumsum(line[c::bpp]) & 255
        else:  # 3 (average) / 4 (paeth): sequential
            cur = np.zeros_like(line)
            lb, pb = line.tolist(), prev.tolist()
            cb = [0] * len(lb)
            for i in range(len(lb)):
                a = cb[i - bpp] if i >= bpp else 0
                b = pb[i]
                c = pb[i - bpp] if i >= bpp else 0
                if ft == 3:
                    pred = (a + b) >> 1
                else:
                    p = a + b - c
                    pa, pbb, pc = abs(p - a), abs(p - b), abs(p - c)
                    pred = a if (pa <= pbb and pa <= pc) else (b if pbb <= pc else c)
                cb[i] = (lb[i] + pred) & 255
            cur = np.array(cb, dtype=np.int32)
        out[y] = cur
        prev = cur
    img = out.reshape(h, w, 3, 2).astype(np.uint16)
    return (img[..., 0] << 8) | img[..., 1]


def main():
    store = {}
    for name, (fn, b) in IMAGES.items():
        img = decode_png16(os.path.join(MEDIA, fn))
        h, w, _ = img.shape
        assert h % b == 0 and w % b == 0
        blocks = img.astype(np.float64).reshape(h // b, b, w // b, b, 3).mean(axis=(1, 3))
        store[name + "_blocks"] = blocks.astype(np.float32)
        store[name + "_px00"] = img[0, 0].copy()
        store[name + "_lit"] = np.float64((img.max(axis=2) > 0).mean())
        store[name + "_shape"] = np.array([h, w, b])
        print(name, img.shape, "px00", img[0, 0], "lit", store[name + "_lit"])
    np.savez_compressed(OUT, **store)
    print("wrote", OUT, os.path.getsize(OUT))


if __name__ == "__main__":
    main()
