"""Output conversion + raw containers (src/bin/main.rs:40-100): device path vs the oracle restatement.

CPU part: container headers and sizes (host logic of the library, no GPU).  GPU part: every
(pixel format, container) pair byte for byte, the encoded sequence, and the RGB16 known answer of
the reference's published image (pixel (0,0) of media/solar-sail-220deg.png = 65535, 65535, 60849).
The u16 -> u8 rule and the container layouts come from the `image` 0.25 crate, which is not under
/root/reference: both sides restate its published source (parity unpinned, third-party)."""
import ctypes as C
import math

import numpy as np
import pytest

FORMATS = [0, 1, 2, 3]           # SAR_PIX_RGBA16, RGB16, RGBA8, RGB8
CONTAINERS = [0, 1, 2, 3]        # SAR_FILE_RAW, PAM, BMP, PNG (stored deflate)


def test_headers_and_sizes_match_the_oracle(oracle):
    from strange_attractor_renderer_b200 import _native as N

    L = N.lib()
    rng = np.random.default_rng(5)
    for (w, h) in [(1, 1), (3, 2), (5, 7), (450, 501), (2048, 2048), (9000, 3)]:
        img = rng.integers(0, 65536, (h, w, 4), dtype=np.uint16) if w * h < 10_000 else np.zeros((h, w, 4), np.uint16)
        for fmt in FORMATS:
            for cont in CONTAINERS:
                ref = oracle.encode(img, fmt, cont)
                size = L.sar_encoded_size(w, h, fmt, cont)
                if ref is None:                       # BMP cannot hold 16-bit samples (the reference panics)
                    assert size == 0 and cont == 2 and fmt in (0, 1)
                    assert L.sar_encode_header(w, h, fmt, cont, None, 0, None) == N.SAR_ERR_UNSUPPORTED
                    continue
                assert size == ref.size
                hb = C.c_size_t()
                N.check(L.sar_encode_header(w, h, fmt, cont, None, 0, C.byref(hb)))
                head = np.zeros(hb.value, np.uint8)
                N.check(L.sar_encode_header(w, h, fmt, cont, head.ctypes.data_as(N._u8p), head.size, None))
                assert np.array_equal(head, ref[:hb.value])
    # the PAM header is plain text: spot-check it against the format definition
    hb = C.c_size_t()
    buf = np.zeros(160, np.uint8)
    N.check(L.sar_encode_header(12, 34, 1, 1, buf.ctypes.data_as(N._u8p), 160, C.byref(hb)))
    assert bytes(buf[:hb.value]) == b"P7\nWIDTH 12\nHEIGHT 34\nDEPTH 3\nMAXVAL 65535\nTUPLTYPE RGB\nENDHDR\n"
    assert L.sar_encoded_size(12, 34, 9, 0) == 0 and L.sar_encoded_size(0, 4, 0, 0) == 0


def test_oracle_png_decodes_to_the_converted_pixels(oracle):
    """The stored-deflate PNG (oracle side) is a valid PNG: Pillow / zlib read back exactly the converted samples."""
    import io
    import zlib

    from PIL import Image

    rng = np.random.default_rng(7)
    img = rng.integers(0, 65536, (37, 53, 4), dtype=np.uint16)
    for fmt, mode, nch in ((2, "RGBA", 4), (3, "RGB", 3)):
        png = oracle.encode(img, fmt, 3)
        got = np.asarray(Image.open(io.BytesIO(png.tobytes())).convert(mode))
        want = ((img[..., :nch].astype(np.uint32) + 128) // 257).astype(np.uint8)
        assert np.array_equal(got, want)
    # 16-bit: Pillow has no 16-bit RGB(A) mode, so inflate the IDAT by hand and compare the big-endian scanlines
    for fmt, nch in ((0, 4), (1, 3)):
        png = oracle.encode(img, fmt, 3).tobytes()
        assert png[:8] == b"\x89PNG\r\n\x1a\n" and png[12:16] == b"IHDR" and png[24] == 16 and png[25] == (6 if nch == 4 else 2)
        n = int.from_bytes(png[33:37], "big")
        assert png[37:41] == b"IDAT" and zlib.crc32(png[37:41 + n]) == int.from_bytes(png[41 + n:45 + n], "big")
        raw = zlib.decompress(png[41:41 + n])
        rows = np.frombuffer(raw, np.uint8).reshape(37, 1 + 53 * nch * 2)
        assert (rows[:, 0] == 0).all()
        assert np.array_equal(rows[:, 1:].reshape(37, 53, nch, 2).astype(np.uint16) @ np.array([256, 1], np.uint16), img[..., :nch])
        assert png[-12:] == b"\x00\x00\x00\x00IEND\xaeB`\x82"


def test_u16_to_u8_rule_is_round_to_nearest(oracle):
    """image 0.25: (c + 128) / 257 == round(c * 255 / 65535) for every u16."""
    c = np.arange(65536, dtype=np.uint32)
    rule = (c + 128) // 257
    assert np.array_equal(rule, np.floor(c * 255 / 65535 + 0.5).astype(np.uint32))
    img = np.zeros((1, 65536, 4), np.uint16)
    img[0, :, 0] = c
    out = oracle.encode(img, 3, 0).reshape(65536, 3)
    assert np.array_equal(out[:, 0], rule.astype(np.uint8))


@pytest.mark.gpu
def test_device_conversion_matches_oracle_bytes(oracle):
    import strange_attractor_renderer_b200 as S

    cfg = S.Config.solar_sail()
    cfg.width, cfg.height, cfg.iterations, cfg.angle = 203, 97, 30_000, 220.0 * math.pi / 180.0
    rt = S.Runtime.new(cfg)
    S.render(cfg, rt, initial_points=S.seed_points(7, 0, 128))
    for transparent in (True, False):
        cfg.transparent = transparent
        img = S.colorize(cfg, rt)
        for fmt in S.PixelFormat:
            for cont in S.Container:                              # PngDeflate: no fixed-size form on either side
                ref = oracle.encode(img, fmt.value, cont.value)
                if ref is None:
                    with pytest.raises(S.SarError):
                        S.encode_image(rt, fmt, cont)
                    continue
                got = S.encode_image(rt, fmt, cont)
                assert np.array_equal(got, ref), (fmt, cont, int((got != ref).sum()))
    # a frame wider than one stored deflate block (65 535 bytes): PNG scanlines straddle block boundaries
    wide = S.Config.poisson_saturne()
    wide.width, wide.height, wide.iterations = 8300, 5, 2_000
    wrt = S.Runtime.new(wide)
    S.render(wide, wrt, initial_points=S.seed_points(2, 0, 64))
    wimg = S.colorize(wide, wrt)
    assert np.array_equal(S.encode_image(wrt, S.PixelFormat.Rgba16, S.Container.Png), oracle.encode(wimg, 0, 3))
    import io
    from PIL import Image
    png8 = S.encode_image(rt, S.PixelFormat.Rgba8, S.Container.Png)
    assert np.array_equal(np.asarray(Image.open(io.BytesIO(png8.tobytes())).convert("RGBA")),
                          ((S.colorize(cfg, rt).astype(np.uint32) + 128) // 257).astype(np.uint8))
    # known answer of the reference's published image: pixel (0,0), RGB16 big-endian in a PAM
    cfg.transparent = False
    S.colorize(cfg, rt)
    pam = S.encode_image(rt, S.PixelFormat.of(False, False), S.Container.Pam)
    head = bytes(pam[:80]).split(b"ENDHDR\n")[0] + b"ENDHDR\n"
    px = [int(v) for v in pam[len(head):len(head) + 6]]
    assert [px[0] << 8 | px[1], px[2] << 8 | px[3], px[4] << 8 | px[5]] == [65535, 65535, 60849]


@pytest.mark.gpu
def test_encoded_sequence_equals_per_frame_encoding(oracle, tmp_path):
    import strange_attractor_renderer_b200 as S

    cfg = S.Config.poisson_saturne()
    cfg.width, cfg.height, cfg.iterations, cfg.transparent = 160, 120, 2_000_000, True
    angles = S.angle_iter(0.0, 60.0, 20.0)
    r = S.ParallelRenderer.new(threads=128)
    plain = S.render_sequence(r, cfg, angles, 1, seed=3)
    for fmt, cont in ((S.PixelFormat.Rgb8, S.Container.Bmp), (S.PixelFormat.Rgba16, S.Container.Pam), (S.PixelFormat.Rgb16, S.Container.Raw),
                      (S.PixelFormat.Rgba16, S.Container.Png), (S.PixelFormat.Rgb8, S.Container.Png)):
        enc = S.render_sequence_encoded(r, cfg, angles, 1, fmt, cont, seed=3)
        seen = []
        S.render_sequence_encoded(r, cfg, angles, 1, fmt, cont, seed=3, callback=lambda f, b: seen.append((f, b.copy())))
        assert [f for f, _ in seen] == list(range(len(angles)))
        for f in range(len(angles)):
            ref = oracle.encode(plain[f], fmt.value, cont.value)
            assert np.array_equal(enc[f], ref) and np.array_equal(seen[f][1], ref)
    # write_image_matches: convert, set the extension, write (main.rs:40-100)
    rt = S.Runtime.new(cfg)
    S.render(cfg, rt, initial_points=S.seed_points(1, 0, 32))
    img = S.colorize(cfg, rt)
    path = S.write_image(rt, str(tmp_path / "frame.png"), transparent=False, eight_bit=True, container=S.Container.Bmp)
    assert path.endswith("frame.bmp")
    assert np.array_equal(np.fromfile(path, dtype=np.uint8), oracle.encode(img, 3, 2))
    r.shutdown()


def _png_rows(png: bytes, w: int, h: int, bpp: int):
    """Structure checks of a one-IDAT PNG + inflate + undo the per-row filter (types 0 and 1) -> sample bytes [h, w*bpp]."""
    import zlib

    assert png[:8] == b"\x89PNG\r\n\x1a\n" and png[12:16] == b"IHDR"
    assert int.from_bytes(png[16:20], "big") == w and int.from_bytes(png[20:24], "big") == h
    assert zlib.crc32(png[12:29]) == int.from_bytes(png[29:33], "big")
    n = int.from_bytes(png[33:37], "big")
    assert png[37:41] == b"IDAT" and zlib.crc32(png[37:41 + n]) == int.from_bytes(png[41 + n:45 + n], "big")
    assert png[45 + n:] == b"\x00\x00\x00\x00IEND\xaeB`\x82"
    raw = zlib.decompress(png[41:41 + n])                       # checks the Adler-32 as well
    rows = np.frombuffer(raw, np.uint8).reshape(h, 1 + w * bpp)
    out = np.empty((h, w * bpp), np.uint8)
    for y in range(h):
        ft, line = int(rows[y, 0]), rows[y, 1:]
        assert ft in (0, 1)
        out[y] = line if ft == 0 else (np.cumsum(line.reshape(w, bpp).astype(np.uint32), axis=0) & 255).astype(np.uint8).reshape(-1)
    return out, n


@pytest.mark.gpu
def test_device_png_deflate_decodes_to_the_reference_pixels(oracle):
    """sar_runtime_encode_png (main.rs:78-89 with the compressor, on the device): a valid PNG whose pixels are exactly the
    converted image — checked with zlib's inflate + CRC/Adler and with Pillow — and far smaller than the stored form."""
    import io
    import time

    from PIL import Image

    import strange_attractor_renderer_b200 as S

    cfg = S.Config.solar_sail()
    cfg.width, cfg.height, cfg.iterations, cfg.angle = 203, 97, 30_000, 220.0 * math.pi / 180.0
    rt = S.Runtime.new(cfg)
    S.render(cfg, rt, initial_points=S.seed_points(7, 0, 128))
    for transparent in (True, False):
        cfg.transparent = transparent
        img = S.colorize(cfg, rt)
        for fmt in S.PixelFormat:
            wide, alpha = fmt in (S.PixelFormat.Rgba16, S.PixelFormat.Rgb16), fmt in (S.PixelFormat.Rgba16, S.PixelFormat.Rgba8)
            nch = 4 if alpha else 3
            bpp = nch * (2 if wide else 1)
            png = S.encode_png(rt, fmt).tobytes()
            assert png[24] == (16 if wide else 8) and png[25] == (6 if alpha else 2)
            samples, n = _png_rows(png, 203, 97, bpp)
            if wide:
                want = img[..., :nch].astype(">u2").view(np.uint8).reshape(97, 203 * bpp)
            else:
                want = ((img[..., :nch].astype(np.uint32) + 128) // 257).astype(np.uint8).reshape(97, 203 * bpp)
                pil = np.asarray(Image.open(io.BytesIO(png)).convert("RGBA" if alpha else "RGB"))
                assert np.array_equal(pil.reshape(97, 203 * bpp), want)
            assert np.array_equal(samples, want), fmt
            assert len(png) < S._native.lib().sar_encoded_size(203, 97, fmt.value, 3)
    # shapes around the block (16 KB) and lane (512 B) sizes, a 1x1 image, rows longer than a block
    for w, h in ((1, 1), (2730, 1), (2731, 1), (85, 32), (8300, 5), (3, 1400)):
        c2 = S.Config.poisson_saturne()
        c2.width, c2.height, c2.iterations, c2.transparent = w, h, 3_000, False
        r2 = S.Runtime.new(c2)
        S.render(c2, r2, initial_points=S.seed_points(2, 0, 64))
        im2 = S.colorize(c2, r2)
        samples, _ = _png_rows(S.encode_png(r2, S.PixelFormat.Rgb16).tobytes(), w, h, 6)
        assert np.array_equal(samples, im2[..., :3].astype(">u2").view(np.uint8).reshape(h, w * 6)), (w, h)
    # a full-size frame: the README's poisson-saturne command at 1e8 iterations; size in the range of the reference's file
    big = S.Config.poisson_saturne()
    big.width, big.height, big.iterations, big.transparent = 1920, 1080, 100_000_000, False
    big.colors.brighness.offset = -0.25
    r = S.ParallelRenderer.new()
    im = S.render_parallel(r, big, 1, seed=9)
    S.encode_png(r.runtime(), S.PixelFormat.Rgb16)             # warm the scratch allocation
    t0 = time.perf_counter()
    png = S.encode_png(r.runtime(), S.PixelFormat.Rgb16).tobytes()
    dt = time.perf_counter() - t0
    samples, n = _png_rows(png, 1920, 1080, 6)
    assert np.array_equal(samples, im[..., :3].astype(">u2").view(np.uint8).reshape(1080, 1920 * 6))
    raw_len = 1080 * (1 + 1920 * 6)
    print(f"\npng deflate 1920x1080 RGB16: {raw_len} -> {len(png)} bytes ({len(png) / raw_len:.3f}), {dt * 1e3:.2f} ms incl. copy-out")
    assert len(png) < 0.45 * raw_len
    r.shutdown()


@pytest.mark.gpu
def test_sequence_of_compressed_png_frames(oracle):
    """sar_render_sequence_encoded with SAR_FILE_PNG_DEFLATE (main.rs:496-512 with the default PNG branch of
    write_image_matches): every frame a complete compressed PNG handed to the callback in frame order; decodes to the
    pixels of the same frame rendered alone; the fixed-size entry points refuse the variable-size container."""
    import strange_attractor_renderer_b200 as S

    cfg = S.Config.solar_sail()
    cfg.width, cfg.height, cfg.iterations, cfg.transparent = 320, 200, 2_000_000, False
    angles = S.angle_iter(0.0, 70.0, 10.0)                        # 7 frames: more than the two slots per device
    r = S.ParallelRenderer.new(threads=2048)
    plain = S.render_sequence(r, cfg, angles, 2, seed=3)
    order = []
    frames = {}
    S.render_sequence_encoded(r, cfg, angles, 2, S.PixelFormat.Rgb16, S.Container.PngDeflate, seed=3,
                              callback=lambda f, b: (order.append(f), frames.__setitem__(f, bytes(b))))
    assert order == list(range(len(angles)))
    for f in range(len(angles)):
        samples, n = _png_rows(frames[f], 320, 200, 6)
        assert np.array_equal(samples, plain[f][..., :3].astype(">u2").view(np.uint8).reshape(200, 320 * 6)), f
        assert len(frames[f]) < 0.7 * 200 * (1 + 320 * 6)
    # without a callback the wrapper collects the frames; 8-bit RGBA
    got = S.render_sequence_encoded(r, cfg, angles[:3], 2, S.PixelFormat.Rgba8, S.Container.PngDeflate, seed=3)
    for f in range(3):
        samples, _ = _png_rows(got[f].tobytes(), 320, 200, 4)
        want = ((plain[f].astype(np.uint32) + 128) // 257).astype(np.uint8)
        want[..., 3] = 255                                        # cfg.transparent = False: alpha 65535
        assert np.array_equal(samples, want.reshape(200, 320 * 4)), f
    L = S._native.lib()
    assert L.sar_encoded_size(320, 200, 1, 4) == 0
    with pytest.raises(S.SarError):
        S.encode_image(r.runtime(), S.PixelFormat.Rgb16, S.Container.PngDeflate)
    r.shutdown()


def test_png_bound_and_variable_size_container_without_a_gpu():
    """Host-only parts of the compressed-PNG API: the worst case is every 16 KB block stored (5 bytes each) around the
    filtered scanlines; the variable-size container has no fixed-size form."""
    from strange_attractor_renderer_b200 import _native as N

    L = N.lib()
    for (w, h, fmt, bpp) in ((1920, 1080, 1, 6), (1, 1, 0, 8), (2048, 2048, 3, 3), (8300, 5, 2, 4)):
        raw = h * (1 + w * bpp)
        blocks = (raw + 16383) // 16384
        assert L.sar_png_bound(w, h, fmt) == 43 + raw + 5 * blocks + 20
    assert L.sar_png_bound(0, 10, 1) == 0 and L.sar_png_bound(10, 10, 9) == 0
    assert L.sar_encoded_size(64, 64, 1, N.SAR_FILE_PNG_DEFLATE) == 0
    assert L.sar_encode_header(64, 64, 1, N.SAR_FILE_PNG_DEFLATE, None, 0, None) == N.SAR_ERR_INVALID
