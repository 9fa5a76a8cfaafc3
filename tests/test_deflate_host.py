"""The compressor of the PNG writer (csrc/sar_deflate.cuh) on the CPU: the inline functions the CUDA kernel calls are
compiled into a host harness (tests/cpp/deflate_host.cpp, the 32 lanes of a warp emulated by a loop) and every stream
they produce must inflate — with zlib, the reference decoder — to the input, byte for byte.  The device kernel itself is
checked the same way on the GPU (tests/test_encoders.py::test_device_png_deflate_decodes_to_the_reference_pixels)."""
import ctypes as C
import os
import subprocess
import zlib

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def dfl(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("dfl") / "libdfl_host.so")
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-Wall", "-Werror", "-o", so,
                    os.path.join(ROOT, "tests", "cpp", "deflate_host.cpp")], check=True)
    L = C.CDLL(so)
    L.dfl_compress_host.restype = C.c_size_t
    L.dfl_compress_host.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p]
    L.dfl_code_lengths_host.argtypes = [C.c_void_p, C.c_void_p]
    L.dfl_chunk_bytes.restype = C.c_uint32
    return L


def _compress(L, data: bytes):
    raw = np.frombuffer(data, np.uint8)
    chunk = L.dfl_chunk_bytes()
    cap = len(raw) + 5 * (len(raw) // chunk + 2) + 64
    out = np.zeros(cap, np.uint8)
    kinds = np.zeros(len(raw) // chunk + 2, np.uint8)
    n = L.dfl_compress_host(raw.ctypes.data, len(raw), out.ctypes.data, cap, kinds.ctypes.data)
    assert n > 0
    return out[:n].tobytes(), kinds[:(len(raw) + chunk - 1) // chunk]


def test_streams_inflate_to_the_input(dfl):
    rng = np.random.default_rng(0)
    chunk = dfl.dfl_chunk_bytes()
    cases = {
        "zeros": bytes(100_000), "one": b"\x07", "two": b"ab", "run3": b"aaa", "run258": b"x" * 259, "run259": b"x" * 260,
        "noise": rng.integers(0, 256, 70_000, dtype=np.uint8).tobytes(),
        "runs": np.repeat(rng.integers(0, 256, 3000, dtype=np.uint8), rng.integers(1, 600, 3000)).tobytes(),
        "skewed": rng.choice(np.arange(256, dtype=np.uint8), 50_000, p=(lambda p: p / p.sum())(0.5 ** np.arange(256))).tobytes(),
        "chunk": bytes(chunk), "chunk+1": bytes(chunk + 1), "chunk-1": bytes([3]) * (chunk - 1),
        "text": b"hello world " * 5000,
        "lane_seams": (bytes([1]) * 511 + bytes([2]) * 3 + bytes([1]) * 700) * 20,    # runs crossing the 512-byte lane ranges
    }
    for name, data in cases.items():
        comp, kinds = _compress(dfl, data)
        assert zlib.decompress(comp, wbits=-15) == data, name
        # never larger than storing every block
        assert len(comp) <= len(data) + 5 * max(1, len(kinds)), name
    comp, kinds = _compress(dfl, cases["noise"])
    assert not kinds.any(), "incompressible blocks must fall back to stored"
    comp, kinds = _compress(dfl, cases["zeros"])
    assert kinds.all() and len(comp) < 2000


def test_random_inputs_round_trip(dfl):
    """Mixed run / literal content with random alphabets and lengths around the block and lane sizes."""
    rng = np.random.default_rng(42)
    chunk = dfl.dfl_chunk_bytes()
    for trial in range(60):
        n = int(rng.choice([1, 2, 3, 5, 257, 258, 259, 511, 512, 513, chunk - 1, chunk, chunk + 1, 3 * chunk + 77,
                            int(rng.integers(1, 6 * chunk))]))
        alphabet = int(rng.choice([1, 2, 3, 16, 256]))
        mean_run = float(rng.choice([1.0, 1.5, 4.0, 40.0, 400.0]))
        vals = rng.integers(0, alphabet, n, dtype=np.uint8)
        runs = rng.geometric(1.0 / mean_run, n)
        data = np.repeat(vals, runs)[:n].tobytes()
        comp, _ = _compress(dfl, data)
        assert zlib.decompress(comp, wbits=-15) == data, (trial, n, alphabet, mean_run)


def test_code_lengths_are_optimal_complete_and_capped(dfl):
    """Kraft sum exactly 1 (zlib rejects incomplete literal codes), at most 15 bits, and — when the cap does not bind —
    the cost of an optimal prefix code (compared with a heap-built Huffman code)."""
    import heapq

    rng = np.random.default_rng(3)

    def lengths(freq):
        f = np.ascontiguousarray(freq, np.uint32)
        out = np.zeros(286, np.uint8)
        dfl.dfl_code_lengths_host(f.ctypes.data, out.ctypes.data)
        return out

    def optimal_cost(freq):
        h = [int(f) for f in freq if f]
        heapq.heapify(h)
        cost = 0
        while len(h) > 1:
            a, b = heapq.heappop(h), heapq.heappop(h)
            cost += a + b
            heapq.heappush(h, a + b)
        return cost

    for trial in range(200):
        k = int(rng.integers(2, 287))
        freq = np.zeros(286, np.uint32)
        used = rng.choice(286, k, replace=False)
        shape = rng.choice(["flat", "geometric", "zipf"])
        if shape == "flat":
            freq[used] = rng.integers(1, 50, k)
        elif shape == "geometric":
            freq[used] = np.maximum(1, (16000 * 0.7 ** np.arange(k)).astype(np.uint32))
        else:
            freq[used] = np.maximum(1, (16000 / (1 + np.arange(k)) ** 1.5).astype(np.uint32))
        ln = lengths(freq)
        assert ((ln > 0) == (freq > 0)).all()
        assert ln.max() <= 15
        assert sum(2.0 ** -int(v) for v in ln if v) == 1.0
        if optimal_cost(freq) == int((ln.astype(np.int64) * freq).sum()):
            continue
        # the cap was binding: a Fibonacci-like tail; cost may exceed the optimum only then
        assert shape != "flat"
    # Fibonacci frequencies force a depth beyond 15 without the cap
    fib = [1, 1]
    while len(fib) < 30:
        fib.append(fib[-1] + fib[-2])
    freq = np.zeros(286, np.uint32)
    freq[:30] = fib
    ln = lengths(freq)
    assert ln.max() == 15 and sum(2.0 ** -int(v) for v in ln if v) == 1.0


def test_filtered_frame_compresses_like_the_reference_file(dfl, oracle):
    """A real frame: the oracle's poisson-saturne render, RGB16 big-endian scanlines, Sub filter — the stream the device
    compresses.  Round trip, and a ratio in the range of the reference's own file (media/poisson-saturne.png: 0.29)."""
    cfg = oracle.poisson_saturne()
    cfg.width, cfg.height, cfg.iterations, cfg.transparent, cfg.bright_offset = 480, 270, 20_000_000, 0, -0.25
    img = oracle.render_parallel(cfg, 8, 12, oracle.seed_points(5, 0, 96))[..., :3]
    be = img.astype(">u2").view(np.uint8).reshape(270, 480 * 6).astype(np.int16)
    sub = be.copy()
    sub[:, 6:] -= be[:, :-6]
    rows = np.zeros((270, 1 + 480 * 6), np.uint8)
    rows[:, 0] = 1
    rows[:, 1:] = (sub & 255).astype(np.uint8)
    data = rows.tobytes()
    comp, kinds = _compress(dfl, data)
    assert zlib.decompress(comp, wbits=-15) == data
    assert len(comp) < 0.6 * len(data)
