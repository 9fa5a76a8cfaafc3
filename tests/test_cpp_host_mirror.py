"""The C++ host mirror (include/sar.hpp) compiles against the C ABI and behaves like api.py."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _fnv(arr: np.ndarray) -> int:
    h = 1469598103934665603
    for b in arr.tobytes():
        h = ((h ^ b) * 1099511628211) & (2**64 - 1)
    return h


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    import strange_attractor_renderer_b200 as S

    if S.build.needs_build():
        S.build.build()
    out = tmp_path_factory.mktemp("cpp") / "host_mirror"
    libdir = os.path.dirname(S._native.LIB_PATH)
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "host_mirror.cpp"),
           "-o", str(out), "-L", libdir, "-l:libsar_b200.so", f"-Wl,-rpath,{libdir}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return str(out)


def test_cpp_mirror_compiles_and_fails_loudly_without_gpu(exe):
    import strange_attractor_renderer_b200 as S

    n = C.c_int(0)
    if S._native.lib().sar_device_count(C.byref(n)) == 0 and n.value > 0:
        pytest.skip("a GPU is visible; covered by the gpu test")
    r = subprocess.run([exe, "--no-gpu"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "PRESETS_OK" in r.stdout and "NO_GPU_ERROR code=-3" in r.stdout


@pytest.mark.gpu
def test_cpp_mirror_matches_python_api(exe):
    import strange_attractor_renderer_b200 as S

    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    lines = dict()
    for ln in r.stdout.splitlines():
        parts = ln.split()
        lines.setdefault(parts[0], []).append({k: v for k, v in (p.split("=") for p in parts[1:] if "=" in p)})
    assert lines["MERGE_MISMATCH"][0]["code"] == str(S._native.SAR_ERR_DIMS)
    # the same calls through api.py
    cfg = S.Config.poisson_saturne()
    cfg.iterations, cfg.width, cfg.height = 200000, 320, 200
    rt = S.Runtime.new(cfg, seed=42)
    for frame in range(2):
        cfg.angle = 0.3 * frame
        S.render(cfg, rt)
        img = S.colorize(cfg, rt)
        assert str(_fnv(img)) == lines["SINGLE"][frame]["hash"]
        rt.reset()
    sol = S.Config.solar_sail()
    sol.iterations, sol.width, sol.height, sol.angle = 3000000, 180, 200, 3.839724354387525
    pr = S.ParallelRenderer.new(threads=256)
    img = S.render_parallel(pr, sol, 12, seed=7)
    assert lines["PARALLEL"][0]["threads"] == "256"
    assert str(_fnv(img)) == lines["PARALLEL"][0]["hash"]
    pr.shutdown()
    # encoders and auto-framing through the C++ mirror == through api.py
    cfg.angle, cfg.transparent = 0.0, False
    S.render(cfg, rt)
    S.colorize(cfg, rt)
    pam = S.encode_image(rt, S.PixelFormat.of(False, False), S.Container.Pam)
    bmp = S.encode_image(rt, S.PixelFormat.of(False, True), S.Container.Bmp)
    png = S.encode_png(rt, S.PixelFormat.of(False, False))      # deterministic: same bytes from both hosts
    assert lines["ENCODED"][0] == {"pam": str(_fnv(pam)), "bmp": str(_fnv(bmp)), "png": str(_fnv(png))}
    assert lines["BMP16"][0]["code"] == str(S._native.SAR_ERR_UNSUPPORTED)
    # the frame loop through the C++ mirror == through api.py: same frames, same compressed PNG files
    seq = S.Config.solar_sail()
    seq.iterations, seq.width, seq.height, seq.transparent = 600000, 160, 120, False
    r2 = S.ParallelRenderer.new(threads=512)
    angles = S.angle_iter(0.0, 50.0, 10.0)
    frames = S.render_sequence(r2, seq, angles, 2, seed=5)
    pngs = S.render_sequence_encoded(r2, seq, angles, 2, S.PixelFormat.Rgb16, S.Container.PngDeflate, seed=5)
    r2.shutdown()
    got = lines["SEQUENCE"][0]
    assert got["frames"] == got["pngframes"] == str(len(angles)) == "5"
    assert got["hash"] == str(_fnv(frames)) and got["pngbytes"] == str(sum(p.size for p in pngs))
    assert got["pnghash"] == str(_fnv(np.concatenate(pngs)))
    af = S.autoframe(S.Config.poisson_saturne(), n_jobs=1024, iterations=2000, seed=3)
    assert float(lines["AUTOFRAME"][0]["xmin"]) == af.box[0] and float(lines["AUTOFRAME"][0]["ymax"]) == af.box[3]
    assert lines["AUTOFRAME"][0]["diverged"] == str(af.diverged)
