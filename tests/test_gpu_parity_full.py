"""GPU vs oracle, bit-exact, on the BASELINE.json configurations THEMSELVES (full size).

The oracle side uses its multithreaded mode (orc_render_jobs_mt: contiguous job slices into private
Runtimes merged in slice order — pinned to the serial job loop in tests/test_oracle_mt.py), so a
1e9-iteration frame costs a few seconds of host time.  Compared: count (u32), zbuf (f32), steps
(f64), max, and the RGBA16 image — all exact.  These are the cases that exercise the depth-test CAS
races at the benchmark's 132 608 lanes and the > 2^23-pixel (unscrambled) `fast` layout.
"""
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def S():
    import strange_attractor_renderer_b200 as S

    S._native.lib()
    return S


def _check_frame(S, oracle, cfg, seed, jobs_per_thread=1):
    """render_parallel with the library's default lanes (what bench.py times) vs the oracle on the
    same decomposition (lib.rs:1058-1062) and the same start points."""
    r = S.ParallelRenderer.new()
    n, per_job = r.plan(cfg.iterations, jobs_per_thread)
    img = S.render_parallel(r, cfg, jobs_per_thread, seed=seed)
    count, steps, zbuf, mx = r.runtime().download()
    ocfg = cfg.to_pod()
    ocfg.iterations = per_job
    ort = oracle.Runtime(cfg.width, cfg.height)
    st = oracle.OrcStats()
    oracle.render_jobs_mt(ocfg, ort, oracle.seed_points(seed, 0, n * jobs_per_thread), 0, st)
    assert np.array_equal(count, ort.count), f"count differs in {(count != ort.count).sum()} pixels"
    assert mx == ort.max
    assert np.array_equal(zbuf.view(np.uint32), ort.zbuf.view(np.uint32)), f"zbuf differs in {(zbuf != ort.zbuf).sum()} pixels"
    assert np.array_equal(steps.view(np.uint64), ort.steps.view(np.uint64)), f"steps differs in {(steps != ort.steps).sum()} pixels"
    oimg = oracle.colorize(ocfg, ort)
    d = img.astype(np.int32) - oimg.astype(np.int32)
    assert not d.any(), f"RGBA16 image differs in {(d != 0).sum()} values (max {np.abs(d).max()} LSB)"
    r.shutdown()
    return n, per_job, st, count, mx


def test_baseline_cfg1_poisson_1e9_2048_full_size(S, oracle):
    """BASELINE configs[1] — the benchmarked frame: poisson-saturne, 1e9 iterations, 2048x2048."""
    cfg = S.Config.poisson_saturne()
    cfg.width, cfg.height, cfg.iterations, cfg.transparent = 2048, 2048, 1_000_000_000, False
    n, per_job, st, count, mx = _check_frame(S, oracle, cfg, seed=1234)
    assert st.recorded == n * per_job == int(count.sum(dtype=np.uint64)) > 999_000_000
    assert mx < (1 << 20)


def test_baseline_cfg2_solar_sail_1e9_1800x2000_220deg_full_size(S, oracle):
    """BASELINE configs[2]: solar-sail, 1e9 iterations, 1800x2000, angle 220 degrees (radians,
    SURVEY §0.7), including the ~38 % of start points that diverge into the NaN sink (§0.5)."""
    cfg = S.Config.solar_sail()
    cfg.width, cfg.height, cfg.iterations, cfg.transparent = 1800, 2000, 1_000_000_000, False
    cfg.angle = 220.0 * math.pi / 180.0
    n, per_job, st, count, mx = _check_frame(S, oracle, cfg, seed=1234)
    assert st.nan_iters > 0.3 * n * per_job, "expected the diverging start points"
    assert count[0, 0] == mx >= (1 << 20), "the NaN sink pixel holds max, beyond the ln table (lib.rs:860 path with host max)"


def test_baseline_cfg0_single_trajectory_1e7_512(S, oracle):
    """BASELINE configs[0]: the reference's own CPU case — render() = ONE serial trajectory of 1e7
    chaotic steps at 512x512 (lib.rs:769-837)."""
    cfg = S.Config.poisson_saturne()
    cfg.width, cfg.height, cfg.iterations = 512, 512, 10_000_000
    pts = S.seed_points(2024, 0, 1)
    rt = S.Runtime.new(cfg)
    S.render(cfg, rt, initial_points=pts)
    count, steps, zbuf, mx = rt.download()
    ort = oracle.Runtime(512, 512)
    oracle.render_jobs(cfg.to_pod(), ort, pts)
    assert np.array_equal(count, ort.count) and mx == ort.max and int(count.sum()) == 10_000_000
    assert np.array_equal(zbuf.view(np.uint32), ort.zbuf.view(np.uint32))
    assert np.array_equal(steps.view(np.uint64), ort.steps.view(np.uint64))
    assert np.array_equal(S.colorize(cfg, rt), oracle.colorize(cfg.to_pod(), ort))


def test_more_than_2_pow_23_pixels_layout(S, oracle):
    """Above 2^23 pixels the `fast` array is not scrambled (slot = pixel) — the layout
    BASELINE configs[3] (4096x4096) runs on.  4096x2049 > 2^23 pixels, 2e8 iterations."""
    cfg = S.Config.poisson_saturne()
    cfg.width, cfg.height, cfg.iterations, cfg.transparent = 4096, 2049, 200_000_000, True
    assert cfg.width * cfg.height > (1 << 23)
    n, per_job, st, count, mx = _check_frame(S, oracle, cfg, seed=99)
    # the frame is wider than tall, so part of the attractor falls outside it (lib.rs:789)
    assert 0.5 * n * per_job < st.recorded == int(count.sum(dtype=np.uint64)) < n * per_job


def test_baseline_cfg4_sequence_frames_full_size(S, oracle):
    """BASELINE configs[4]: frames of the solar-sail angle sweep at full size (1e8 iterations per frame, 2048x2048,
    default lanes, fresh start points per frame as the reference draws them) through sar_render_sequence — every
    frame's RGBA16 image equals the oracle's colorize bit for bit although max (the NaN sink) is far beyond the
    host-libm ln table."""
    cfg = S.Config.solar_sail()
    cfg.width, cfg.height, cfg.iterations, cfg.transparent = 2048, 2048, 100_000_000, True
    angles = S.angle_iter(0.0, 360.0, 120.0)                      # 0, 120, 240 degrees of the 360-frame sweep
    r = S.ParallelRenderer.new()
    n, per_job = r.plan(cfg.iterations, 1)
    frames = S.render_sequence(r, cfg, angles, 1, seed=7)
    ocfg = cfg.to_pod()
    ocfg.iterations = per_job
    for f, a in enumerate(angles):
        ocfg.angle = a
        ort = oracle.Runtime(2048, 2048)
        oracle.render_jobs_mt(ocfg, ort, oracle.seed_points(7, f * n, n))
        assert ort.max >= (1 << 20)
        d = frames[f].astype(np.int32) - oracle.colorize(ocfg, ort).astype(np.int32)
        assert not d.any(), f"frame {f}: {(d != 0).sum()} values differ"
    r.shutdown()


@pytest.mark.parametrize("name", ["poisson_saturne", "solar_sail", "solar_sail_220"])
def test_gpu_render_against_the_reference_s_published_image(S, name):
    """GPU vs the REFERENCE's output, no oracle in between: the three README commands (README.md:72-77; 1e9 iterations,
    1920x1080 / 1800x2000 / 1800x2000 at 220 degrees) rendered by sar_render_parallel with the library's default
    decomposition; the count field must equal the one recovered from the published PNG (tests/test_reference_images.py)
    at Poisson noise per pixel.  Seeds differ (the reference's are unknowable), so this is a chi-square, not an equality:
    ~1.0 expected (the oracle with the same decomposition: 0.9975), a field one pixel off gives > 60.  For solar-sail the
    recorded mass is scaled by the share of bounded jobs (the rest sits in the NaN sink, pixel (0,0): 58 / 54 of the
    author's 144 jobs), and the palette positions are compared under the images' AdjustedVelocity constants."""
    import os

    import test_reference_images as T

    inv = np.load(os.path.join(T.ROOT, "tests", "golden", "media_inverse.npz"))
    preset, off, angle = T.IMAGES[name]
    cfg = getattr(S.Config, preset)()
    w, h = (1920, 1080) if name == "poisson_saturne" else (1800, 2000)
    cfg.iterations, cfg.width, cfg.height, cfg.transparent, cfg.angle = 1_000_000_000, w, h, False, angle
    cfg.colors.brighness.offset = off
    if preset == "solar_sail":
        cfg.color_transform = S.color_transforms.AdjustedVelocity(offset=-0.2, factor=0.8)   # see T._config
    r = S.ParallelRenderer.new()
    img = S.render_parallel(r, cfg, 1, seed=4321)
    count, steps, zbuf, mx = r.runtime().download()
    jobs, per_job = r.plan(cfg.iterations, 1)
    total = int(count.sum(dtype=np.uint64))
    idx, n, v = inv[name + "_idx"].astype(np.int64), inv[name + "_n"], inv[name + "_v"].astype(np.float64)
    c = count.ravel()
    rmax = int(inv[name + "_max"])
    if name == "poisson_saturne":
        assert total == jobs * per_job > 999_000_000                                          # fully in view
        s = 1.0
        assert abs(mx - rmax) < 5.0 * np.sqrt(float(mx))                                      # hottest pixel, 95 125
        lit = float((img[..., :3].max(axis=2) > 0).mean())
        assert abs(lit - 0.33337) < 2e-3, lit                                                 # lit fraction of the PNG
    else:
        sink = int(c[0])
        assert mx == sink and sink % per_job < per_job // 100                                 # whole jobs (+ a few stray hits)
        frac_ref, frac_gpu = rmax / (144.0 * T.PER_JOB), sink / float(jobs * per_job)
        assert abs(frac_ref - frac_gpu) < 4.0 * np.sqrt(0.39 * 0.61 / 144.0), (frac_ref, frac_gpu)   # diverging share
        s = (144.0 * T.PER_JOB - rmax) / float(total - sink)
    chi = T._chi2(n, c[idx], s)
    assert 0.85 < chi < 1.25, chi
    for shift in (1, -1, w, -w):
        assert T._chi2(n, c[np.clip(idx + shift, 0, c.size - 1)], s) > 15.0
    assert abs(float(n.sum()) / (s * float(c[idx].sum())) - 1.0) < 2e-3
    o = steps.ravel()[idx]
    keep = (v < 5.0 / 6.0 - 1e-3) & (o < 5.0 / 6.0 - 1e-3) & (c[idx] > 0)
    assert np.median(np.abs(v[keep] - o[keep])) < 1e-4
    r.shutdown()
