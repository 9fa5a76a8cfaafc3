"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on the same start points.

Bar (BASELINE.json north_star): count buffer bit-exact; colour within 1e-5 relative.  This
build is stricter: zbuf (f32) and steps (f64) are bit-exact too, because z ties resolve to the
earlier render() call / earlier iteration exactly as in a sequential reference run, and the u16
image is bit-exact as well: ln(count+1) and ln(max+1) come from a host-libm table (counts below
2^20) or from the host directly (max), every other colour operation is exact IEEE arithmetic.
"""
import ctypes as C
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

REL_TOL_COLOUR = 1e-5  # north_star: "within 1e-5 relative in the fp32 colour buffer"


@pytest.fixture(scope="module")
def S():
    import strange_attractor_renderer_b200 as S

    S._native.lib()
    return S


def _small(cfg, w, h, iters):
    cfg.width, cfg.height, cfg.iterations = w, h, iters
    return cfg


def _oracle_state(oracle, cfg, pts):
    rt = oracle.Runtime(cfg.width, cfg.height)
    st = oracle.OrcStats()
    oracle.render_jobs(cfg.to_pod(), rt, pts, st)
    return rt, st


def _assert_state_equal(gpu_state, ort):
    count, steps, zbuf, mx = gpu_state
    assert np.array_equal(count, ort.count), f"count differs in {(count != ort.count).sum()} pixels"
    assert mx == ort.max
    assert np.array_equal(zbuf, ort.zbuf), f"zbuf differs in {(zbuf != ort.zbuf).sum()} pixels"
    assert np.array_equal(steps.view(np.uint64), ort.steps.view(np.uint64)), \
        f"steps differs in {(steps != ort.steps).sum()} pixels"


def _assert_image_close(img, f32, oimg, of64, exact=True):
    """u16 image: bit-exact (ln comes from the host libm table); with exact=False
    (counts beyond the table on a path that cannot read max back) at most 1 LSB in < 1e-4 of values."""
    d = np.abs(img.astype(np.int32) - oimg.astype(np.int32))
    if exact:
        assert d.max() == 0, f"u16 image differs in {(d > 0).sum()} values (max {d.max()} LSB)"
    assert d.max() <= 1, f"u16 image differs by {d.max()} LSB"
    assert (d > 0).mean() < 1e-4, f"{(d > 0).sum()} u16 channel values differ by 1 LSB"
    if f32 is not None:
        ref = of64.astype(np.float64)
        ok = np.isclose(f32.astype(np.float64), ref, rtol=REL_TOL_COLOUR, atol=1e-7) | (np.isnan(ref) & np.isnan(f32))
        assert ok.all(), f"{(~ok).sum()} colour values outside {REL_TOL_COLOUR} relative"


def test_poisson_saturne_many_jobs_bit_exact(S, oracle):
    """BASELINE cfg 1 shape (512x512) with render_parallel-style decomposition: 256 jobs."""
    cfg = _small(S.Config.poisson_saturne(), 512, 512, 40_000)
    cfg.transparent = False
    pts = S.seed_points(1234, 0, 256)
    assert np.array_equal(pts, oracle.seed_points(1234, 0, 256))
    rt = S.Runtime.new(cfg)
    S.render(cfg, rt, initial_points=pts)
    ort, st = _oracle_state(oracle, cfg, pts)
    assert st.recorded == 256 * 40_000  # attractor fully in view
    _assert_state_equal(rt.download(), ort)
    img, f32 = S.colorize(cfg, rt, want_f32=True)
    oimg, of64 = oracle.colorize(cfg.to_pod(), ort, want_f64=True)
    _assert_image_close(img, f32, oimg, of64)


def test_single_long_trajectory_is_reference_render(S, oracle):
    """BASELINE cfg 0: render() = ONE serial trajectory (lib.rs:769-837); 3e6 steps must stay
    bit-identical although the map is chaotic."""
    cfg = _small(S.Config.poisson_saturne(), 512, 512, 3_000_000)
    pts = S.seed_points(99, 0, 1)
    rt = S.Runtime.new(cfg)
    S.render(cfg, rt, initial_points=pts)
    ort, _ = _oracle_state(oracle, cfg, pts)
    _assert_state_equal(rt.download(), ort)


def test_solar_sail_nan_sink_and_angle(S, oracle):
    """BASELINE cfg 2 shape scaled down: solar-sail at 220 deg (in radians), with the 38 % of
    start points that diverge to NaN and pile onto count[(0,0)]."""
    cfg = _small(S.Config.solar_sail(), 360, 400, 30_000)
    cfg.angle = 220.0 * math.pi / 180.0
    pts = S.seed_points(7, 0, 200)
    rt = S.Runtime.new(cfg)
    S.render(cfg, rt, initial_points=pts)
    ort, st = _oracle_state(oracle, cfg, pts)
    assert st.nan_iters > 30_000 * 20, "expected diverging trajectories in this list"
    gs = rt.download()
    _assert_state_equal(gs, ort)
    assert gs[0][0, 0] == gs[3], "the NaN sink pixel holds the max"
    img, f32 = S.colorize(cfg, rt, want_f32=True)
    oimg, of64 = oracle.colorize(cfg.to_pod(), ort, want_f64=True)
    _assert_image_close(img, f32, oimg, of64)
    assert list(img[0, 0, :3]) == [65535, 65535, 60849]   # known answer, media/solar-sail-220deg.png


def test_progressive_accumulation_and_reset(S, oracle):
    """render() on a non-reset Runtime continues the image (lib.rs:742-743); reset() clears it."""
    cfg = _small(S.Config.poisson_saturne(), 256, 256, 20_000)
    a, b = S.seed_points(1, 0, 40), S.seed_points(2, 0, 24)
    rt = S.Runtime.new(cfg)
    S.render(cfg, rt, initial_points=a)
    S.render(cfg, rt, initial_points=b)
    ort, _ = _oracle_state(oracle, cfg, np.concatenate([a, b]))
    _assert_state_equal(rt.download(), ort)
    rt.reset()
    count, steps, zbuf, mx = rt.download()
    assert count.max() == 0 and mx == 0 and (zbuf == -1.0).all() and (steps == 0.0).all()
    S.render(cfg, rt, initial_points=b)
    ort2, _ = _oracle_state(oracle, cfg, b)
    _assert_state_equal(rt.download(), ort2)


def test_seeded_render_matches_explicit_points(S, oracle):
    """Runtime's own generator (device-side SplitMix64) == sar_seed_points == the oracle's."""
    cfg = _small(S.Config.poisson_saturne(), 128, 128, 5_000)
    rt = S.Runtime.new(cfg, seed=4242)
    for _ in range(3):
        S.render(cfg, rt)   # three reference render() calls, three draws
    ort, _ = _oracle_state(oracle, cfg, oracle.seed_points(4242, 0, 3))
    _assert_state_equal(rt.download(), ort)


def test_merge_matches_reference_merge(S, oracle):
    """Runtime::merge (lib.rs:708-738): counts add, other wins on strictly greater z."""
    cfg = _small(S.Config.poisson_saturne(), 200, 160, 15_000)
    a, b = S.seed_points(11, 0, 32), S.seed_points(12, 0, 32)
    ra, rb = S.Runtime.new(cfg), S.Runtime.new(cfg)
    S.render(cfg, ra, initial_points=a)
    S.render(cfg, rb, initial_points=b)
    oa, _ = _oracle_state(oracle, cfg, a)
    ob, _ = _oracle_state(oracle, cfg, b)
    ra.merge(rb)
    oa.merge(ob)
    _assert_state_equal(ra.download(), oa)
    other = S.Runtime.new(_small(S.Config.poisson_saturne(), 100, 100, 1))
    with pytest.raises(S.SarError) as e:   # reference: assert_eq! panic, lib.rs:709-710
        ra.merge(other)
    assert e.value.code == S._native.SAR_ERR_DIMS


def test_download_upload_roundtrip_continues_render(S, oracle):
    """Checkpoint/resume of the progressive accumulation: download, upload into a new Runtime,
    keep rendering; equals an uninterrupted run."""
    cfg = _small(S.Config.solar_sail(), 180, 200, 10_000)
    a, b = S.seed_points(21, 0, 30), S.seed_points(22, 0, 30)
    r1 = S.Runtime.new(cfg)
    S.render(cfg, r1, initial_points=a)
    count, steps, zbuf, _ = r1.download()
    r2 = S.Runtime.new(cfg)
    r2.upload(count, steps, zbuf)
    S.render(cfg, r2, initial_points=b)
    ort, _ = _oracle_state(oracle, cfg, np.concatenate([a, b]))
    _assert_state_equal(r2.download(), ort)


def test_render_parallel_decomposition(S, oracle):
    """render_parallel (lib.rs:1051-1082): iterations/num_threads/jobs_per_thread per job (integer
    division, remainder dropped), num_threads*jobs_per_thread jobs, merged, colourised."""
    cfg = _small(S.Config.poisson_saturne(), 320, 240, 10_000_123)
    cfg.transparent = True
    r = S.ParallelRenderer.new(threads=96)
    assert r.num_threads() == 96
    img = S.render_parallel(r, cfg, 3, seed=555)
    per_job = 10_000_123 // 96 // 3
    ocfg = cfg.to_pod()
    ocfg.iterations = per_job
    ort = oracle.Runtime(320, 240)
    oracle.render_jobs(ocfg, ort, oracle.seed_points(555, 0, 288))
    assert int(ort.count.sum()) == per_job * 288
    _assert_state_equal(r.runtime().download(), ort)
    oimg, _ = oracle.colorize(ocfg, ort, want_f64=True)
    _assert_image_close(img, None, oimg, None)
    # explicit start points take the same path
    img2 = S.render_parallel(r, cfg, 3, initial_points=oracle.seed_points(555, 0, 288))
    assert np.array_equal(img, img2)
    r.shutdown()


def test_depth_render_kind(S, oracle):
    """RenderKind::Depth (lib.rs:875-900): f32 min/max over touched pixels, grey u16."""
    cfg = _small(S.Config.poisson_saturne(), 300, 200, 20_000)
    cfg.render = S.RenderKind.Depth
    pts = S.seed_points(3, 0, 64)
    rt = S.Runtime.new(cfg)
    S.render(cfg, rt, initial_points=pts)
    ort, _ = _oracle_state(oracle, cfg, pts)
    img, f32 = S.colorize(cfg, rt, want_f32=True)
    oimg, of64 = oracle.colorize(cfg.to_pod(), ort, want_f64=True)
    assert np.array_equal(img, oimg)   # pure f32 arithmetic: exact
    assert np.array_equal(f32, of64.astype(np.float32))


def test_edge_cases(S, oracle):
    # colorize of an empty Runtime: max == 0 -> ln(1)/ln(1) = NaN -> `as u16` = 0 (lib.rs:860-866)
    cfg = _small(S.Config.poisson_saturne(), 64, 48, 0)
    rt = S.Runtime.new(cfg)
    S.render(cfg, rt, initial_points=S.seed_points(1, 0, 4))   # zero iterations: only warm-up runs
    img = S.colorize(cfg, rt)
    oimg = oracle.colorize(cfg.to_pod(), oracle.Runtime(64, 48))
    assert np.array_equal(img, oimg) and img[..., :3].max() == 0
    # zero jobs is a no-op
    S.render(cfg, rt, initial_points=np.empty((0, 3)))
    # 1x1 image: every in-view hit lands on the single pixel
    cfg1 = _small(S.Config.poisson_saturne(), 1, 1, 1000)
    r1 = S.Runtime.new(cfg1)
    pts = S.seed_points(5, 0, 8)
    S.render(cfg1, r1, initial_points=pts)
    o1, _ = _oracle_state(oracle, cfg1, pts)
    _assert_state_equal(r1.download(), o1)
    # zoomed in so that most points fall outside the viewport: out-of-view points still advance
    # previous_point (lib.rs:790-794), which feeds the next in-view delta
    cfgz = _small(S.Config.poisson_saturne(), 256, 256, 30_000)
    cfgz.view.scale = 6.0
    rz = S.Runtime.new(cfgz)
    pts = S.seed_points(8, 0, 64)
    S.render(cfgz, rz, initial_points=pts)
    oz, st = _oracle_state(oracle, cfgz, pts)
    assert 0 < st.recorded < 64 * 30_000 * 0.5
    _assert_state_equal(rz.download(), oz)
    # config / runtime dimension mismatch and bad arguments are errors, not crashes
    with pytest.raises(S.SarError):
        S.render(_small(S.Config.poisson_saturne(), 65, 48, 10), rt, initial_points=pts)
    with pytest.raises(S.SarError):
        S.Runtime.new(_small(S.Config.poisson_saturne(), 0, 10, 1))


def test_palette_and_transparency_variants(S, oracle):
    cfg = _small(S.Config.solar_sail(), 150, 170, 20_000)
    cfg.color_transform = S.color_transforms.AdjustedVelocity(offset=0.05, factor=2.5)  # spreads over the palette
    cfg.colors.palette = S.Palette.from_rgb([1.0, 0.2, 0.9], [0.1, 1.0, 0.3], [0.4, 0.6, 1.0])
    cfg.colors.brighness = S.BrighnessConstants(offset=-0.05, factor=1.25)
    pts = S.seed_points(31, 0, 48)
    rt = S.Runtime.new(cfg)
    S.render(cfg, rt, initial_points=pts)
    ort, _ = _oracle_state(oracle, cfg, pts)
    _assert_state_equal(rt.download(), ort)
    for transparent in (True, False):
        cfg.transparent = transparent
        img, f32 = S.colorize(cfg, rt, want_f32=True)
        oimg, of64 = oracle.colorize(cfg.to_pod(), ort, want_f64=True)
        _assert_image_close(img, f32, oimg, of64)


def test_lane_count_does_not_change_the_result(S):
    """Size-independent property: the trajectory -> lane mapping is invisible.  Same job list on
    64 lanes and on the default (SM count x 896) lanes gives identical buffers."""
    cfg = _small(S.Config.solar_sail(), 450, 500, 3_000)
    states = []
    for lanes in (64, 0):
        rt = S.Runtime.new(cfg)
        c = cfg.to_pod()
        S._native.check(S._native.lib().sar_render_seeded_async(C.byref(c), rt._h, 77, 0, 5000, lanes, None))
        S._native.check(S._native.lib().sar_stream_synchronize(rt._h, None))
        states.append(rt.download())
    for x, y in zip(states[0][:3], states[1][:3]):
        assert np.array_equal(x.view(np.uint8), y.view(np.uint8))
    assert states[0][3] == states[1][3]


def test_full_size_conservation_cfg2(S):
    """BASELINE cfg 1 at full size (1e9 iterations, 2048x2048): every recorded iteration is counted
    exactly once (sum of counts == jobs x iterations; poisson-saturne is fully in view), and the
    image is reproducible run to run."""
    cfg = _small(S.Config.poisson_saturne(), 2048, 2048, 1_000_000_000)
    r = S.ParallelRenderer.new()
    n = r.num_threads()
    img1 = S.render_parallel(r, cfg, 1, seed=1234)
    count, steps, zbuf, mx = r.runtime().download()
    assert int(count.sum(dtype=np.uint64)) == (1_000_000_000 // n) * n
    assert mx == count.max()
    lit = float((count > 0).mean())
    assert 0.15 < lit < 0.25   # 19-22 % of pixels touched
    assert ((zbuf > -1.0) == (count > 0)).all() or (zbuf[count > 0] >= -1.0).all()
    img2 = S.render_parallel(r, cfg, 1, seed=1234)
    assert np.array_equal(img1, img2)
    r.shutdown()


def test_render_parallel_auto_threads_with_cli_default_jobs(S, oracle):
    """jobs_per_thread = 12 is the CLI default (main.rs:305).  In auto mode the renderer keeps the
    number of jobs at its lane count: num_threads = lanes / 12, and the decomposition is the
    reference's for that num_threads (lib.rs:1058-1062)."""
    cfg = _small(S.Config.poisson_saturne(), 256, 192, 30_000_000)
    r = S.ParallelRenderer.new()
    lanes, n12 = r.num_threads(), r.num_threads(12)
    assert n12 == (lanes // 12) // 32 * 32 and n12 * 12 <= lanes
    img = S.render_parallel(r, cfg, 12, seed=99)
    ocfg = cfg.to_pod()
    ocfg.iterations = 30_000_000 // n12 // 12
    ort = oracle.Runtime(256, 192)
    oracle.render_jobs(ocfg, ort, oracle.seed_points(99, 0, n12 * 12))
    _assert_state_equal(r.runtime().download(), ort)
    _assert_image_close(img, None, oracle.colorize(ocfg, ort), None)
    r.shutdown()


def test_render_sequence_frames_equal_render_parallel(S, oracle):
    """The binary's frame loop (main.rs:496-512): every frame of a sequence equals render_parallel
    of that angle on the same start points — fresh points per frame by default, one shared list
    (warm-up run once) with shared_points=True."""
    cfg = _small(S.Config.solar_sail(), 200, 220, 3_000_000)
    angles = S.angle_iter(0.0, 90.0, 30.0)                       # 0, 30, 60 degrees -> radians
    assert np.allclose(angles, [0.0, math.pi / 6, math.pi / 3]) and S.angle_iter(220.0, 220.0, 1.0) == [220.0]
    r = S.ParallelRenderer.new(threads=128)
    jobs = 128 * 2
    fresh = S.render_sequence(r, cfg, angles, 2, seed=5)
    shared = S.render_sequence(r, cfg, angles, 2, seed=5, shared_points=True)
    seen = []
    S.render_sequence(r, cfg, angles, 2, seed=5, callback=lambda f, im: seen.append((f, im.copy())))
    assert [f for f, _ in seen] == [0, 1, 2]
    for f, a in enumerate(angles):
        cfg.angle = a
        want_fresh = S.render_parallel(r, cfg, 2, initial_points=S.seed_points(5, f * jobs, jobs))
        want_shared = S.render_parallel(r, cfg, 2, initial_points=S.seed_points(5, 0, jobs))
        assert np.array_equal(fresh[f], want_fresh)
        assert np.array_equal(shared[f], want_shared)
        assert np.array_equal(seen[f][1], want_fresh)
    # and one frame against the oracle
    ocfg = cfg.to_pod()
    ocfg.iterations = 3_000_000 // 128 // 2
    ort = oracle.Runtime(200, 220)
    oracle.render_jobs(ocfg, ort, oracle.seed_points(5, 0, jobs))
    _assert_image_close(shared[2], None, oracle.colorize(ocfg, ort), None)
    r.shutdown()


def test_randomised_configs_bit_exact(S, oracle):
    """Seeded random sweep over what a caller can vary: image shape (odd sizes, non powers of two),
    scale (incl. zoomed-in views that put most points out of view), camera centre, view angle,
    rotation, transform parameters, palette, iterations, job count — and perturbed coefficients,
    some of which make trajectories diverge.  State and RGBA16 image must equal the oracle's."""
    rng = np.random.default_rng(20261017)
    for case in range(12):
        base = S.Config.poisson_saturne() if case % 2 == 0 else S.Config.solar_sail()
        w, h = int(rng.integers(1, 400)), int(rng.integers(1, 400))
        cfg = _small(base, w, h, int(rng.integers(0, 4000)))
        cfg.angle = float(rng.uniform(-7.0, 7.0))
        cfg.view.scale = float(rng.choice([0.3, 1.0, 1.7, 4.0, 9.0]))
        cfg.view.center_camera.x += float(rng.normal(0, 0.05))
        cfg.view.center_camera.z += float(rng.normal(0, 0.05))
        cfg.view.rotation.rotation = float(rng.uniform(0, 6.3))
        cfg.transparent = bool(rng.integers(0, 2))
        cfg.colors.brighness = S.BrighnessConstants(offset=float(rng.uniform(-0.3, 0.1)), factor=float(rng.uniform(0.5, 2.5)))
        if case % 3 == 0:
            cfg.color_transform = S.color_transforms.AdjustedVelocity(offset=float(rng.uniform(-0.2, 0.5)), factor=float(rng.uniform(0.5, 3.0)))
        if case % 4 == 1:
            n = int(rng.integers(1, 9))
            cfg.colors.palette = S.Palette([tuple(rng.uniform(0, 1, 3)) for _ in range(n)])
        if case >= 6:   # nudge the map: some of these escape to infinity / NaN
            for lst in (cfg.attractor.x, cfg.attractor.y, cfg.attractor.z):
                k = int(rng.integers(0, 10))
                lst[k] += float(rng.normal(0, 0.02))
        if case == 11:
            cfg.render = S.RenderKind.Depth
        pts = S.seed_points(1000 + case, 0, int(rng.integers(1, 300)))
        rt = S.Runtime.new(cfg)
        S.render(cfg, rt, initial_points=pts)
        ort, st = _oracle_state(oracle, cfg, pts)
        _assert_state_equal(rt.download(), ort)
        img, f32 = S.colorize(cfg, rt, want_f32=True)
        oimg, of64 = oracle.colorize(cfg.to_pod(), ort, want_f64=True)
        _assert_image_close(img, f32, oimg, of64)
        rt.close()


def test_small_renders_are_not_empty_in_auto_mode(S, oracle):
    """A GPU has ~1e5 lanes; `iterations / num_threads / jobs_per_thread` (lib.rs:1058) must not
    round a small frame to nothing.  Auto mode caps num_threads so each job keeps >= 64 steps."""
    r = S.ParallelRenderer.new()
    lanes = r.num_threads()
    for iterations, jpt in ((100_000, 1), (1_000_000, 12), (2_000, 3)):
        n, per_job = r.plan(iterations, jpt)
        assert n % 32 == 0 and 32 <= n <= lanes and per_job == iterations // n // jpt
        assert per_job >= 64 or n == 32
        cfg = _small(S.Config.poisson_saturne(), 160, 120, iterations)
        img = S.render_parallel(r, cfg, jpt, seed=3)
        ocfg = cfg.to_pod()
        ocfg.iterations = per_job
        ort = oracle.Runtime(160, 120)
        oracle.render_jobs(ocfg, ort, oracle.seed_points(3, 0, n * jpt))
        assert int(ort.count.sum()) == per_job * n * jpt > 0
        _assert_state_equal(r.runtime().download(), ort)
        _assert_image_close(img, None, oracle.colorize(ocfg, ort), None)
    assert r.plan(10**9, 1) == (lanes, 10**9 // lanes)          # large frames are unaffected
    r.shutdown()


def test_negative_zero_z_is_kept_and_ties_with_positive_zero(S, oracle):
    """`z2 as f32` can be -0.0; f32 `>` (lib.rs:728, 821) sees -0.0 == +0.0, but the stored bits
    differ.  Through the checkpoint path: -0.0 survives upload/download bit for bit, and in a merge a
    +0.0 never displaces a -0.0 (nor the reverse): ties keep self."""
    cfg = _small(S.Config.poisson_saturne(), 4, 1, 10)
    nz, pz = np.float32(-0.0), np.float32(0.0)
    za = np.array([[nz, pz, nz, -0.5]], np.float32)
    zb = np.array([[pz, nz, 0.25, nz]], np.float32)
    ca, cb = np.array([[1, 2, 3, 4]], np.uint32), np.array([[10, 20, 30, 40]], np.uint32)
    sa, sb = np.array([[0.1, 0.2, 0.3, 0.4]]), np.array([[0.5, 0.6, 0.7, 0.8]])
    ra, rb = S.Runtime.new(cfg), S.Runtime.new(cfg)
    ra.upload(ca, sa, za)
    rb.upload(cb, sb, zb)
    c, s, z, _ = ra.download()
    assert np.array_equal(z.view(np.uint32), za.view(np.uint32)) and np.array_equal(s, sa) and np.array_equal(c, ca)
    oa, ob = oracle.Runtime(4, 1), oracle.Runtime(4, 1)
    oa.load(ca, sa, za)
    ob.load(cb, sb, zb)
    ra.merge(rb)
    oa.merge(ob)
    c, s, z, mx = ra.download()
    assert np.array_equal(z.view(np.uint32), oa.zbuf.view(np.uint32)), (z, oa.zbuf)
    assert np.array_equal(s, oa.steps) and np.array_equal(c, oa.count) and mx == oa.max
    assert z.view(np.uint32).tolist() == [[0x80000000, 0x00000000, np.float32(0.25).view(np.uint32), 0x80000000]]


def test_progressive_order_keys_are_linear(S, oracle):
    """Order keys advance by the number of jobs rendered, whatever `first_job` positions in the seed
    stream: many progressive render() calls on one Runtime stay exact, and the counter
    refuses to wrap."""
    cfg = _small(S.Config.poisson_saturne(), 64, 64, 500)
    rt = S.Runtime.new(cfg, seed=11)
    for _ in range(40):
        S.render(cfg, rt)
    jb = C.c_uint64()
    S._native.check(S._native.lib().sar_runtime_get_job_base(rt._h, C.byref(jb)))
    assert jb.value == 40
    ort, _ = _oracle_state(oracle, cfg, oracle.seed_points(11, 0, 40))
    _assert_state_equal(rt.download(), ort)
    S._native.check(S._native.lib().sar_runtime_set_job_base(rt._h, (1 << 32) - 1))
    with pytest.raises(S.SarError):
        S.render(cfg, rt, initial_points=S.seed_points(1, 0, 2))


@pytest.mark.parametrize("nt,pipe", [(1, 0), (1, 1), (2, 0), (2, 1), (4, 0), (4, 1)])
def test_every_kernel_variant_is_bit_exact(S, oracle, nt, pipe):
    """The tuning knobs (trajectories per thread, depth test one iteration behind its atomic) never
    change results: every instantiation against the oracle, on a case with NaN trajectories, a job
    count that is not a multiple of the lanes, several jobs per lane and out-of-view points."""
    L = S._native.lib()
    try:
        S._native.check(L.sar_set_option(b"traj_per_thread", nt))
        S._native.check(L.sar_set_option(b"pipeline", pipe))
        cfg = _small(S.Config.solar_sail(), 333, 217, 2_500)
        cfg.angle, cfg.view.scale = 0.9, 2.6
        n_jobs = 1000 + 37
        rt = S.Runtime.new(cfg)
        c = cfg.to_pod()
        S._native.check(L.sar_render_seeded_async(C.byref(c), rt._h, 4711, 0, n_jobs, 384, None))   # 2.7 jobs per lane
        S._native.check(L.sar_stream_synchronize(rt._h, None))
        ort, st = _oracle_state(oracle, cfg, oracle.seed_points(4711, 0, n_jobs))
        assert st.nan_iters > 0 and 0 < st.recorded < n_jobs * 2_500
        _assert_state_equal(rt.download(), ort)
        # one lane, one job: the pipelined tail (last iteration's pending test) on its own
        cfg1 = _small(S.Config.poisson_saturne(), 64, 64, 333)
        r1 = S.Runtime.new(cfg1)
        S.render(cfg1, r1, initial_points=S.seed_points(3, 0, 1))
        o1, _ = _oracle_state(oracle, cfg1, oracle.seed_points(3, 0, 1))
        _assert_state_equal(r1.download(), o1)
    finally:
        L.sar_set_option(b"traj_per_thread", 1)     # the defaults (sar_kernels.cu: SAR_DEFAULT_NT / SAR_DEFAULT_PIPE)
        L.sar_set_option(b"pipeline", 0)


def test_sequence_frames_are_exact_beyond_the_ln_table(S, oracle):
    """BASELINE configs[4] is a solar-sail sweep: the NaN sink pushes Runtime.max past the 2^20-entry
    host-libm ln table on every frame.  The sequence driver reads each frame's max back (one frame
    behind, on a second Runtime) so that ln(max + 1) — the log base of lib.rs:860 — comes from the
    host libm: every frame equals the oracle's colorize bit for bit (d.max() == 0)."""
    cfg = _small(S.Config.solar_sail(), 200, 220, 8_000_000)
    cfg.transparent = True
    angles = S.angle_iter(200.0, 245.0, 15.0)                    # 200, 215, 230 degrees
    r = S.ParallelRenderer.new(threads=256)
    frames = S.render_sequence(r, cfg, angles, 1, seed=77, shared_points=True)
    ocfg = cfg.to_pod()
    ocfg.iterations = 8_000_000 // 256
    pts = oracle.seed_points(77, 0, 256)
    for f, a in enumerate(angles):
        ocfg.angle = a
        ort = oracle.Runtime(200, 220)
        oracle.render_jobs_mt(ocfg, ort, pts)
        assert ort.max >= (1 << 20), "this test is about max beyond the ln table"
        d = np.abs(frames[f].astype(np.int32) - oracle.colorize(ocfg, ort).astype(np.int32))
        assert d.max() == 0, f"frame {f}: {(d > 0).sum()} values differ"
    r.shutdown()


def test_cross_gpu_wait_times_out_instead_of_hanging(S):
    """A rank whose peers never signal (a dead process) must cost a timeout, not a hung GPU or a frame computed from
    half-delivered data: the wait records sync_error, later protocol kernels of that Runtime do nothing,
    and the error can be read and cleared."""
    L = S._native.lib()
    cfg = _small(S.Config.poisson_saturne(), 64, 64, 100)
    rt = S.Runtime.new(cfg)
    err = C.c_uint32(7)
    S._native.check(L.sar_runtime_sync_error(rt._h, C.byref(err), 0))
    assert err.value == 0
    S._native.check(L.sar_set_option(b"sync_timeout_ms", 50))
    try:
        S._native.check(L.sar_frame_image_wait_async(rt._h, 2, 5, None))      # waits for IMAGE_DONE(5) from 2 ranks: never comes
        S._native.check(L.sar_stream_synchronize(rt._h, None))
        S._native.check(L.sar_runtime_sync_error(rt._h, C.byref(err), 0))
        assert err.value == 1 + 3, "IMAGE_DONE is event kind 3"
        # a later kernel of the protocol on this Runtime is a no-op: the accumulators stay as they are
        S.render(cfg, rt, initial_points=S.seed_points(1, 0, 8))
        before = rt.download()[0].copy()
        assert before.sum() == 8 * 100
        S._native.check(L.sar_frame_reset_async(rt._h, 1, 6, None))            # would zero the accumulators
        S._native.check(L.sar_stream_synchronize(rt._h, None))
        assert np.array_equal(rt.download()[0], before), "a protocol kernel ran on a Runtime whose wait had failed"
        S._native.check(L.sar_runtime_sync_error(rt._h, C.byref(err), 1))      # read and clear
        assert err.value == 4
        S._native.check(L.sar_runtime_sync_error(rt._h, C.byref(err), 0))
        assert err.value == 0
    finally:
        S._native.check(L.sar_set_option(b"sync_timeout_ms", 10_000))
    # with the error cleared the protocol works again (a 1-rank frame: reset waits for nothing new)
    S._native.check(L.sar_frame_reset_async(rt._h, 1, 1, None))
    S._native.check(L.sar_stream_synchronize(rt._h, None))
    assert rt.download()[0].sum() == 0


def test_tile_scatter_path_is_bit_exact(S, oracle):
    """Images that fit a shared-memory tile (W*H <= 25 600 pixels) with at least one full block of lanes take the
    per-block privatised histogram path (the north star's scatter).  Same bar as the L2 path: count,
    zbuf, steps and image equal the oracle's — incl. NaN trajectories, out-of-view points, a 1x1 image, Depth, a
    progressive mix of tile and L2 launches on one Runtime, and the cubic attractor."""
    L = S._native.lib()
    cases = [
        (S.Config.poisson_saturne(), 64, 64, 400, 2000, {}),
        (S.Config.solar_sail(), 120, 100, 300, 3000, {"angle": 220.0 * math.pi / 180.0}),      # NaN sink
        (S.Config.poisson_saturne(), 1, 1, 50, 1000, {}),
        (S.Config.poisson_saturne(), 160, 160, 200, 1500, {"scale": 5.0}),                     # the largest tile; mostly out of view
        (S.Config.solar_sail(), 97, 131, 250, 1000, {"depth": True}),
    ]
    for base, w, h, iters, jobs, opt in cases:
        cfg = _small(base, w, h, iters)
        cfg.angle = opt.get("angle", 0.3)
        if "scale" in opt:
            cfg.view.scale = opt["scale"]
        if opt.get("depth"):
            cfg.render = S.RenderKind.Depth
        pts = S.seed_points(40 + w, 0, jobs)
        rt = S.Runtime.new(cfg)
        S.render(cfg, rt, initial_points=pts)
        ort, st = _oracle_state(oracle, cfg, pts)
        _assert_state_equal(rt.download(), ort)
        img, f32 = S.colorize(cfg, rt, want_f32=True)
        oimg, of64 = oracle.colorize(cfg.to_pod(), ort, want_f64=True)
        _assert_image_close(img, f32, oimg, of64)
        # the knob: the L2 path on the same job list gives the same bytes
        S._native.check(L.sar_set_option(b"tile_scatter", 0))
        try:
            rt2 = S.Runtime.new(cfg)
            S.render(cfg, rt2, initial_points=pts)
            for a, b in zip(rt.download()[:3], rt2.download()[:3]):
                assert np.array_equal(a.view(np.uint8), b.view(np.uint8))
        finally:
            S._native.check(L.sar_set_option(b"tile_scatter", 1))
    # progressive: tile launch, then a small launch (L2 path: fewer than 896 jobs), then a tile launch again; max tracked across them
    cfg = _small(S.Config.poisson_saturne(), 100, 80, 300)
    a, b, c = S.seed_points(1, 0, 1200), S.seed_points(2, 0, 40), S.seed_points(3, 0, 900)
    rt = S.Runtime.new(cfg)
    for pts in (a, b, c):
        S.render(cfg, rt, initial_points=pts)
    ort, _ = _oracle_state(oracle, cfg, np.concatenate([a, b, c]))
    _assert_state_equal(rt.download(), ort)
    assert np.array_equal(S.colorize(cfg, rt), oracle.colorize(cfg.to_pod(), ort))
    # just above the tile limit: 161 x 160 pixels go the L2 way (and still match)
    cfg = _small(S.Config.poisson_saturne(), 161, 160, 100)
    pts = S.seed_points(9, 0, 1000)
    rt = S.Runtime.new(cfg)
    S.render(cfg, rt, initial_points=pts)
    ort, _ = _oracle_state(oracle, cfg, pts)
    _assert_state_equal(rt.download(), ort)
    # the cubic attractor kind has a tile instantiation too
    base = S.Config.poisson_saturne()
    cfg = _small(S.Config.poisson_saturne(), 90, 90, 200)
    cfg.attractor = S.attractors.PolynomialSprott3Degree(base.attractor.x, base.attractor.y, base.attractor.z,
                                                         [-0.05] + [0.0] * 9, [0.0] * 10, [0.0] * 9 + [-0.05])
    pts = S.seed_points(4, 0, 1000)
    rt = S.Runtime.new(cfg)
    S.render(cfg, rt, initial_points=pts)
    ort, _ = _oracle_state(oracle, cfg, pts)
    _assert_state_equal(rt.download(), ort)


def test_randomised_tile_configs_bit_exact(S, oracle):
    """The randomised sweep of test_randomised_configs_bit_exact on images that fit a shared-memory tile with enough
    jobs to take the tile path (>= 896): shapes, scale, camera, angle, rotation, transforms, palettes, perturbed
    coefficients (some diverge), Depth — state and RGBA16 image must equal the oracle's."""
    rng = np.random.default_rng(20261018)
    for case in range(10):
        base = S.Config.poisson_saturne() if case % 2 == 0 else S.Config.solar_sail()
        w = int(rng.integers(1, 161))
        h = int(rng.integers(1, max(2, min(161, 25_600 // w + 1))))
        assert w * h <= 25_600
        cfg = _small(base, w, h, int(rng.integers(1, 400)))
        cfg.angle = float(rng.uniform(-7.0, 7.0))
        cfg.view.scale = float(rng.choice([0.3, 1.0, 1.7, 4.0]))
        cfg.view.center_camera.x += float(rng.normal(0, 0.05))
        cfg.view.rotation.rotation = float(rng.uniform(0, 6.3))
        cfg.transparent = bool(rng.integers(0, 2))
        if case % 3 == 0:
            cfg.color_transform = S.color_transforms.AdjustedVelocity(offset=float(rng.uniform(-0.2, 0.5)), factor=float(rng.uniform(0.5, 3.0)))
        if case % 4 == 1:
            cfg.color_transform = S.color_transforms.ScreenBlend([float(v) for v in rng.uniform(-1, 1, 4)], offset=0.3, factor=0.8)
        if case >= 5:
            for lst in (cfg.attractor.x, cfg.attractor.y, cfg.attractor.z):
                lst[int(rng.integers(0, 10))] += float(rng.normal(0, 0.02))
        if case == 9:
            cfg.render = S.RenderKind.Depth
        pts = S.seed_points(2000 + case, 0, int(rng.integers(896, 2500)))
        rt = S.Runtime.new(cfg)
        S.render(cfg, rt, initial_points=pts)
        ort, _ = _oracle_state(oracle, cfg, pts)
        _assert_state_equal(rt.download(), ort)
        img, f32 = S.colorize(cfg, rt, want_f32=True)
        oimg, of64 = oracle.colorize(cfg.to_pod(), ort, want_f64=True)
        _assert_image_close(img, f32, oimg, of64)
        rt.close()


@pytest.mark.parametrize("name", ["poisson_saturne", "solar_sail", "solar_sail_220"])
def test_gpu_colorize_reproduces_the_reference_s_published_pixels(S, name):
    """Not GPU-vs-oracle: GPU vs the REFERENCE's own output.  tests/golden/media_inverse.npz holds, for 50 000 pixels of
    each media/*.png, the (count, steps) and the image's Runtime.max recovered by inverting colorize (lib.rs:853-874) such
    that the formula returns the PNG's 16-bit channels exactly (tests/test_reference_images.py).  Uploaded as a 1-row
    Runtime, sar_colorize must give those bytes back — solar-sail's max (k x 6 944 444) is far beyond the ln table, so
    this is also the host-ln(max+1) path.  Pixels the pre-clamp revision coloured from a negative position are skipped."""
    import os

    import test_reference_images as T

    inv = np.load(os.path.join(T.ROOT, "tests", "golden", "media_inverse.npz"))
    count, steps, zbuf, mx = T.strip_runtime_arrays(inv, name)
    preset, off, _ = T.IMAGES[name]
    cfg = getattr(S.Config, preset)()
    cfg.width, cfg.height, cfg.transparent = count.shape[1], 1, False
    cfg.colors.brighness.offset = off
    rt = S.Runtime.new(cfg)
    rt.upload(count, steps, zbuf)
    c2, s2, z2, gmax = rt.download()
    assert gmax == mx and np.array_equal(c2, count) and np.array_equal(s2, steps)
    img = S.colorize(cfg, rt)[0, 1:, :]
    head = inv[name + "_v"] >= 0
    assert head.sum() > 30_000
    assert np.array_equal(img[head, :3], inv[name + "_rgb"][head]), "GPU colourise differs from the reference's published pixels"
    assert (img[:, 3] == 65535).all()
