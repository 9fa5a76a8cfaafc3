"""Attractor / ColorTransform kinds beyond the two the reference ships (README.md:8,
lib.rs:71-77, 241-249): PolynomialSprott3Degree and ScreenBlend, defined in include/sar.h.  There is no
reference code for them, so they are pinned the way the rest is: C oracle == independent Python
restatement (CPU), CUDA path == C oracle bit for bit (GPU)."""
import numpy as np
import pytest

import pyref

# a cubic perturbation of poisson-saturne that stays bounded: small cubic damping terms
X3 = [-0.05, 0.0, 0.0, 0.02, 0.0, 0.0, 0.0, 0.0, 0.01, 0.0]
Y3 = [0.0, -0.03, 0.0, 0.0, 0.0, 0.0, -0.05, 0.0, 0.0, 0.0]
Z3 = [0.0, 0.0, -0.02, 0.0, 0.01, 0.0, 0.0, 0.0, 0.0, -0.05]


def _cubic(cfg):
    cfg.attractor_kind = 1
    for k, lst in enumerate((X3, Y3, Z3)):
        for i, v in enumerate(lst):
            cfg.coef3[k][i] = v
    return cfg


def _blend(cfg):
    cfg.ct_kind, cfg.ct_offset, cfg.ct_factor = 2, 0.35, 0.9
    cfg.ct_weights[:] = [0.4, -0.3, 0.25, 1.5]
    return cfg


def test_oracle_and_python_restatement_agree_on_the_new_kinds(oracle):
    rng = np.random.default_rng(11)
    cfg = _blend(_cubic(oracle.poisson_saturne()))
    coef = [list(cfg.coef[k]) for k in range(3)]
    coef3 = [list(cfg.coef3[k]) for k in range(3)]
    for _ in range(300):
        p = rng.uniform(-0.6, 0.6, 3)
        assert oracle.next_point(cfg, p).tolist() == pyref.next_point(coef, [float(v) for v in p], coef3)
        d, s = rng.normal(0, 0.3, 3), rng.normal(0, 0.4, 3)
        assert oracle.color_transform(cfg, d, s) == pyref.color_transform(cfg, list(d), list(s))
    # with all cubic coefficients zero the cubic kind walks the quadratic trajectory (x + 0.0 == x)
    z = oracle.poisson_saturne()
    z.attractor_kind = 1
    p = np.array([0.01, 0.02, 0.03])
    q = p.copy()
    for _ in range(500):
        p, q = oracle.next_point(z, p), oracle.next_point(oracle.poisson_saturne(), q)
    assert np.array_equal(p, q)
    cfg.width, cfg.height, cfg.iterations = 40, 36, 3000
    crt, prt = oracle.Runtime(40, 36), pyref.Runtime(40, 36)
    for pt in oracle.seed_points(3, 0, 6):
        oracle.render(cfg, crt, pt)
        pyref.render(cfg, prt, [float(v) for v in pt])
    assert crt.count.sum() > 0, "the cubic test attractor must stay in view"
    assert crt.count.ravel().tolist() == prt.count and crt.zbuf.ravel().tolist() == prt.zbuf
    assert crt.steps.ravel().tolist() == prt.steps


@pytest.mark.gpu
def test_gpu_matches_oracle_on_the_new_kinds(oracle):
    import strange_attractor_renderer_b200 as S

    base = S.Config.poisson_saturne()
    cubic = S.attractors.PolynomialSprott3Degree(base.attractor.x, base.attractor.y, base.attractor.z, X3, Y3, Z3)
    blend = S.color_transforms.ScreenBlend([0.4, -0.3, 0.25, 1.5], offset=0.35, factor=0.9)
    for att, ct in ((cubic, base.color_transform), (base.attractor, blend), (cubic, blend)):
        cfg = S.Config.poisson_saturne()
        cfg.attractor, cfg.color_transform = att, ct
        cfg.width, cfg.height, cfg.iterations, cfg.angle = 257, 190, 6_000, 0.4
        pts = S.seed_points(21, 0, 300)
        rt = S.Runtime.new(cfg)
        S.render(cfg, rt, initial_points=pts)
        ort = oracle.Runtime(257, 190)
        st = oracle.OrcStats()
        oracle.render_jobs(cfg.to_pod(), ort, pts, st)
        assert st.recorded > 0.5 * 300 * 6_000
        count, steps, zbuf, mx = rt.download()
        assert np.array_equal(count, ort.count) and mx == ort.max
        assert np.array_equal(zbuf.view(np.uint32), ort.zbuf.view(np.uint32))
        assert np.array_equal(steps.view(np.uint64), ort.steps.view(np.uint64))
        assert np.array_equal(S.colorize(cfg, rt), oracle.colorize(cfg.to_pod(), ort))
        # round trip of the Config mirror
        back = S.Config._from_pod(cfg.to_pod())
        assert type(back.attractor) is type(att) and type(back.color_transform) is type(ct)
    # the auto-framing pass and the warm-up kernel follow the attractor kind too
    cfg = S.Config.poisson_saturne()
    cfg.attractor = cubic
    af = S.autoframe(cfg, n_jobs=512, iterations=2_000, seed=4)
    box, bad = oracle.screen_bbox_jobs(cfg.to_pod(), oracle.seed_points(4, 0, 512), 2_000)
    assert np.array_equal(np.array(af.box), box) and af.diverged == bad
    cfg.width, cfg.height, cfg.iterations = 120, 100, 400_000
    r = S.ParallelRenderer.new(threads=128)
    frames = S.render_sequence(r, cfg, [0.0, 0.5], 1, seed=6, shared_points=True)      # warm_kernel path
    want = S.render_parallel(r, cfg, 1, initial_points=S.seed_points(6, 0, 128))
    assert np.array_equal(frames[0], want)
    r.shutdown()
