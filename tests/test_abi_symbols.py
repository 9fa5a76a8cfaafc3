"""CPU checks of the drop-in boundary: the library loads, exports every symbol include/sar.h
declares, the struct layout matches, and — with no GPU — compute entry points fail loudly
instead of falling back to anything."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "sar.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sar_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def S():
    import strange_attractor_renderer_b200 as S

    if S.build.needs_build():
        S.build.build()
    return S


def test_every_declared_symbol_is_exported_and_bound(S):
    declared = _declared_symbols()
    assert len(declared) >= 35
    L = C.CDLL(S._native.LIB_PATH)
    missing = [s for s in declared if not hasattr(L, s)]
    assert not missing, f"declared in include/sar.h but not exported: {missing}"
    assert sorted(S._native.SYMBOLS) == declared, set(declared) ^ set(S._native.SYMBOLS)
    assert S._native.lib().sar_abi_version() == 2


def test_library_does_not_link_the_oracle(S):
    """The product must not route through oracle/ (or any CPU restatement)."""
    import subprocess

    out = subprocess.run(["nm", "-D", "--defined-only", S._native.LIB_PATH], capture_output=True, text=True).stdout
    assert "orc_" not in out
    needed = subprocess.run(["ldd", S._native.LIB_PATH], capture_output=True, text=True).stdout
    assert "liboracle" not in needed


def test_config_struct_layout_matches_header(S, oracle):
    assert C.sizeof(S.SarConfig) == C.sizeof(oracle.SarConfig) == 8 + 6 * 4 + 8 + 240 + 24 + 24 + 8 * 4 + 8 + 16 * 24 + 16 + 240 + 32
    # presets held by the library == presets restated by the oracle, byte for byte
    assert bytes(S.Config.poisson_saturne().to_pod()) == bytes(oracle.poisson_saturne())
    assert bytes(S.Config.solar_sail().to_pod()) == bytes(oracle.solar_sail())
    c = S.Config.poisson_saturne()
    assert (c.iterations, c.width, c.height, c.transparent, c.angle, c.silent) == (10_000_000, 1920, 1080, True, 0.0, True)
    assert c.colors.brighness.offset == -0.15 and c.colors.brighness.factor == 5.0 / 3.0
    assert c.view.center_camera.z == -0.366 + 0.12
    s = S.Config.solar_sail()
    assert isinstance(s.color_transform, S.color_transforms.AdjustedVelocity)
    assert (s.color_transform.factor, s.color_transform.offset, s.view.scale) == (-0.2, 0.8, 1.7)


def test_seed_points_generator(S, oracle):
    a = S.seed_points(1234, 0, 1000)
    assert np.array_equal(a, oracle.seed_points(1234, 0, 1000))
    assert np.array_equal(a[100:110], S.seed_points(1234, 100, 10))       # random access
    assert a.min() >= 0.0 and a.max() < 0.1                               # `* 0.1`, lib.rs:748
    assert not np.array_equal(a, S.seed_points(1235, 0, 1000))
    # definition check in plain Python (SplitMix64)
    M = 2**64 - 1

    def sm(seed, n):
        z = (seed + (n + 1) * 0x9E3779B97F4A7C15) & M
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M
        return z ^ (z >> 31)

    for k in (0, 7, 999):
        for c in range(3):
            assert a[k, c] == ((sm(1234, 3 * k + c) >> 11) * 2.0**-53) * 0.1


def test_invalid_configs_are_rejected_on_the_host(S):
    cfg = S.Config.poisson_saturne()
    with pytest.raises(S.SarError):
        S.Palette([])                                   # Palette::new panics on empty, lib.rs:415
    cfg.color_transform = lambda d, s, v: 0.5           # arbitrary closures cannot cross to the GPU
    with pytest.raises(S.SarError) as e:
        cfg.to_pod()
    assert e.value.code == S._native.SAR_ERR_UNSUPPORTED
    cfg = S.Config.poisson_saturne()
    cfg.attractor.x = cfg.attractor.x[:9]
    with pytest.raises(S.SarError):
        cfg.to_pod()


def test_no_gpu_means_loud_failure_not_fallback(S):
    n = C.c_int(-1)
    rc = S._native.lib().sar_device_count(C.byref(n))
    if rc == 0 and n.value > 0:
        pytest.skip("a GPU is visible here")
    with pytest.raises(S.SarError) as e:
        S.Runtime.new(S.Config.poisson_saturne())
    assert e.value.code == S._native.SAR_ERR_CUDA
    with pytest.raises(S.SarError):
        S.ParallelRenderer.new()
    assert S._native.lib().sar_last_error()
