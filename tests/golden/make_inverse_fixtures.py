"""Invert the reference's colourise on its own published images -> exact per-pixel fixtures.

Run in the build container only (it reads /root/reference/media):

    python tests/golden/make_inverse_fixtures.py

Idea.  `colorize` (lib.rs:853-874) maps (count, steps, max) of a pixel to three u16 channels
    c = floor((sqrt(lerp(palette, steps)) * ln(count+1)/ln(max+1) + offset) * factor * 65535)
with a piecewise-linear palette whose entries are 0.5 or 1.0.  Three integers, two unknowns
(count is an integer, steps a real): for every pixel whose channels are neither 0 nor 65535
the system is over-determined, so an exact preimage (count, steps) exists only if the formula,
its constants, the truncating cast, the +1 in both logarithms and `max` are what the reference
used.  `max` itself is recovered from the image: only one integer makes the pixels solvable
(poisson-saturne: 95 125), and for the solar-sail images it equals k * (1e9 / 12 / 12) — the
NaN sink of pixel (0,0) fed by k diverging jobs of `render_parallel`'s decomposition
(lib.rs:1056-1058) on the author's 12 threads x 12 jobs.

Findings recorded in the fixture (`*_stats`): every fully informative pixel of the three images
(679 266 + 772 323 + 681 661) has an exact preimage — under ONE difference from lib.rs @ HEAD:
the published images were made before `Palette::interpolate` clamped negative positions to 0
(lib.rs:443-444): 10 / 150 081 / 140 947 pixels need a negative position in the first palette
segment (extrapolation).  Pixels with position >= 0 (all but 10 of poisson-saturne) are what the
tests require the oracle (and the GPU) to reproduce bit for bit.

Output: tests/golden/media_inverse.npz, per image a random sample of SAMPLE solved pixels
  <name>_idx  u32  flat pixel index y*W+x          <name>_rgb  u16 [K,3]  the PNG's channels
  <name>_n    u32  recovered count                 <name>_v    f32  recovered palette position (steps)
  <name>_max  i64  recovered Runtime.max           <name>_stats i64 [full, solved, negative, sum_n, W, H]
The recovered count field also feeds a pixel-level chi-square test against the oracle's render.
"""
import math
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from make_media_fixtures import MEDIA, decode_png16  # noqa: E402

OUT = os.path.join(HERE, "media_inverse.npz")
SAMPLE = 50_000
PER_JOB = 1_000_000_000 // 12 // 12        # lib.rs:1058 on 12 threads x 12 jobs (main.rs:305)
# name -> (file, brightness offset of the README command, max or None = search)
IMAGES = {
    "poisson_saturne": ("poisson-saturne.png", -0.25, None),
    "solar_sail": ("solar-sail.png", -0.1, 58 * PER_JOB),
    "solar_sail_220": ("solar-sail-220deg.png", -0.15, 54 * PER_JOB),
}
PAL = np.array([(1, 1, .5), (.5, 1, .5), (1, .5, .5), (.5, 1, 1), (.5, .5, 1), (1, .5, 1), (1, .5, 1)], dtype=np.float64)
BF = 5.0 / 3.0                              # lib.rs:397-404


def estimate_f(P, off):
    """Least-squares log factor per pixel (palette position free, negative allowed in segment 0)."""
    x2 = ((P + 0.5) / (65535.0 * BF) - off) ** 2
    best = np.full(len(P), np.inf)
    f = np.zeros(len(P))
    for s in range(6):
        A = np.stack([PAL[s], PAL[s + 1]], axis=1) if s < 5 else PAL[s][:, None]
        sol, *_ = np.linalg.lstsq(A, x2.T, rcond=None)
        res = np.abs(A @ sol - x2.T).max(axis=0)
        okk = (sol >= -1e-6).all(axis=0) if s else (sol[0] > 0)
        res = np.where(okk, res, np.inf)
        upd = res < best
        best[upd] = res[upd]
        f[upd] = np.sqrt(np.maximum(sol.sum(axis=0), 0))[upd]
    return f


def intervals(P, n, L, off):
    """For counts n: per palette segment the interval of t that reproduces all three channels."""
    glo = P / (BF * 65535.0) - off
    ghi = (P + 1.0) / (BF * 65535.0) - off
    f = np.log(n + 1.0) / L
    qlo = (glo / f[:, None]) ** 2
    qhi = (ghi / f[:, None]) ** 2
    for s in range(6):
        A, B = PAL[s], PAL[s + 1]
        tlo = np.full(len(P), -2.0 if s == 0 else 0.0)
        thi = np.ones(len(P))
        feas = n >= 1
        for c in range(3):
            if A[c] == B[c]:
                feas = feas & (qlo[:, c] <= A[c] * (1 + 1e-13)) & (A[c] * (1 - 1e-13) < qhi[:, c])
            else:
                a = (qlo[:, c] - A[c]) / (B[c] - A[c])
                b = (qhi[:, c] - A[c]) / (B[c] - A[c])
                tlo = np.maximum(tlo, np.minimum(a, b))
                thi = np.minimum(thi, np.maximum(a, b))
        yield s, feas & (thi - tlo > 1e-9), tlo, thi


def solve(P, f_est, L, off, K=3):
    N = len(P)
    n_est = np.rint(np.exp(f_est * L) - 1.0).astype(np.int64)
    out_n = np.zeros(N, np.int64)
    out_v = np.zeros(N)
    ok = np.zeros(N, bool)
    for d in sorted(range(-K, K + 1), key=abs):
        n = np.maximum(n_est + d, 0)
        for s, feas, tlo, thi in intervals(P, n.astype(np.float64), L, off):
            new = feas & ~ok
            out_n[new] = n[new]
            out_v[new] = (s + 0.5 * (tlo + thi)[new]) / 6.0
            ok |= new
    return out_n, out_v, ok


def brute(P1, L, off, n_top):
    """One pixel, every count up to n_top (palette vertices defeat the estimator)."""
    nn = np.arange(1, n_top, dtype=np.float64)
    PP = np.repeat(P1[None, :], len(nn), axis=0)
    for s, feas, tlo, thi in intervals(PP, nn, L, off):
        k = np.nonzero(feas)[0]
        if len(k):
            return int(nn[k[0]]), (s + 0.5 * (tlo[k[0]] + thi[k[0]])) / 6.0
    return None


def forward(n, v, mx, off, clamp_low):
    """Plain-Python colourise of one pixel (lib.rs:442-472, 853-868); clamp_low=False = the pre-clamp revision."""
    if v < 0.0 and clamp_low:
        v = 0.0
    elif v >= 1.0:
        v = 0.999999
    val = v * 6.0
    s = int(math.floor(val)) if val >= 0.0 else 0       # negative `as usize` saturates to 0
    t = math.fmod(val, 1.0)
    t1 = 1.0 - t
    factor = math.log(float(n + 1)) / math.log(float(mx + 1))
    out = []
    for c in range(3):
        p = math.sqrt(PAL[s + 1][c] * t + PAL[s][c] * t1)
        x = (p * factor + off) * BF * 65535.0
        out.append(0 if x <= 0 else (65535 if x >= 65535 else int(x)))
    return out


def find_max(P, f_est, off):
    """The integer max that makes the pixels solvable (sharp: +-1 loses 8 % of them)."""
    lo = np.sort(f_est)[:4000]
    lev = lo[np.r_[True, np.diff(lo) > 3e-5]][:3]
    cands = []
    for n0 in range(1, 400):
        L = math.log(n0 + 1) / lev[0]
        if abs(math.log(n0 + 2) / L - lev[1]) < 2e-5 and abs(math.log(n0 + 3) / L - lev[2]) < 3e-5:
            cands.append(int(round(math.exp(L))))
    rng = np.random.default_rng(0)
    sub = rng.choice(len(P), min(20000, len(P)), replace=False)
    best = (0.0, None)
    for c in cands:
        for m1 in range(c - 60, c + 61):
            ok = solve(P[sub], f_est[sub], math.log(m1), off, K=2)[2].mean()
            if ok > best[0]:
                best = (ok, m1)
    return best[1] - 1, best[0]


def main():
    from oracle import oracle as O

    store = {}
    for name, (fn, off, mx) in IMAGES.items():
        img = decode_png16(os.path.join(MEDIA, fn))
        H, W, _ = img.shape
        full = ((img > 0) & (img < 65535)).all(axis=2)
        idx = np.flatnonzero(full.ravel())
        P = img.reshape(-1, 3)[idx].astype(np.float64)
        f_est = estimate_f(P, off)
        if mx is None:
            mx, frac = find_max(P, f_est, off)
            print(name, "max found", mx, "solvable fraction of the probe", frac)
        L = math.log(mx + 1)
        n, v, ok = solve(P, f_est, L, off)
        for i in np.flatnonzero(~ok):
            r = brute(P[i], L, off, 1_000_000)
            if r:
                n[i], v[i], ok[i] = r[0], r[1], True
        v32 = v.astype(np.float32)
        neg = ok & (v32 < 0)
        print(name, (H, W), "full", len(P), "solved", int(ok.sum()), "negative position", int(neg.sum()),
              "sum n", int(n[ok].sum()), "max", mx)
        # verification 1: the oracle (HEAD semantics) reproduces EVERY solved pixel with position >= 0
        cfg = O.poisson_saturne() if name == "poisson_saturne" else O.solar_sail()
        cfg.width, cfg.height, cfg.transparent, cfg.bright_offset = W, H, 0, off
        cnt = np.zeros(H * W, np.uint32)
        st = np.zeros(H * W)
        pos = ok & ~neg
        cnt[idx[pos]] = n[pos]
        st[idx[pos]] = v32[pos].astype(np.float64)
        assert not full.ravel()[0]
        cnt[0] = mx                                  # the pixel that holds Runtime.max
        rt = O.Runtime(W, H)
        rt.load(cnt.reshape(H, W), st.reshape(H, W), np.zeros((H, W), np.float32))
        assert rt.max == mx
        got = O.colorize(cfg, rt).reshape(-1, 4)[idx[pos], :3]
        exact = (got == img.reshape(-1, 3)[idx[pos]]).all(axis=1)
        print("   oracle reproduces", int(exact.sum()), "of", int(pos.sum()), "pixels with position >= 0")
        good = np.zeros(len(P), bool)
        good[np.flatnonzero(pos)[exact]] = True
        # the few misses sit on the last palette vertex (position 5/6, where f32 rounding crosses into the next
        # segment's neighbour): any position inside the constant last segment gives the same colour
        for i in np.flatnonzero(pos)[~exact]:
            target = list(img.reshape(-1, 3)[idx[i]])
            for cand in [np.nextafter(v32[i], np.float32(k), dtype=np.float32) for k in (0, 1)] + [np.float32(0.9)]:
                if forward(int(n[i]), float(cand), mx, off, True) == target:
                    v32[i] = cand
                    good[i] = True
                    break
        print("   after moving vertex positions:", int((good & pos).sum()), "of", int(pos.sum()))
        # verification 2: the pre-clamp formula reproduces the negative-position pixels
        for i in np.flatnonzero(neg):
            if forward(int(n[i]), float(v32[i]), mx, off, clamp_low=False) == list(img.reshape(-1, 3)[idx[i]]):
                good[i] = True
        print("   pre-clamp formula reproduces", int((good & neg).sum()), "of", int(neg.sum()), "negative-position pixels")
        rng = np.random.default_rng(2)
        pick = np.sort(rng.choice(np.flatnonzero(good), min(SAMPLE, int(good.sum())), replace=False))
        store[name + "_idx"] = idx[pick].astype(np.uint32)
        store[name + "_rgb"] = img.reshape(-1, 3)[idx[pick]].astype(np.uint16)
        store[name + "_n"] = n[pick].astype(np.uint32)
        store[name + "_v"] = v32[pick]
        store[name + "_max"] = np.int64(mx)
        store[name + "_stats"] = np.array([len(P), int(good.sum()), int((good & neg).sum()), int(n[good].sum()), W, H], np.int64)
    np.savez_compressed(OUT, **store)
    print("wrote", OUT, os.path.getsize(OUT))


if __name__ == "__main__":
    main()
