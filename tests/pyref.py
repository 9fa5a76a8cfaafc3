"""A second, independent restatement of the reference's render path in plain Python floats
(IEEE f64, no fusing) — small cases only.  It exists to catch transcription errors in the C
oracle: two restatements written separately from src/lib.rs must agree bit for bit.
Citations are into the reference's src/lib.rs @ e571d19."""
import math
import struct


def f32(x: float) -> float:
    """`x as f32` (round to nearest even, overflow to inf), back as a Python float."""
    try:
        return struct.unpack("f", struct.pack("f", x))[0]
    except OverflowError:
        return math.copysign(math.inf, x)


def as_uint(v: float, top: int) -> int:
    """Rust float -> unsigned `as` cast: saturating, NaN -> 0, truncating."""
    if v != v or v <= 0.0:
        return 0
    if v >= float(top):
        return top
    return int(v)


def next_point(coef, p, coef3=None):                         # lib.rs:585-620; coef3: the cubic kind of include/sar.h
    x, y, z = p
    m = [1.0, x, x * x, x * y, x * z, y, y * y, y * z, z, z * z]
    if coef3 is not None:
        m3 = [m[2] * x, m[2] * y, m[2] * z, m[3] * y, m[3] * z, m[4] * z, m[6] * y, m[6] * z, m[7] * z, m[9] * z]
    out = []
    for k in range(3):
        s = 0.0
        for i in range(10):
            s += m[i] * coef[k][i]
        if coef3 is not None:
            for i in range(10):
                s += m3[i] * coef3[k][i]
        out.append(s)
    return out


def rotation_matrix(axis, rot):                              # lib.rs:179-195, release build
    x, y, z = axis
    c = math.cos(rot)
    c1 = 1.0 - c
    s = math.sin(rot)
    return [[c + x * x * c1, x * y * c1 - z * s, x * z * c1 + y * s],
            [y * x * c1 + z * s, c + y * y * c1, y * z * c1 - x * s],
            [z * x * c1 - y * s, z * y * c1 + x * s, c + z * z * c1]]


def mul_right(m, v):                                         # lib.rs:208-215
    return [m[r][0] * v[0] + m[r][1] * v[1] + m[r][2] * v[2] for r in range(3)]


def magnitude(v):                                            # lib.rs:129-131
    return math.sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2])


def color_transform(cfg, delta, p):                          # lib.rs:511-516, 520-558
    if cfg.ct_kind == 1:
        return (magnitude(delta) + cfg.ct_offset) * cfg.ct_factor
    if cfg.ct_kind == 2:                                     # ScreenBlend, include/sar.h
        w = list(cfg.ct_weights)
        t = p[0] * w[0]
        t = t + p[1] * w[1]
        t = t + p[2] * w[2]
        t = t + magnitude(delta) * w[3]
        return (t + cfg.ct_offset) * cfg.ct_factor
    COS = 0.7009092642998508981833083453238941729068756103515625
    SIN = 0.7132504491541815649924274111981503665447235107421875
    x2 = (p[0] + cfg.center_camera[0]) * COS + (p[2] + cfg.center_camera[1]) * SIN
    if (x2 < -0.0839 or 10.55 * x2 + p[1] < 0.46 - 1.0941 or 1.0426 * x2 + p[1] < 0.179 - 0.1576
            or 0.5139 * x2 - p[1] > -0.04 - 0.04092):
        part = 0.0
    else:
        part = 1.0
    color = (part + magnitude(delta)) / 2.0
    return (color - 0.1) / 0.9


def palette_interpolate(cfg, value):                         # lib.rs:442-472
    if value < 0.0:
        value = 0.0
    elif value >= 1.0:
        value = 0.999999
    n_col = cfg.palette_len
    lst = [list(cfg.palette_rgb[i]) for i in range(n_col)]
    lst.append(lst[-1])                                      # lib.rs:418
    value = value * float(n_col)
    n = as_uint(math.floor(value), 2**64 - 1) if value == value else 0
    t = math.fmod(value, 1.0) if value == value else value
    t1 = 1.0 - t
    return [math.sqrt(lst[n + 1][c] * t + lst[n][c] * t1) for c in range(3)]


class Runtime:                                               # lib.rs:631-699
    def __init__(self, w, h):
        self.w, self.h = w, h
        self.count = [0] * (w * h)
        self.steps = [0.0] * (w * h)
        self.zbuf = [-1.0] * (w * h)
        self.max = 0


def render(cfg, rt, init):                                   # lib.rs:747-838
    cur = list(init)
    coef = [list(cfg.coef[k]) for k in range(3)]
    coef3 = [list(cfg.coef3[k]) for k in range(3)] if cfg.attractor_kind == 1 else None
    for _ in range(1000):
        cur = next_point(coef, cur, coef3)
    R = rotation_matrix(list(cfg.axis), cfg.rotation)
    sin_v, cos_v = math.sin(cfg.angle), math.cos(cfg.angle)
    cc = list(cfg.center_camera)
    width, height = float(cfg.width), float(cfg.height)
    width_scaled = width * cfg.scale
    scale_adjusted_mid = 0.5 / cfg.scale
    prev = list(cur)
    for _ in range(cfg.iterations):
        cur = next_point(coef, cur, coef3)
        s = mul_right(R, cur)
        x2 = (s[0] + cc[0]) * cos_v + (s[2] + cc[1]) * sin_v
        z2 = (s[0] + cc[0]) * sin_v - (s[2] + cc[1]) * cos_v
        i = (scale_adjusted_mid - x2) * width_scaled
        j = height / 2.0 - (s[1] + cc[2]) * width_scaled
        if i >= width or j >= height or i < 0.0 or j < 0.0:
            prev = cur
            continue
        ii, jj = as_uint(i, 2**32 - 1), as_uint(j, 2**32 - 1)
        idx = jj * rt.w + ii
        rt.count[idx] = (rt.count[idx] + 1) & 0xFFFFFFFF
        if rt.count[idx] > rt.max:
            rt.max = rt.count[idx]
        zf = f32(z2)
        if zf > rt.zbuf[idx]:
            delta = [cur[0] - prev[0], cur[1] - prev[1], cur[2] - prev[2]]
            rt.steps[idx] = color_transform(cfg, delta, s)
            rt.zbuf[idx] = zf
        prev = cur


def _ln(v: int) -> float:
    return math.log(v) if v > 0 else -math.inf


def colorize(cfg, rt):                                       # lib.rs:841-904
    out = []
    if cfg.render_kind == 0:
        for p in range(rt.w * rt.h):
            r, g, b = palette_interpolate(cfg, rt.steps[p])
            num, den = _ln((rt.count[p] + 1) & 0xFFFFFFFF), _ln((rt.max + 1) & 0xFFFFFFFF)
            factor = num / den if den != 0.0 else (math.nan if num == 0.0 or num != num else math.copysign(math.inf, num))
            px = [as_uint((c * factor + cfg.bright_offset) * cfg.bright_factor * 65535.0, 65535) for c in (r, g, b)]
            px.append(as_uint(factor * 65535.0, 65535) if cfg.transparent else 65535)
            out.append(px)
    else:
        mx, mn = 0.0, f32(3.4028234663852886e38)
        for z in rt.zbuf:
            if z != -1.0:
                mx, mn = max(mx, z), min(mn, z)
        diff = f32(mx - mn)
        for z in rt.zbuf:
            zz = 0.0 if z == -1.0 else f32(f32(z - mn) / diff)
            g = as_uint(f32(zz * 65535.0), 65535)
            out.append([g, g, g, 65535])
    return out
