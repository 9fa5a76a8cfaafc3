#!/usr/bin/env python
"""bench.py — attractor iterations/sec on BASELINE.json's headline configuration.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One step = one whole frame of the hot path: Runtime::reset → render (all trajectories) →
max → colorize, for `poisson-saturne, 1e9 iterations, 2048x2048` (BASELINE.json configs[1]) per
GPU.  At N > 1 every rank renders its own 1e9-iteration share of ONE N*1e9-iteration frame
(weak scaling) and the ranks exchange row stripes over NVLink peer memory.

  value    : recorded iterations / s, device-resident (start points generated on the GPU, image
             left in HBM), timed with CUDA events on the launching stream, max over ranks.
  e2e      : the same frame through the reference-facing C-ABI call sar_render_parallel with HOST
             buffers: start points copied in from pinned host memory, RGBA16 image copied out.
  roofline : the iterate kernel alone, 12 algorithmic bytes per recorded iteration
             against the measured HBM copy bandwidth (MEASURED_PEAKS.json).
  cpu_baseline : the CPU oracle run as the reference's render_parallel (nproc threads x 12 jobs)
             on this host.  `--impl reference` prints that as its own line (the reference is Rust
             and there is no Rust toolchain, so kind = "port").
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "attractor iterations/sec"
UNIT = "iterations/s"
WORKLOAD = "poisson-saturne, 1e9 iterations, 2048x2048 (BASELINE.json configs[1])"
ITERATIONS = 1_000_000_000
WIDTH = HEIGHT = 2048
ALGO_BYTES_PER_ITER = 12  # count u32 read+write + zbuf f32 read
SEED = 1234


def measured_hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons while the timed region runs (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 6 for n, v in zip(names, r[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# --------------------------------------------------------------------------------------------
# CPU arm: the oracle as render_parallel on the host cores (test infrastructure used as baseline)
# --------------------------------------------------------------------------------------------
def cpu_render_parallel(iterations: int, threads: int, jobs_per_thread: int = 12):
    from oracle import oracle as O

    cfg = O.poisson_saturne()
    cfg.iterations, cfg.width, cfg.height, cfg.transparent = iterations, WIDTH, HEIGHT, 0
    pts = O.seed_points(SEED, 0, threads * jobs_per_thread)
    t0 = time.perf_counter()
    O.render_parallel(cfg, threads, jobs_per_thread, pts)
    dt = time.perf_counter() - t0
    recorded = (iterations // threads // jobs_per_thread) * threads * jobs_per_thread
    return recorded / dt, dt


def cpu_sample_size(threads: int, target_s: float = 15.0) -> int:
    est_rate = 0.7e7 * threads   # ~0.7e7 it/s/thread when memory-bound
    return int(min(ITERATIONS, max(2e7, est_rate * target_s)))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = os.cpu_count() or 8
    iters = cpu_sample_size(threads, 10.0)
    for _ in range(max(0, min(args.warmup, 1))):
        cpu_render_parallel(max(iters // 8, 10_000_000), threads)
    vals, times = [], []
    for _ in range(args.steps):
        v, dt = cpu_render_parallel(iters, threads)
        vals.append(v)
        times.append(dt)
    value = sum(vals) / len(vals)
    sample = f"{iters:.3g} of 1e9 iterations per step, render_parallel semantics: {threads} threads x 12 jobs, private 16 B/px buffers, serial merge + colorize"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        # same workload keys and values as the GPU arm's `config`; the bounded sample a step runs is `iterations_per_step`
        "config": {"workload": WORKLOAD, "width": WIDTH, "height": HEIGHT, "iterations_per_gpu": ITERATIONS,
                   "iterations_per_frame": ITERATIONS, "seed": SEED, "iterations_per_step": iters,
                   "l2": "n/a (host arm: private 16 B/px buffers per thread, 64 MiB each, far beyond the CPU caches)",
                   "note": "reference is Rust; no rustc/cargo in this image, so the CPU arm is the C restatement (oracle/) of lib.rs:747-1082"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))
    return 0


# --------------------------------------------------------------------------------------------
# Parity of the N-rank frame against the oracle (outside every timed region)
# --------------------------------------------------------------------------------------------
def parity_frame(world, rank, local, group, preset="solar", depth=False, per_gpu=20_000_000, lanes=2048, jpt=2,
                 width=450, height=501, seed=4321):
    """One small frame on `world` ranks (trajectory-sharded render, NVLink stripe merge, stripe
    colourise into rank 0) against the CPU oracle on the same job list: count / zbuf / steps /
    RGBA16 image, all bit-exact.  Returns the dict of checks on rank 0 (None elsewhere)."""
    import ctypes as C

    import numpy as np
    import torch
    import torch.distributed as dist

    import strange_attractor_renderer_b200 as S
    from strange_attractor_renderer_b200 import _native as N
    from strange_attractor_renderer_b200 import dist as D

    L = N.lib()
    cfg = S.Config.solar_sail() if preset == "solar" else S.Config.poisson_saturne()
    cfg.width, cfg.height, cfg.angle = width, height, 1.25          # height not divisible by the world size
    if depth:
        cfg.render = S.RenderKind.Depth
    frame = D.Frame(cfg, device=local, world=world, rank=rank, group=group, lanes=lanes, jobs_per_thread=jpt,
                    iterations_per_gpu=per_gpu, seed=seed)
    stream = torch.cuda.Stream(device=local)
    sp = C.c_void_p(stream.cuda_stream)
    for _ in range(2):                                              # twice: reset / re-merge must be clean
        frame.step_device(sp)
    torch.cuda.synchronize(local)
    frame.check_sync()
    if group is not None:
        D.barrier(group)
    h, w = frame.h, frame.w
    count = np.empty((h, w), np.uint32)
    steps = np.empty((h, w), np.float64)
    zbuf = np.empty((h, w), np.float32)
    N.check(L.sar_runtime_download(frame.rt, count.ctypes.data_as(N._u32p), steps.ctypes.data_as(N._f64p),
                                   zbuf.ctypes.data_as(N._f32p), None))
    r0, n = frame.row0, frame.rows
    part = (r0, n, count[r0:r0 + n].copy(), steps[r0:r0 + n].copy(), zbuf[r0:r0 + n].copy())
    parts = [part]
    if group is not None:
        parts = [None] * world
        dist.all_gather_object(parts, part)
    checks = None
    if rank == 0:
        from oracle import oracle as O          # the checker, never the thing measured

        img = np.empty((h, w, 4), np.uint16)
        N.check(L.sar_runtime_image_download(frame.rt, 0, 0, img.ctypes.data_as(N._u16p), None))
        for (a, m, c, s_, z) in parts:
            count[a:a + m], steps[a:a + m], zbuf[a:a + m] = c, s_, z
        ocfg = cfg.to_pod()
        ocfg.iterations = frame.iterations_per_job
        ort = O.Runtime(w, h)
        O.render_jobs_mt(ocfg, ort, O.seed_points(seed, 0, frame.lanes * jpt * world))
        oimg = O.colorize(ocfg, ort)
        d = np.abs(img.astype(np.int32) - oimg.astype(np.int32))
        checks = {
            "count": bool(np.array_equal(count, ort.count)),
            "zbuf": bool(np.array_equal(zbuf.view(np.uint32), ort.zbuf.view(np.uint32))),
            "steps": bool(np.array_equal(steps.view(np.uint64), ort.steps.view(np.uint64))),
            # the device-resident frame cannot read max back; when max + 1 is beyond the host-libm ln table
            # (solar-sail's NaN sink) the log base comes from the device log: <= 1 LSB
            "image": bool(d.max() == 0) if ort.max + 1 < (1 << 20) else bool(d.max() <= 1 and (d > 0).mean() < 1e-4),
            "image_exact": bool(d.max() == 0),
            "vs": "oracle",
            "frame": f"{preset}{' depth' if depth else ''} {w}x{h}, {world} rank(s) x {frame.lanes * jpt} jobs x {frame.iterations_per_job} iterations",
            "recorded": int(ort.count.sum(dtype=np.uint64)), "max": int(ort.max),
        }
    if rank == 0:
        N.check(L.sar_stream_synchronize(frame.rt, sp))
    if group is not None:
        D.barrier(group)
    frame.close()
    return checks


# --------------------------------------------------------------------------------------------
# The GPU's frame against the REFERENCE's own published output (outside every timed region)
# --------------------------------------------------------------------------------------------
def reference_image_check(count, steps, mx):
    """`count` / `steps` / `max` of a 1e9-iteration poisson-saturne frame at 1920x1080 (the README's command, README.md:73)
    against what tests/golden/make_inverse_fixtures.py recovered from media/poisson-saturne.png by inverting colorize
    (integer counts and palette positions of 50 000 pixels that reproduce the PNG's 16-bit channels exactly, max 95 125).
    The reference's seeds are unknowable, so equality is statistical: chi-square per pixel ~ 1 when both count fields are
    Poisson draws of the same density (oracle vs oracle with another seed: 0.992); one pixel of shift gives > 60."""
    import numpy as np

    inv = np.load(os.path.join(ROOT, "tests", "golden", "media_inverse.npz"))
    idx = inv["poisson_saturne_idx"].astype(np.int64)
    n = inv["poisson_saturne_n"].astype(np.float64)
    v = inv["poisson_saturne_v"].astype(np.float64)
    rmax = int(inv["poisson_saturne_max"])
    c = np.asarray(count).ravel().astype(np.float64)
    st = np.asarray(steps).ravel()

    def chi2(at):
        y = c[np.clip(at, 0, c.size - 1)]
        return float((((n - y) ** 2) / np.maximum(n + y, 1.0)).mean())

    chi = chi2(idx)
    off = min(chi2(idx + d) for d in (1, -1, 1920, -1920))
    o = st[idx]
    keep = (v < 5.0 / 6.0 - 1e-3) & (o < 5.0 / 6.0 - 1e-3) & (c[idx] > 0)      # the last palette segment is constant
    dv = float(np.median(np.abs(v[keep] - o[keep]))) if keep.any() else None
    mass = float(n.sum() / max(c[idx].sum(), 1.0))
    ok = bool(0.85 < chi < 1.25 and off > 15.0 and abs(mass - 1.0) < 2e-3 and abs(mx - rmax) < 5.0 * rmax ** 0.5
              and dv is not None and dv < 1e-4)
    return {"vs": "media/poisson-saturne.png (reference's published image; count field recovered by inverting colorize)",
            "pixels": int(idx.size), "chi2_per_pixel": chi, "chi2_one_pixel_off": off, "mass_ratio": mass,
            "max": int(mx), "max_reference": rmax, "palette_position_median_abs_diff": dv, "ok": ok}


# --------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------
def run_ours(args):
    import torch

    import strange_attractor_renderer_b200 as S
    from strange_attractor_renderer_b200 import _native as N
    from strange_attractor_renderer_b200 import dist as D

    L = N.lib()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus N > 1 must be launched with torch.distributed.run (one process per GPU)")
    if not torch.cuda.is_available():
        raise SystemExit("no CUDA device: the render path has no CPU fallback")
    torch.cuda.set_device(local)
    # stdout carries exactly one JSON line: native libraries (NCCL prints its version banner on stdout)
    # get stderr for the duration of the run
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    group = D.init_process_group(world, rank, local) if world > 1 else None

    global WIDTH, HEIGHT, ITERATIONS, WORKLOAD
    if args.size or args.iterations_per_gpu or args.preset != "poisson-saturne":      # other BASELINE configs (profiles/, not the default line)
        if args.size:
            WIDTH, HEIGHT = (int(v) for v in args.size.lower().split("x"))
        ITERATIONS = int(float(args.iterations_per_gpu)) if args.iterations_per_gpu else ITERATIONS
        WORKLOAD = f"{args.preset}, {ITERATIONS:.3g} iterations per GPU, {WIDTH}x{HEIGHT} (non-default)"
    cfg = S.Config.poisson_saturne() if args.preset == "poisson-saturne" else S.Config.solar_sail()
    if args.preset != "poisson-saturne":
        cfg.angle = 3.839724354387525   # 220 degrees (BASELINE configs[2])
    cfg.width, cfg.height, cfg.transparent = WIDTH, HEIGHT, False
    strong = args.scaling == "strong"
    if strong:                                   # fixed total work: ITERATIONS over the whole job, 1/N of it per GPU
        ITERATIONS = ITERATIONS // world
    lanes = args.lanes or 0
    jpt = args.jobs_per_thread
    frame = D.Frame(cfg, device=local, world=world, rank=rank, group=group, lanes=lanes, jobs_per_thread=jpt,
                    iterations_per_gpu=ITERATIONS, seed=SEED)
    recorded_per_step = frame.recorded_iterations_total()   # over all ranks

    stream = torch.cuda.Stream(device=local)
    sp = C.c_void_p(stream.cuda_stream)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=f"cuda:{local}")   # > 126 MB L2

    def sync_all():
        torch.cuda.synchronize(local)
        if group is not None:
            D.barrier(group)
        torch.cuda.synchronize(local)

    def timed_steps(fn, k):
        """k steps, each bracketed by events on `stream`; L2 flushed between steps (outside the brackets)."""
        total = 0.0
        for _ in range(k):
            with torch.cuda.stream(stream):
                flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            sync_all()
            e0.record(stream)
            fn()
            e1.record(stream)
            sync_all()
            total += e0.elapsed_time(e1)
        return total   # ms

    # ---- value: device-resident frame -------------------------------------------------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()                          # sampled from the warm-up steps (same load) through the timed region
        time.sleep(0.3)
    for _ in range(args.warmup):
        frame.step_device(sp)
    sync_all()
    launches0 = L.sar_launch_count()
    ms_total = timed_steps(lambda: frame.step_device(sp), args.steps)
    launches = int(L.sar_launch_count() - launches0)
    clocks = sampler.stop() if rank == 0 else None
    ms_total = D.max_over_ranks(ms_total, group, local)
    value = recorded_per_step * args.steps / (ms_total * 1e-3)

    # ---- roofline: the iterate kernel alone ---------------------------------------------------
    it_ms = 0.0
    for _ in range(args.steps):
        frame.reset_async(sp)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sync_all()
        e0.record(stream)
        frame.render_async(sp)
        e1.record(stream)
        sync_all()
        it_ms += e0.elapsed_time(e1)
    it_ms /= args.steps
    peak, peak_src = measured_hbm_peak()
    achieved = frame.recorded_iterations_local() * ALGO_BYTES_PER_ITER / (it_ms * 1e-3) / 1e9
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            traffic = json.load(f).get("iterate_kernel_dram_bytes_per_launch")
    except Exception:
        pass

    # ---- e2e: through sar_render_parallel with host buffers ----------------------------------
    e2e = frame.make_e2e()   # pinned host start points + pinned host image
    for _ in range(min(args.warmup, 3)):
        e2e.step()
    sync_all()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e.step()
    sync_all()
    e2e_s = D.max_over_ranks(time.perf_counter() - t0, group, local)
    e2e_value = recorded_per_step * args.steps / e2e_s

    out = None
    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "width": WIDTH, "height": HEIGHT, "iterations_per_gpu": ITERATIONS,
                       "iterations_per_frame": ITERATIONS * world,
                       "lanes_per_gpu": frame.lanes, "jobs_per_thread": jpt, "iterations_per_job": frame.iterations_per_job,
                       "warmup_iterations_per_job": 1000, "seed": SEED,
                       "l2": "256 MB buffer written between timed steps (L2 flush); accumulators (96 MB) are rewritten by reset each step",
                       "parallelism": "1 GPU" if world == 1 else f"{world} ranks: trajectory-sharded, row-stripe merge over NVLink peer loads"},
            "clocks": clocks, "gpu_launches": launches,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": e2e.h2d_bytes * world, "d2h_bytes_per_step": e2e.d2h_bytes,
                    "ms_per_step": 1e3 * e2e_s / args.steps, "api": "sar_render_parallel (C ABI) with pinned host buffers" if world == 1
                    else "dist.Frame.step_host: per-rank C-ABI calls + NVLink stripe merge, image gathered on rank 0"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of this kernel from the round's ncu --set full capture "
                                           "(profiles/roofline_traffic.json, profiles/r2_iterate_ncu_full.md); per launch, not re-measured per run",
                         "kernel": "iterate_kernel", "kernel_ms": it_ms, "algorithmic_bytes_per_iteration": ALGO_BYTES_PER_ITER,
                         "peak_source": peak_src,
                         "note": "the working set is L2-resident, so HBM is not what binds: the ceiling of this kernel is the rate of "
                                 "scattered L2 atomics with return (one per recorded iteration), measured 128e9/s raw on this GPU "
                                 "(profiles/r1_micro_atomics.md); see l2_atomic",
                         "l2_atomic": {"bound": "l2 atomic with return, 1 per recorded iteration", "peak": 127.9e9, "unit": "ops/s",
                                       "achieved": frame.recorded_iterations_local() / (it_ms * 1e-3),
                                       "frac": frame.recorded_iterations_local() / (it_ms * 1e-3) / 127.9e9,
                                       "peak_source": "tools/micro_atomics.cu, ATOM.ADD.64 uniform random over 32 MB, 1x B200"}},
        }
    frame.check_sync()
    frame.close()
    # ---- parity of the N-rank path against the oracle (small frames, outside the timed regions) ------
    if not args.no_parity:
        p1 = parity_frame(world, rank, local, group, "solar", depth=False)
        p2 = parity_frame(world, rank, local, group, "poisson", depth=True)
        if rank == 0:
            out["parity"] = {k: bool(p1[k] and p2[k]) for k in ("count", "zbuf", "steps", "image")}
            out["parity"].update({"vs": "oracle", "frames": [p1, p2]})
    if rank == 0 and world == 1 and not args.no_parity and "parity" in out:
        try:   # one more frame, the README's own command, against the reference's published image
            rcfg = S.Config.poisson_saturne()
            rcfg.width, rcfg.height, rcfg.iterations, rcfg.transparent = 1920, 1080, 1_000_000_000, False
            rcfg.colors.brighness.offset = -0.25
            rr = S.ParallelRenderer.new(devices=[local])
            S.render_parallel(rr, rcfg, 1, seed=SEED)
            rcount, rsteps, _rz, rmx = rr.runtime().download()
            rr.shutdown()
            out["parity"]["reference_image"] = reference_image_check(rcount, rsteps, rmx)
        except Exception as e:  # never lose the bench line over the extra check
            out["parity"]["reference_image"] = {"error": str(e)[:300]}
    if rank == 0 and world == 1 and not args.no_cpu_baseline and not (args.size or args.iterations_per_gpu or args.preset != "poisson-saturne"):
        threads = os.cpu_count() or 8
        iters = cpu_sample_size(threads)
        v, dt = cpu_render_parallel(iters, threads)
        out["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                               "sample": f"{iters:.3g} of 1e9 iterations ({dt:.1f} s), render_parallel semantics: {threads} threads x 12 jobs"}
    if group is not None:
        D.shutdown(group)
    sys.stdout.flush()
    os.dup2(real_stdout, 1)
    os.close(real_stdout)
    if rank == 0:
        print(json.dumps(out), flush=True)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--lanes", type=int, default=0, help="trajectory lanes per GPU (0 = library default, SM count x 896)")
    ap.add_argument("--jobs-per-thread", type=int, default=1)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the N-rank-vs-oracle parity frames")
    ap.add_argument("--scaling", choices=["weak", "strong"], default="weak",
                    help="weak (default): 1e9 iterations per GPU; strong: 1e9 iterations per frame, split over the GPUs")
    ap.add_argument("--preset", choices=["poisson-saturne", "solar-sail"], default="poisson-saturne")
    ap.add_argument("--size", default="", help="WxH override (default 2048x2048)")
    ap.add_argument("--iterations-per-gpu", default="", help="override of 1e9")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
