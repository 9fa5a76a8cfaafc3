"""BASELINE configs[4]: solar-sail 360-frame angle sweep, 1e8 iterations per frame, 2048x2048, frames round-robin over
the visible GPUs in one process (sar_render_sequence), every frame copied to the host.
Usage: python tools/seq_bench.py [n_frames] [kinds]   (wall time for raw RGBA16 frames, device-converted RGB8 frames and — kinds
containing "png" — complete compressed RGB16 PNG files, main.rs:496-512 with the default branch of write_image_matches)"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import strange_attractor_renderer_b200 as S

n_frames = int(sys.argv[1]) if len(sys.argv) > 1 else 360
cfg = S.Config.solar_sail()
cfg.width = cfg.height = 2048
cfg.iterations = 100_000_000
angles = S.angle_iter(0.0, 360.0, 1.0)[:n_frames]
nd = torch.cuda.device_count()
for devs in ([0], list(range(nd))) if nd > 1 else ([0],):
    for shared in (False, True):
        for what in (sys.argv[2].split(",") if len(sys.argv) > 2 else ("rgba16", "rgb8")):
            r = S.ParallelRenderer.new(devices=devs)
            n = [0, 0]

            def cb(f, im):
                n[0] += 1
                n[1] += im.size

            def run(a):
                if what == "rgba16":
                    S.render_sequence(r, cfg, a, 1, seed=7, shared_points=shared, callback=cb)
                elif what == "png":
                    S.render_sequence_encoded(r, cfg, a, 1, S.PixelFormat.Rgb16, S.Container.PngDeflate, seed=7, shared_points=shared, callback=cb)
                else:
                    S.render_sequence_encoded(r, cfg, a, 1, S.PixelFormat.Rgb8, S.Container.Raw, seed=7, shared_points=shared, callback=cb)

            run(angles[:4 * len(devs)])
            n[0] = n[1] = 0
            t0 = time.perf_counter()
            run(angles)
            dt = time.perf_counter() - t0
            lanes = r.num_threads() // len(devs)
            rec = (100_000_000 // lanes) * lanes * len(angles)
            print(f"cfg4 sweep {len(angles)} frames on {len(devs)} GPU(s), {'shared' if shared else 'fresh'} points, {what}: {dt:.3f} s, "
                  f"{1e3 * dt / len(angles):.3f} ms/frame, {rec / dt:.4g} it/s, frames delivered {n[0]}, {n[1] / max(n[0], 1) / 1e6:.2f} MB per frame", flush=True)
            r.shutdown()
