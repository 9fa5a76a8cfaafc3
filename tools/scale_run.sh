#!/bin/bash
# 1 -> 8 GPU scaling of bench.py plus the 8-GPU BASELINE configs; run under `gpurun --gpus 8`.
set -u
mkdir -p gpurun_out
python -m pytest tests/test_multi_gpu.py -x -q 2>&1 | tail -3
for n in 1 2 4 8; do
  if [ "$n" = 1 ]; then
    python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2950$n bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
  fi
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/scale_n$n.json"))
    print("N=$n value %.4g it/s  ms/step %.3f  e2e %.4g" % (d["value"], d["ms_per_step"], d["e2e"]["value"]))
except Exception as e:
    print("N=$n failed", e); print(open("gpurun_out/scale_n$n.err").read()[-1500:])
PY
done
# BASELINE configs[3]: poisson-saturne, 8e9 iterations, 4096x4096, 8 GPUs row-striped
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --steps 5 --warmup 3 --size 4096x4096 > gpurun_out/cfg3_8gpu.json 2> gpurun_out/cfg3_8gpu.err
python -c "
import json; d=json.load(open('gpurun_out/cfg3_8gpu.json')); print('cfg3 8e9 4096^2 8 GPUs: %.4g it/s, %.3f ms/frame, e2e %.4g' % (d['value'], d['ms_per_step'], d['e2e']['value']))" || tail -5 gpurun_out/cfg3_8gpu.err
# BASELINE configs[4]: 360-frame solar-sail sweep, 1e8/frame, 2048x2048, frames round-robin over 8 GPUs in one process
python - <<'PY'
import sys, time
sys.path.insert(0, ".")
import strange_attractor_renderer_b200 as S
cfg = S.Config.solar_sail(); cfg.width = cfg.height = 2048; cfg.iterations = 100_000_000
angles = S.angle_iter(0.0, 360.0, 1.0)
import torch
nd = torch.cuda.device_count()
for devs in ([0], list(range(nd))):
    for shared in (False, True):
        r = S.ParallelRenderer.new(devices=devs)
        S.render_sequence(r, cfg, angles[:2 * len(devs)], 1, seed=7, shared_points=shared, callback=lambda f, im: None)
        t0 = time.perf_counter(); n = [0]
        def cb(f, im): n[0] += 1
        S.render_sequence(r, cfg, angles, 1, seed=7, shared_points=shared, callback=cb)
        dt = time.perf_counter() - t0
        lanes = r.num_threads() // len(devs)
        rec = (100_000_000 // lanes) * lanes * len(angles)
        print(f"cfg4 sweep 360 frames on {len(devs)} GPU(s), {'shared' if shared else 'fresh'} points: {dt:.3f} s, {1e3*dt/360:.3f} ms/frame, {rec/dt:.4g} it/s, frames delivered {n[0]}", flush=True)
        r.shutdown()
PY
