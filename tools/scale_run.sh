#!/bin/bash
# 1 -> 8 GPU scaling of bench.py (weak and strong), the N-rank parity tests and the 8-GPU BASELINE
# config; run under `gpurun --gpus 8`.  Every bench line carries "parity" (N-rank frame vs oracle).
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multi_gpu.py -x -q 2>&1 | tail -3
run() {  # $1 = N, $2 = tag, rest = extra bench args
  local n=$1 tag=$2; shift 2
  if [ "$n" = 1 ]; then
    timeout 300 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/${tag}_n$n.json 2> gpurun_out/${tag}_n$n.err
  else
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2950$n bench.py --gpus $n --steps 5 --warmup 3 "$@" > gpurun_out/${tag}_n$n.json 2> gpurun_out/${tag}_n$n.err
  fi
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${tag}_n$n.json").read().strip().splitlines()[-1])
    p=d.get("parity",{})
    print("${tag} N=$n value %.4g it/s  ms/step %.3f  e2e %.4g  parity %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], {k:p.get(k) for k in ("count","zbuf","steps","image")}))
except Exception as e:
    print("${tag} N=$n failed", e); print(open("gpurun_out/${tag}_n$n.err").read()[-1500:])
PY
}
for n in 1 2 4 8; do run $n weak; done
for n in 2 4 8; do run $n strong --scaling strong --no-parity; done
# BASELINE configs[3]: poisson-saturne, 8e9 iterations, 4096x4096, 8 GPUs row-striped
run 8 cfg3_4096 --size 4096x4096 --no-parity
# BASELINE configs[4]: 360-frame solar-sail sweep, 1e8 iterations per frame, frames round-robin over the GPUs of one process
timeout 300 python tools/seq_bench.py 360 2>&1 | grep cfg4
