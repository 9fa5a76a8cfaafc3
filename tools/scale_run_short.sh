#!/bin/bash
# N = 4, 8 of bench.py and BASELINE configs[3] on 8 GPUs (the rest of tools/scale_run.sh, shortened)
set -u
mkdir -p gpurun_out
for n in 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
  python -c "
import json; d=json.loads(open('gpurun_out/scale_n$n.json').read().strip().splitlines()[-1]); print('N=$n value %.4g it/s  ms/step %.3f  e2e %.4g' % (d['value'], d['ms_per_step'], d['e2e']['value']))" || tail -5 gpurun_out/scale_n$n.err
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --steps 5 --warmup 3 --size 4096x4096 > gpurun_out/cfg3_8gpu.json 2> gpurun_out/cfg3_8gpu.err
python -c "
import json; d=json.loads(open('gpurun_out/cfg3_8gpu.json').read().strip().splitlines()[-1]); print('cfg3 8e9 4096^2 8 GPUs: %.4g it/s, %.3f ms/frame, e2e %.4g' % (d['value'], d['ms_per_step'], d['e2e']['value']))" || tail -5 gpurun_out/cfg3_8gpu.err
