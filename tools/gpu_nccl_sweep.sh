#!/bin/bash
N=${1:-8}
for ch in 1 2 4; do
  NCCL_MAX_NCHANNELS=$ch NCCL_MIN_NCHANNELS=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 30 --warmup 5 --no-cpu-baseline --no-torch-gpu --no-e2e > gpurun_out/sweep_ch$ch.json 2> gpurun_out/sweep_ch$ch.err
  python - <<P
import json
try:
    d=json.loads(open("gpurun_out/sweep_ch$ch.json").read().strip().splitlines()[-1])
    print("channels $ch: N=%d value %.1f vol/s (%.3f ms/step) clocks %s" % (d["n_gpus"], d["value"], d["ms_per_step"], d["clocks"]))
except Exception as e:
    print("channels $ch: no line", e)
P
done
