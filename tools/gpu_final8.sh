#!/bin/bash
# final multi-GPU evidence on one 8-GPU box: NCCL DP test (2 ranks), training bench at 2/4/8 ranks, batched inference at 8
timeout 600 python -m pytest tests/test_gpu_dp.py -m gpu -q > gpurun_out/f8_dp_test.log 2>&1; echo "dp test rc=$?"; tail -2 gpurun_out/f8_dp_test.log
for N in 2 4 8; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --no-torch-gpu > gpurun_out/f8_bench_n$N.json 2> gpurun_out/f8_bench_n$N.err
  echo "bench N=$N rc=$?"
  python - <<P
import json
try:
    d=json.loads(open("gpurun_out/f8_bench_n$N.json").read().strip().splitlines()[-1])
    print("N=%d value %.1f vol/s (%.3f ms/step) e2e %.1f clocks %s" % (d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"], d["clocks"]))
except Exception as e:
    print("no line", e)
P
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 8 --inference --steps 4 > gpurun_out/f8_inference_n8.json 2> gpurun_out/f8_inference_n8.err
echo "inference N=8 rc=$?"; cut -c1-260 gpurun_out/f8_inference_n8.json
