#!/bin/bash
# multi-GPU checks: NCCL DP test + bench at N ranks, NCCL captured in the graph vs segmented replay
N=${1:-2}
tag=${2:-dp}
run_bench() {  # $1 = label, env passed through
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --no-torch-gpu > gpurun_out/${tag}_$1.json 2> gpurun_out/${tag}_$1.err
  echo "bench $1 rc=$?"
  python - <<P
import json
try:
    d=json.loads(open("gpurun_out/${tag}_$1.json").read().strip().splitlines()[-1])
    print("$1: N=%d value %.1f vol/s (%.3f ms/step) e2e %.1f clocks %s" % (d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"], d["clocks"]))
except Exception as e:
    print("$1: no line", e)
P
}
if [ "${3:-}" != "nobench_tests" ]; then timeout 600 python -m pytest tests/test_gpu_dp.py -m gpu -q -x 2>&1 | tail -5; fi
if [ "${3:-}" != "nobench_tests" ]; then B2_DP_CAPTURE_NCCL=0 timeout 600 python -m pytest tests/test_gpu_dp.py -m gpu -q -x 2>&1 | tail -3; fi
run_bench captured
B2_DP_CAPTURE_NCCL=0 run_bench segmented
tail -3 gpurun_out/${tag}_captured.err
