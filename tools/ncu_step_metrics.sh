#!/bin/bash
# ncu DRAM / tensor-pipe counters for EVERY kernel of one eager training step (BASELINE shape):
#   tools/ncu_step_metrics.sh <tag>   ->  gpurun_out/<tag>_step_metrics.csv
tag=$1
cmd="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-torch-gpu --no-cuda-graph --no-e2e"
$cmd > gpurun_out/${tag}_plain.json 2> gpurun_out/${tag}_plain.err || { echo plain failed; tail -5 gpurun_out/${tag}_plain.err; exit 1; }
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed \
    --clock-control none -c 4000 --csv --log-file gpurun_out/${tag}_step_metrics.csv $cmd > gpurun_out/${tag}_ncu.log 2>&1
echo "ncu rc=$?"
python tools/summarize_step_metrics.py gpurun_out/${tag}_step_metrics.csv | tee gpurun_out/${tag}_step_metrics_summary.txt | head -60
