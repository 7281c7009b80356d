#!/bin/bash
# ncu launch list (gpu__time_duration) of eager training steps; usage: tools/launch_list.sh <tag>
tag=$1
cmd="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-torch-gpu --no-cuda-graph --no-e2e"
$cmd > gpurun_out/${tag}_plain.json 2> gpurun_out/${tag}_plain.err || { echo plain failed; tail -5 gpurun_out/${tag}_plain.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/${tag}_launches.csv $cmd > gpurun_out/${tag}_ncu.log 2>&1
echo "ncu rc=$?"
python tools/summarize_launches.py gpurun_out/${tag}_launches.csv | tee gpurun_out/${tag}_summary.txt
