#!/bin/bash
# tensor-pipe / DRAM counters of the fprop + dgrad launches of one eager training step: tools/ncu_conv.sh <tag>
tag=$1
cmd="python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-torch-gpu --no-cuda-graph --no-e2e"
ncu -k regex:"conv3d_igemm_kernel|conv3d_slab_kernel" -c 26 --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
    --clock-control none --csv --log-file gpurun_out/${tag}_conv_metrics.csv $cmd > gpurun_out/${tag}_ncu_conv.log 2>&1
echo "ncu rc=$?"
