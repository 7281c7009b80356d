#!/bin/bash
# A/B/C of B2_WGRAD_REDUCE modes on one box
for mode in inline side defer; do
  B2_WGRAD_REDUCE=$mode timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-torch-gpu --no-e2e > gpurun_out/ab_$mode.json 2> gpurun_out/ab_$mode.err
  python -c "
import json; d=json.loads(open('gpurun_out/ab_$mode.json').read().strip().splitlines()[-1]); print('$mode: %.2f vol/s %.3f ms/step' % (d['value'], d['ms_per_step']), d['clocks'])" || tail -3 gpurun_out/ab_$mode.err
done
