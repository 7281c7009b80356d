#!/bin/bash
# A/B of an environment switch on the resident-input step: tools/ab.sh VAR   (runs bench twice: VAR unset, VAR=1)
v=$1
for mode in off on; do
  if [ $mode = on ]; then export $v=1; else unset $v; fi
  timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-torch-gpu --no-e2e > gpurun_out/ab_$mode.json 2> gpurun_out/ab_$mode.err
  python -c "
import json; d=json.loads(open('gpurun_out/ab_$mode.json').read().strip().splitlines()[-1]); print('$v $mode: %.2f vol/s %.3f ms/step' % (d['value'], d['ms_per_step']))"
done
