#!/bin/bash
# One gpurun call of the build->measure loop: GPU tests, bench arms, per-kernel timings.  Logs go to gpurun_out/<tag>_*.
tag=${1:-run}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/${tag}_build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/${tag}_build.log; exit 1; }
timeout 900 python -m pytest tests -m gpu -q -s > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${tag}_pytest.log
grep -E "passed|failed|error" gpurun_out/${tag}_pytest.log | tail -3
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/${tag}_bench.json
if [ "${2:-}" != "short" ]; then
  timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${tag}_ref.json 2> gpurun_out/${tag}_ref.err; echo "ref rc=$?"
  timeout 300 python bench.py --inference --steps 4 > gpurun_out/${tag}_inf.json 2> gpurun_out/${tag}_inf.err; echo "inf rc=$?"; cat gpurun_out/${tag}_inf.json
  timeout 300 python tools/bench_kernels.py > gpurun_out/${tag}_kernels.log 2>&1; echo "kernels rc=$?"
fi
