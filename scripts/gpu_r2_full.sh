#!/bin/bash
# N=1: the whole GPU suite, the default bench line (58.8 M cells, with CPU arm and parity check), a 7.35 M-cell line.
cd "$(dirname "$0")/.."
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2_gputest2.log 2>&1; echo "gpu tests rc=$?"
tail -n 6 gpurun_out/r2_gputest2.log
timeout 300 python bench.py --levels 3 --steps 3 --warmup 3 --pc mg --no-cpu-baseline --no-fallback --no-e2e > gpurun_out/r2_v_l3.json 2> gpurun_out/r2_v_l3.err; echo "L3 rc=$?"
grep -h "ms/step\|OPERATOR\|safeincave_cuda" gpurun_out/r2_v_l3.err
if [ "$1" == "default" ]; then
  timeout 1200 python bench.py > gpurun_out/r2_default.json 2> gpurun_out/r2_default.err; echo "default rc=$?"
  grep -h "ms/step\|OPERATOR\|CPU arm\|parity\|probe\|safeincave_cuda" gpurun_out/r2_default.err
else
  timeout 600 python bench.py --steps 2 --warmup 2 --pc mg --no-cpu-baseline --no-fallback --levels 4 > gpurun_out/r2_v_l4.json 2> gpurun_out/r2_v_l4.err; echo "L4 rc=$?"
  grep -h "ms/step\|OPERATOR\|safeincave_cuda" gpurun_out/r2_v_l4.err
fi
