#!/bin/bash
# N=1: CUDA-graph tests and A/B timings, then the ncu captures.
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_gpu_mg.py -x -q -m gpu -k "graph or fused" > gpurun_out/r2_mg_test.log 2>&1; echo "mg test rc=$?"
tail -n 15 gpurun_out/r2_mg_test.log
B="python bench.py --levels 3 --steps 3 --warmup 3 --pc mg --no-cpu-baseline --no-fallback --no-e2e"
timeout 300 $B > gpurun_out/r2_ab_graph1_fused1.json 2> gpurun_out/r2_ab_graph1_fused1.err; echo "rc=$?"
timeout 300 $B --graph 0 > gpurun_out/r2_ab_graph0_fused1.json 2> gpurun_out/r2_ab_graph0_fused1.err; echo "rc=$?"
timeout 300 $B --fused-coarse 0 > gpurun_out/r2_ab_graph1_fused0.json 2> gpurun_out/r2_ab_graph1_fused0.err; echo "rc=$?"
grep -h "ms/step\|safeincave_cuda" gpurun_out/r2_ab_*.err
bash scripts/ncu_capture.sh 3
