#!/bin/bash
# N=1: A/B of the k_mg_ebe_pc occupancy variants and the k_tangent variants, then the default bench at 58.8 M cells.
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_gpu_mg.py -x -q -m gpu -k "compressed or graph" > gpurun_out/r2_ab_test.log 2>&1; echo "tests rc=$?"
tail -n 4 gpurun_out/r2_ab_test.log
bash scripts/gpu_variants.sh pc_ python bench.py --levels 3 --steps 2 --warmup 2 --pc mg --no-cpu-baseline --no-fallback --no-e2e 2>&1 | tee gpurun_out/r2_ab_pc.txt
bash scripts/gpu_variants.sh t_ python scripts/time_constitutive.py --levels 3 2>&1 | tee gpurun_out/r2_ab_tangent.txt
timeout 600 python bench.py --steps 2 --warmup 2 --pc mg --no-cpu-baseline --no-fallback --levels 4 > gpurun_out/r2_pc1_l4.json 2> gpurun_out/r2_pc1_l4.err; echo "L4 rc=$?"
grep -h "ms/step\|OPERATOR\|safeincave_cuda" gpurun_out/r2_pc1_l4.err
