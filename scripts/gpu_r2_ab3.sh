#!/bin/bash
cd "$(dirname "$0")/.."
bash scripts/gpu_variants.sh pcmath_ python bench.py --levels 3 --steps 3 --warmup 2 --pc mg --no-cpu-baseline --no-fallback --no-e2e 2>&1 | tee gpurun_out/r2_ab3_pcmath.txt
timeout 600 python -m pytest tests/test_gpu_mg.py -x -q -m gpu > gpurun_out/r2_ab3_test.log 2>&1; echo "mg tests rc=$?"
tail -n 3 gpurun_out/r2_ab3_test.log
