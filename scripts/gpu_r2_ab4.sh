#!/bin/bash
cd "$(dirname "$0")/.."
bash scripts/gpu_variants.sh geom_ python bench.py --levels 4 --steps 2 --warmup 2 --pc mg --no-cpu-baseline --no-fallback --no-e2e 2>&1 | tee gpurun_out/r2_ab4_geom_l4.txt
bash scripts/gpu_variants.sh geom_ python bench.py --levels 3 --steps 3 --warmup 2 --pc mg --no-cpu-baseline --no-fallback --no-e2e 2>&1 | tee gpurun_out/r2_ab4_geom_l3.txt
