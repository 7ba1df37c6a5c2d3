#!/bin/bash
# A/B runs of everything that round 1 could only verify on the host emulation.
# Usage on the GPU box (one GPU):   gpurun --timeout 1500 -- 'bash scripts/gpu_ab.sh'
# Writes one JSON line per configuration to gpurun_out/ab_*.json, then the launch list and one --set full capture of
# the finest-level operator inside the multigrid loop.  Numbers printed under ncu are never bench values.
set -u
mkdir -p gpurun_out
run() {  # name, bench flags...
  local name=$1; shift
  timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err
  echo "$name rc=$? $(head -c 300 gpurun_out/ab_$name.json)"
}
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
run default          --pc mg                               # extrapolated guess + lagged setup (the defaults)
run plain_warmstart  --pc mg --warm-start 1 --mg-lag 0     # what profiles/r1_bench_mg_levels3.json measured, + new coarse solve
run extrap_only      --pc mg --warm-start 2 --mg-lag 0
run lag_only         --pc mg --warm-start 1 --mg-lag 2
SIC_MG_FUSED_COARSE=1 run fused_coarse   --pc mg            # opt-in: coarsest-level sweep as one cooperative launch
run jacobi_extrap    --pc jacobi --levels 2                # block-Jacobi CG with the extrapolated guess, 918k cells
run jacobi_plain     --pc jacobi --levels 2 --warm-start 1
run levels4          --pc mg --levels 4 --steps 2 --warmup 2   # 58.8M cells (about 50 GB)
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_mg|k_tangent|k_post|k_commit|k_diag|k_invert|k_ebe|k_guess" \
    -s 600 -c 900 --csv --log-file gpurun_out/r2_launches_bench_mg_levels3.csv \
    timeout 300 python bench.py --pc mg --no-cpu-baseline --no-e2e --steps 1 --warmup 1 > gpurun_out/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_mg_ebe_dot -s 20 -c 1 -o gpurun_out/r2_k_mg_ebe_dot \
    timeout 300 python bench.py --pc mg --levels 2 --no-cpu-baseline --no-e2e --steps 1 --warmup 1 > gpurun_out/ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_tangent -s 3 -c 1 -o gpurun_out/r2_k_tangent_set0 \
    timeout 300 python bench.py --pc mg --levels 2 --no-cpu-baseline --no-e2e --steps 1 --warmup 1 > gpurun_out/ncu_tangent.log 2>&1
ls -la gpurun_out
