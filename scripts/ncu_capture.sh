#!/bin/bash
# ncu --set full captures of the hot kernels at the bench size (run under gpurun; outputs in gpurun_out/).
cd "$(dirname "$0")/.."
L=${1:-3}
python scripts/ncu_step.py --levels $L > gpurun_out/ncu_step_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:'k_tangent|k_post|k_commit|k_mg_ebe|k_mg_cheb_step|k_mg_cg_update|k_mg_coarse_fused|k_mg_restrict|k_mg_prolong' -c 70 \
    -o gpurun_out/r2_step_l$L -f python scripts/ncu_step.py --levels $L > gpurun_out/ncu_step.log 2>&1
tail -3 gpurun_out/ncu_step.log
python scripts/ncu_step.py --levels 2 --staged > gpurun_out/ncu_cfg3_plain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:'k_tangent|k_post|k_commit' -c 6 \
    -o gpurun_out/r2_cfg3_l2 -f python scripts/ncu_step.py --levels 2 --staged > gpurun_out/ncu_cfg3.log 2>&1
tail -3 gpurun_out/ncu_cfg3.log gpurun_out/ncu_cfg3_plain.log
ls -la gpurun_out/*.ncu-rep
