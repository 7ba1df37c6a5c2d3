#!/bin/bash
# ncu --set full of the FINAL round-2 kernels of one MG-CG iteration at 7.35 M cells (summarised on the box).
cd "$(dirname "$0")/.."
python scripts/ncu_step.py --levels 3 > gpurun_out/ncu_final_plain.log 2>&1 || { tail -n 5 gpurun_out/ncu_final_plain.log; exit 1; }
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:'k_tangent|k_post|k_commit|k_mg_ebe|k_mg_cheb|k_mg_cg_update|k_mg_dot|k_mg_cg_p|k_mg_resid|k_mg_restrict|k_mg_prolong|k_mg_coarse_fused' -c 60 \
    -o gpurun_out/r2_final_l3 -f python scripts/ncu_step.py --levels 3 > gpurun_out/ncu_final.log 2>&1
tail -n 2 gpurun_out/ncu_final.log
python scripts/ncu_table.py <(ncu -i gpurun_out/r2_final_l3.ncu-rep --page raw --csv 2>/dev/null) > gpurun_out/r2_final_l3_table.txt 2>&1
ncu -i gpurun_out/r2_final_l3.ncu-rep --page raw --csv 2>/dev/null | gzip -9 > gpurun_out/r2_final_l3_raw.csv.gz
rm -f gpurun_out/r2_final_l3.ncu-rep
head -40 gpurun_out/r2_final_l3_table.txt
