#!/bin/bash
# ncu --set full captures of the hot kernels at the bench size (run under gpurun).  The reports are summarised ON THE BOX
# (scripts/ncu_summary.py + the raw metric page as gzipped CSV) and deleted: gpurun_out/ only travels back below 64 MiB.
cd "$(dirname "$0")/.."
L=${1:-3}
summarise() {   # $1 = report without extension
  python scripts/ncu_summary.py gpurun_out/$1.ncu-rep > gpurun_out/$1_summary.txt 2>&1
  ncu -i gpurun_out/$1.ncu-rep --page raw --csv 2>/dev/null | gzip -9 > gpurun_out/$1_raw.csv.gz
  ls -la gpurun_out/$1.ncu-rep; rm -f gpurun_out/$1.ncu-rep
}
python scripts/ncu_step.py --levels $L > gpurun_out/ncu_step_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:'k_tangent|k_post|k_commit|k_mg_ebe|k_mg_cheb_step|k_mg_cg_update|k_mg_coarse_fused|k_mg_restrict|k_mg_prolong' -c 70 \
    -o gpurun_out/r2_step_l$L -f python scripts/ncu_step.py --levels $L > gpurun_out/ncu_step.log 2>&1
tail -n 3 gpurun_out/ncu_step.log
summarise r2_step_l$L
python scripts/ncu_step.py --levels 2 --staged > gpurun_out/ncu_cfg3_plain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:'k_tangent|k_post|k_commit' -c 6 \
    -o gpurun_out/r2_cfg3_l2 -f python scripts/ncu_step.py --levels 2 --staged > gpurun_out/ncu_cfg3.log 2>&1
tail -n 3 gpurun_out/ncu_cfg3.log gpurun_out/ncu_cfg3_plain.log
summarise r2_cfg3_l2
# launch list of one whole bench step (kernel by kernel: --graph 0), for the kernels' SHARES of the step
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 40000 --csv --log-file gpurun_out/r2_launches_l$L.csv \
    python bench.py --levels $L --steps 1 --warmup 1 --pc mg --graph 0 --no-cpu-baseline --no-e2e --no-fallback > gpurun_out/ncu_launches.log 2>&1
gzip -9 -f gpurun_out/r2_launches_l$L.csv
du -sh gpurun_out
