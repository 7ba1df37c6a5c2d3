#!/bin/bash
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_gpu_mg.py -x -q -m gpu > gpurun_out/r2_ab2_test.log 2>&1; echo "mg tests rc=$?"
tail -n 4 gpurun_out/r2_ab2_test.log
bash scripts/gpu_variants.sh pc_ python bench.py --levels 3 --steps 2 --warmup 2 --pc mg --no-cpu-baseline --no-fallback --no-e2e 2>&1 | tee gpurun_out/r2_ab2_pc.txt
timeout 600 python bench.py --steps 2 --warmup 2 --pc mg --no-cpu-baseline --no-fallback --levels 4 > gpurun_out/r2_g_l4.json 2> gpurun_out/r2_g_l4.err; echo "L4 rc=$?"
grep -h "ms/step\|OPERATOR\|safeincave_cuda" gpurun_out/r2_g_l4.err
summarise() {
  python scripts/ncu_table.py <(ncu -i gpurun_out/$1.ncu-rep --page raw --csv 2>/dev/null) > gpurun_out/$1_table.txt 2>&1
  ncu -i gpurun_out/$1.ncu-rep --page raw --csv 2>/dev/null | gzip -9 > gpurun_out/$1_raw.csv.gz
  rm -f gpurun_out/$1.ncu-rep
}
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:'k_mg_ebe_pc|k_mg_cheb_step|k_mg_cheb_first|k_mg_cg_update|k_mg_restrict|k_mg_prolong' -c 12 \
    -o gpurun_out/r2_pc2_l3 -f python scripts/ncu_step.py --levels 3 > gpurun_out/ncu_pc2.log 2>&1
summarise r2_pc2_l3
cat gpurun_out/r2_pc2_l3_table.txt
