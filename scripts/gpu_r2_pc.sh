#!/bin/bash
# N=1: compressed preconditioner operator + constant-division change: parity tests, A/B timings, ncu of the changed kernels.
cd "$(dirname "$0")/.."
timeout 1200 python -m pytest tests/test_gpu_mg.py tests/test_gpu_constitutive.py -x -q -m gpu > gpurun_out/r2_pc_test.log 2>&1; echo "tests rc=$?"
tail -n 6 gpurun_out/r2_pc_test.log
B="python bench.py --steps 3 --warmup 3 --pc mg --no-cpu-baseline --no-fallback --no-e2e"
timeout 300 $B --levels 3 > gpurun_out/r2_pc1_l3.json 2> gpurun_out/r2_pc1_l3.err; echo "rc=$?"
timeout 300 $B --levels 3 --compressed 0 > gpurun_out/r2_pc0_l3.json 2> gpurun_out/r2_pc0_l3.err; echo "rc=$?"
timeout 600 python bench.py --steps 2 --warmup 2 --pc mg --no-cpu-baseline --no-fallback --levels 4 > gpurun_out/r2_pc1_l4.json 2> gpurun_out/r2_pc1_l4.err; echo "rc=$?"
grep -h "ms/step\|safeincave_cuda" gpurun_out/r2_pc*.err
summarise() {
  python scripts/ncu_table.py <(ncu -i gpurun_out/$1.ncu-rep --page raw --csv 2>/dev/null) > gpurun_out/$1_table.txt 2>&1
  ncu -i gpurun_out/$1.ncu-rep --page raw --csv 2>/dev/null | gzip -9 > gpurun_out/$1_raw.csv.gz
  rm -f gpurun_out/$1.ncu-rep
}
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:'k_tangent|k_commit|k_mg_ebe_pc|k_mg_ebe_dot|k_mg_ct_compress' -c 8 \
    -o gpurun_out/r2_pc_l3 -f python scripts/ncu_step.py --levels 3 > gpurun_out/ncu_pc.log 2>&1
tail -n 2 gpurun_out/ncu_pc.log
summarise r2_pc_l3
cat gpurun_out/r2_pc_l3_table.txt
for L in 1 0; do
  if timeout 300 python scripts/ncu_step.py --levels $L --staged > gpurun_out/ncu_cfg3_plain.log 2>&1; then
    timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off \
        -k regex:'k_tangent|k_post|k_commit' -c 6 \
        -o gpurun_out/r2_cfg3_l$L -f python scripts/ncu_step.py --levels $L --staged > gpurun_out/ncu_cfg3.log 2>&1
    summarise r2_cfg3_l$L
    cat gpurun_out/r2_cfg3_l${L}_table.txt
    break
  fi
  tail -n 2 gpurun_out/ncu_cfg3_plain.log
done
du -sh gpurun_out
