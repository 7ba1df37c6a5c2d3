#!/bin/bash
# 2-GPU checks of the nested multigrid partition (run under `gpurun --gpus 2`); outputs in gpurun_out/.
cd "$(dirname "$0")/.."
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 900 python -m pytest tests/test_gpu_ranks.py -x -q -m gpu > gpurun_out/r2_ranks_test.log 2>&1; echo "ranks test rc=$?"
tail -5 gpurun_out/r2_ranks_test.log
timeout 600 python bench.py --levels 3 --steps 3 --warmup 3 --pc mg --no-cpu-baseline --no-fallback > gpurun_out/r2_n1_l3_nested.json 2> gpurun_out/r2_n1_l3_nested.err; echo "n1 rc=$?"
timeout 600 $TR --master-port 29511 bench.py --gpus $N --levels 3 --steps 3 --warmup 3 --pc mg > gpurun_out/r2_n${N}_l3_nested.json 2> gpurun_out/r2_n${N}_l3_nested.err; echo "n$N nested rc=$?"
timeout 600 $TR --master-port 29512 bench.py --gpus $N --levels 3 --steps 3 --warmup 3 --pc mg --nested 0 > gpurun_out/r2_n${N}_l3_repl.json 2> gpurun_out/r2_n${N}_l3_repl.err; echo "n$N replicated rc=$?"
timeout 900 $TR --master-port 29513 bench.py --gpus $N --levels 4 --steps 2 --warmup 2 --pc mg > gpurun_out/r2_n${N}_l4_nested.json 2> gpurun_out/r2_n${N}_l4_nested.err; echo "n$N L4 rc=$?"
grep -h "ms/step" gpurun_out/r2_n*.err
