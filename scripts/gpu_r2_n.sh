#!/bin/bash
# N-GPU bench of the default configuration (what the driver's scaling run launches), plus the 2-rank tests when N == 2.
cd "$(dirname "$0")/.."
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
if [ "$N" == "2" ]; then
  timeout 900 python -m pytest tests/test_gpu_ranks.py -x -q -m gpu > gpurun_out/r2_ranks_test.log 2>&1; echo "ranks test rc=$?"
  tail -n 5 gpurun_out/r2_ranks_test.log
fi
timeout 900 $TR --master-port 29521 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r2_scale_n${N}.json 2> gpurun_out/r2_scale_n${N}.err; echo "n$N default rc=$?"
grep -h "ms/step\|OPERATOR\|probe\|rror\|safeincave_cuda" gpurun_out/r2_scale_n${N}.err | tail -n 8
free -g | head -2
