#!/bin/bash
# N GPUs: the fused operator + halo exchange kernel: parity tests, then A/B against separate launches.
cd "$(dirname "$0")/.."
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 900 python -m pytest tests/test_gpu_ranks.py -x -q -m gpu > gpurun_out/r2_fx_test.log 2>&1; echo "ranks test rc=$?"
tail -n 12 gpurun_out/r2_fx_test.log
for FX in 1 0; do
  timeout 600 $TR --master-port 2953$FX bench.py --gpus $N --levels 3 --steps 3 --warmup 3 --pc mg --no-e2e --fused-exchange $FX > gpurun_out/r2_fx${FX}_n${N}_l3.json 2> gpurun_out/r2_fx${FX}_n${N}_l3.err; echo "L3 fx=$FX rc=$?"
  grep -h "ms/step\|OPERATOR\|rror" gpurun_out/r2_fx${FX}_n${N}_l3.err | tail -n 3
done
for FX in 1 0; do
  timeout 600 $TR --master-port 2954$FX bench.py --gpus $N --levels 4 --steps 3 --warmup 3 --pc mg --no-e2e --fused-exchange $FX > gpurun_out/r2_fx${FX}_n${N}_l4.json 2> gpurun_out/r2_fx${FX}_n${N}_l4.err; echo "L4 fx=$FX rc=$?"
  grep -h "ms/step\|OPERATOR\|rror" gpurun_out/r2_fx${FX}_n${N}_l4.err | tail -n 3
done
