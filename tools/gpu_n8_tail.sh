#!/bin/bash
# 8 GPUs: the shower record with the host-driven iterations above 1M tracks (default) and with the whole loop as graphs
mkdir -p gpurun_out; rm -f gpurun_out/n8_tail.log
for tb in 1048576 100000000 1048576 100000000; do
  G4HB200_TAIL_BELOW=$tb python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 3 --warmup 3 --e2e-steps 1 --sustained-seconds 0 --no-variants --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('tail_below=$tb shower ms', d['shower']['ms'], 'showers/s', d['shower']['showers_per_s'])" >> gpurun_out/n8_tail.log
done
cat gpurun_out/n8_tail.log
