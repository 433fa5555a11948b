#!/bin/bash
# A/B: ramped chunks of the host-buffer step; small grids in the graph-driven tail of the stepping loop
mkdir -p gpurun_out; rm -f gpurun_out/ab2.log
for r in 0 1 0 1; do
  echo "== G4HB200_HOST_RAMP=$r" >> gpurun_out/ab2.log
  G4HB200_HOST_RAMP=$r python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 10 --shower-primaries 0 --sustained-seconds 0 --no-variants 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('e2e', d['e2e']['value'], 'value', d['value'])" >> gpurun_out/ab2.log
done
for g in 0 1; do
  echo "== G4HB200_SMALL_TAIL_GRIDS=$g" >> gpurun_out/ab2.log
  for p in 256 256 4096; do
    G4HB200_SMALL_TAIL_GRIDS=$g python tools/bench_shower.py --config 4 --primaries $p 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['primaries_total'], d['ms'], d['loop_iterations_rank0'])" >> gpurun_out/ab2.log
  done
done
cat gpurun_out/ab2.log
