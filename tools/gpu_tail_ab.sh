#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/tail_ab.log
timeout 900 python -m pytest tests/test_shower.py tests/test_sharding.py -m gpu -x -q 2>&1 | tail -8 >> gpurun_out/tail_ab.log
for v in "G4HB200_GRAPH_TAIL=0" "G4HB200_TAIL_BELOW=32768" "G4HB200_TAIL_BELOW=262144" "G4HB200_TAIL_BELOW=1048576" "G4HB200_TAIL_BELOW=100000000"; do
  for p in 256 4096 16384; do
    echo "== $v primaries=$p" >> gpurun_out/tail_ab.log
    env $v python tools/bench_shower.py --config 4 --primaries $p 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print({k:d.get(k) for k in ('ms','value','loop_iterations_rank0','kernel_launches_rank0','peak_electrons_rank0','energy_balance')})
    else: print(l[:300])
" >> gpurun_out/tail_ab.log
  done
done
cat gpurun_out/tail_ab.log
