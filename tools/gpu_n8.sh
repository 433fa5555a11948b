#!/bin/bash
# bench.py on 8 GPUs of one box (torchrun, as the driver launches it), with and without pinning the ranks to core slices
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
nvidia-smi -L > gpurun_out/gpus.txt 2>&1
lscpu | head -30 > gpurun_out/lscpu.txt 2>&1
python -c "import os; print('affinity', sorted(os.sched_getaffinity(0)))" >> gpurun_out/lscpu.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/bench_n8_aff.json 2> gpurun_out/bench_n8_aff.err
echo "rc=$? bytes=$(wc -c < gpurun_out/bench_n8_aff.json) gpus=$(wc -l < gpurun_out/gpus.txt)"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 20 --warmup 3 --no-affinity --sustained-seconds 0 --shower-primaries 0 > gpurun_out/bench_n8_noaff.json 2> gpurun_out/bench_n8_noaff.err
echo "rc=$? bytes=$(wc -c < gpurun_out/bench_n8_noaff.json)"
python - <<'PY'
import json
for f in ('bench_n8_aff','bench_n8_noaff'):
    try:
        d=json.load(open('gpurun_out/%s.json'%f))
        print(f, 'value %.3e e2e %.3e'%(d['value'], d['e2e']['value']), 'shower', (d.get('shower') or {}).get('ms'), (d.get('shower') or {}).get('value'), d.get('cpu_cores_of_rank0'))
    except Exception as e:
        print(f, 'failed', e)
PY
tail -12 gpurun_out/bench_n8_aff.err | cut -c1-250
