#!/bin/bash
# one GPU-box visit: parity suite, bench, ncu launch list of the same bench command, full ncu captures
# usage: bash tools/gpu_round.sh [full]
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/pytest_gpu.log
python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_reference.json 2>> gpurun_out/bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"^(El|Gamma)" -s 44 -c 88 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 5 --warmup 3 --e2e-steps 2 --no-cpu-baseline --sustained-seconds 0 --shower-primaries 0 --no-variants > gpurun_out/bench_under_ncu.log 2>&1
python tools/bench_configs.py > gpurun_out/configs01.jsonl 2>> gpurun_out/bench.err
python tools/bench_shower.py --config 3 > gpurun_out/configs34.jsonl 2>> gpurun_out/bench.err
python tools/bench_shower.py --config 4 --primaries 256 >> gpurun_out/configs34.jsonl 2>> gpurun_out/bench.err
python tools/bench_shower.py --config 4 --primaries 4096 >> gpurun_out/configs34.jsonl 2>> gpurun_out/bench.err
python tools/bench_shower.py --config 4 --primaries 16384 >> gpurun_out/configs34.jsonl 2>> gpurun_out/bench.err
if [ "$1" == "full" ]; then
G4HB200_SPLIT_PARTS=1 ncu --set full --clock-control none --import-source on -k regex:"^(El|Gamma)" -c 16 -f -o gpurun_out/prof_pipeline \
    python tools/kernel_probe.py 1048576 1 > gpurun_out/prof_pipeline.log 2>&1
# the lane-refill executor (opt-in) and the single-precision SampleMSC (offered variant): lanes and durations for the record
G4HB200_REFILL=4 G4HB200_MSC_F32=1 G4HB200_SPLIT_PARTS=1 ncu --set full --clock-control none --import-source on -k regex:"(Refill|F32Kernel)" -c 9 -f -o gpurun_out/prof_variants \
    python tools/kernel_probe.py 1048576 1 > gpurun_out/prof_variants.log 2>&1
fi
tail -5 gpurun_out/pytest_gpu.log; cat gpurun_out/bench.json | cut -c1-600; tail -3 gpurun_out/bench.err; cat gpurun_out/configs34.jsonl | cut -c1-400
