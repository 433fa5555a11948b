python tools/msc_f32_probe.py 1048576 > gpurun_out/msc_f32_probe.json 2> gpurun_out/msc_f32_probe.err; cat gpurun_out/msc_f32_probe.json; tail -8 gpurun_out/msc_f32_probe.err
