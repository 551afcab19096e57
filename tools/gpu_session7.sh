#!/bin/bash
mkdir -p gpurun_out
(time timeout 600 python -m pytest tests -m gpu -x -q) > gpurun_out/s7_pytest_gpu.log 2>&1; tail -4 gpurun_out/s7_pytest_gpu.log | head -2
python tools/gpu_modeac_timing.py 2>&1 | tail -1 | cut -c1-200
sed -i "s/synth.modeac_stream(900 + i, B \* BUF)/synth.config2_stream(900 + i, B * BUF)/" tools/gpu_modeac_timing.py; python tools/gpu_modeac_timing.py 2>&1 | tail -1 | cut -c1-200
timeout 200 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:modeac -s 12 -c 3 python tools/gpu_modeac_timing.py 2>&1 | grep -E "gpu__time|inst_executed" | head -6
