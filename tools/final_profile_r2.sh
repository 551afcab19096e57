#!/bin/bash
# Round 2: every profile that DESIGN.md / profiles/README.md cite, taken on the code as committed (one B200, under gpurun).
mkdir -p gpurun_out
# 1. launch list of the bench command (step by step: a profiler serialises the kernels anyway), kernel durations only
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 1 --warmup 1 --launches-per-step 6 --no-cpu --no-extra --depth 1 > gpurun_out/r02_ncu_launches.log 2>&1
# 2. full captures: the warm many-receiver step (scan, stage B, finalizer), the Mode A/C scan, the one-receiver kernels
for k in scan_kernel resolve_kernel finalize_kernel; do
  timeout 250 ncu --set full --clock-control none --import-source on -k regex:^$k -s 4 -c 1 -f -o gpurun_out/r02_$k python tools/gpu_timing.py > gpurun_out/r02_ncu_$k.log 2>&1; tail -1 gpurun_out/r02_ncu_$k.log
done
timeout 250 ncu --set full --clock-control none --import-source on -k regex:^modeac_scan -s 3 -c 1 -f -o gpurun_out/r02_modeac_scan python tools/gpu_modeac_timing.py > gpurun_out/r02_ncu_modeac.log 2>&1; tail -1 gpurun_out/r02_ncu_modeac.log
timeout 250 ncu --set full --clock-control none --import-source on -k regex:resolve_solo -s 200 -c 1 -f -o gpurun_out/r02_resolve_solo python tools/gpu_latency.py > gpurun_out/r02_ncu_solo.log 2>&1; tail -1 gpurun_out/r02_ncu_solo.log
timeout 250 ncu --set full --clock-control none --import-source on -k regex:^scan_kernel -s 200 -c 1 -f -o gpurun_out/r02_scan_one_buffer python tools/gpu_latency.py > gpurun_out/r02_ncu_scan1.log 2>&1; tail -1 gpurun_out/r02_ncu_scan1.log
timeout 250 ncu --set full --clock-control none --import-source on -k regex:^scan_kernel -s 4 -c 1 -f -o gpurun_out/r02_scan_dense python tools/gpu_timing.py cfg5 > gpurun_out/r02_ncu_scan5.log 2>&1; tail -1 gpurun_out/r02_ncu_scan5.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:modeac -c 12 --csv --log-file gpurun_out/r02_modeac_launches.csv python tools/gpu_modeac_timing.py > /dev/null 2>&1
# 3. the numbers themselves (never taken under the profiler): clocks beside them
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/r02_clocks.csv &
SMI=$!
timeout 500 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; tail -2 gpurun_out/r02_bench_n1.err
kill $SMI
timeout 100 python tools/gpu_timing.py 2>&1 | tail -2 | cut -c1-200 > gpurun_out/r02_timing_cfg2.txt
timeout 100 python tools/gpu_timing.py cfg5 2>&1 | tail -2 | cut -c1-200 > gpurun_out/r02_timing_cfg5.txt
timeout 100 python tools/gpu_scan_probe.py 2>&1 | tail -4 > gpurun_out/r02_scan_probe.txt
timeout 100 python tools/gpu_modeac_timing.py 2>&1 | tail -1 | cut -c1-220 > gpurun_out/r02_modeac_timing.txt
cat gpurun_out/r02_scan_probe.txt gpurun_out/r02_modeac_timing.txt; cut -c1-400 gpurun_out/r02_bench_n1.json
timeout 200 python tools/gpu_latency.py > gpurun_out/r02_latency.json 2> gpurun_out/r02_latency.err; head -c 600 gpurun_out/r02_latency.json
# 4. parity on this box: the gpu-marked tests, then 300 fuzz cases against the oracle on the hardware (tools/emu_fuzz.py without --emu)
( time timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 ) > gpurun_out/r02_pytest_gpu.log 2>&1; tail -6 gpurun_out/r02_pytest_gpu.log
timeout 900 python tools/emu_fuzz.py --seed 2024 --cases 300 2>&1 | tail -2 | tee gpurun_out/r02_hw_fuzz.txt
# 5. compute-sanitizer (memcheck, racecheck, synccheck) over a selection of the gpu-marked tests
bash tools/gpu_sanitize.sh
