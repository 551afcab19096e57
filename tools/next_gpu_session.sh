#!/bin/bash
# First GPU call of the next session (one B200, under gpurun, ~6 minutes): everything that changed after the last measurement
# of round 1 was validated on the emulated kernels only (tests/emu) — time it, profile it, and fuzz it on the hardware.
#   gpurun --timeout 600 -- 'bash tools/next_gpu_session.sh'
mkdir -p gpurun_out
(time timeout 400 python -m pytest tests -m gpu -x -q) > gpurun_out/r2_pytest_gpu.log 2>&1; tail -3 gpurun_out/r2_pytest_gpu.log
timeout 200 python bench.py --no-cpu > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; cut -c1-400 gpurun_out/r2_bench_n1.json; tail -2 gpurun_out/r2_bench_n1.err
# launch list of the bench command (committed under profiles/: the one of the final round-1 code was lost)
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 6 --no-cpu --depth 1 > gpurun_out/r2_ncu_launches.log 2>&1
# full captures: scan (gate merge, fold, pack), finalize (shorter load chain), resolve
for k in scan_kernel finalize_kernel resolve_kernel; do
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:^$k -s 3 -c 1 -f -o gpurun_out/r2_$k python tools/gpu_timing.py > gpurun_out/r2_ncu_$k.log 2>&1; tail -1 gpurun_out/r2_ncu_$k.log
done
timeout 100 python tools/gpu_timing.py 2>&1 | tail -2 | cut -c1-220
# scan grid smaller than the chip: does stage B of the step before overlap the scan then?
bash tools/gpu_scan_sms.sh
# the fuzzer on the hardware (same cases the emulated kernels passed)
timeout 300 python tools/emu_fuzz.py --emu 0 --seed 31 --cases 300 2>&1 | tail -3
ls -la gpurun_out | head -40
