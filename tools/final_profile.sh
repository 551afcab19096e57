#!/bin/bash
# Round-end measurement pass on one B200 (run under gpurun): tests, bench (both arms), ncu launch lists, full captures.
mkdir -p gpurun_out
(time timeout 500 python -m pytest tests -m gpu -x -q) > gpurun_out/final_pytest_gpu.log 2>&1; tail -4 gpurun_out/final_pytest_gpu.log
timeout 300 python bench.py > gpurun_out/final_bench_n1.json 2> gpurun_out/final_bench_n1.err; cut -c1-200 gpurun_out/final_bench_n1.json; tail -2 gpurun_out/final_bench_n1.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_bench_ref.json 2> gpurun_out/final_bench_ref.err; cut -c1-200 gpurun_out/final_bench_ref.json
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final_launches_cold.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/final_ncu_cold.log 2>&1
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 400 --csv --log-file gpurun_out/final_launches_warm.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/final_ncu_warm.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:^scan_kernel -s 3 -c 1 -f -o gpurun_out/scan_final python tools/gpu_timing.py > gpurun_out/final_ncu_scan.log 2>&1; tail -1 gpurun_out/final_ncu_scan.log
timeout 200 ncu --set full --clock-control none --import-source on -k regex:^resolve_kernel -s 3 -c 1 -f -o gpurun_out/resolve_final python tools/gpu_timing.py > gpurun_out/final_ncu_resolve.log 2>&1; tail -1 gpurun_out/final_ncu_resolve.log
timeout 100 python tools/gpu_timing.py 2>&1 | tail -2 | cut -c1-220
timeout 100 python tools/gpu_timing.py cfg5 2>&1 | tail -1 | cut -c1-220
for lib in readsb_b200/libb200demod_v*.so; do [ -f $lib ] && B200_DEMOD_LIB=$PWD/$lib timeout 100 python tools/gpu_timing.py 2>&1 | tail -1 | sed -e "s|^|$(basename $lib) |" | cut -c1-220; done
ls -la gpurun_out | head -40
