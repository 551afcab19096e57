#!/bin/bash
# Third measurement pass: launch lists with the receivers' address filters warm (the first steps of a context teach the filter
# every address, which makes stage B repeat its speculation: 0.27 ms instead of 0.07 ms), final capture of stage B, bench.
mkdir -p gpurun_out
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final3_launches.csv python bench.py --steps 2 --warmup 6 --no-cpu > gpurun_out/final3_ncu.log 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final3_launches_depth1.csv python bench.py --steps 2 --warmup 6 --no-cpu --depth 1 > gpurun_out/final3_ncu_d1.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:^resolve_kernel -s 3 -c 1 -f -o gpurun_out/resolve_final3 python tools/gpu_timing.py > gpurun_out/final3_ncu_resolve.log 2>&1; tail -1 gpurun_out/final3_ncu_resolve.log
timeout 300 python bench.py > gpurun_out/final3_bench_n1.json 2> gpurun_out/final3_bench_n1.err; cat gpurun_out/final3_bench_n1.json; tail -2 gpurun_out/final3_bench_n1.err
timeout 100 python tools/gpu_timing.py 2>&1 | tail -6 | cut -c1-200
(timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edges.py -m gpu -x -q) 2>&1 | tail -1
