#!/bin/bash
# Second measurement pass: step-by-step launch lists (what a profiler's kernel serialisation matches), host topology, bench, variants.
mkdir -p gpurun_out
(nvidia-smi topo -m; lscpu | grep -i -E "numa|socket|model name|^cpu\(s\)"; python -c "import os; print('affinity', sorted(os.sched_getaffinity(0)))"; cat /sys/fs/cgroup/cpu.max; for d in /sys/bus/pci/devices/*; do if [ "$(cat $d/vendor)" = "0x10de" ]; then echo $d $(cat $d/numa_node) $(cat $d/class); fi; done) > gpurun_out/final_topology.txt 2>&1; head -30 gpurun_out/final_topology.txt
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final_launches_depth1_cold.csv python bench.py --steps 2 --warmup 1 --no-cpu --depth 1 > gpurun_out/final_ncu_d1_cold.log 2>&1
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 400 --csv --log-file gpurun_out/final_launches_depth1_warm.csv python bench.py --steps 2 --warmup 1 --no-cpu --depth 1 > gpurun_out/final_ncu_d1_warm.log 2>&1
timeout 300 python bench.py > gpurun_out/final2_bench_n1.json 2> gpurun_out/final2_bench_n1.err; cat gpurun_out/final2_bench_n1.json; tail -2 gpurun_out/final2_bench_n1.err
timeout 200 python bench.py --no-cpu --depth 1 > gpurun_out/final2_bench_depth1.json 2>/dev/null; cut -c1-160 gpurun_out/final2_bench_depth1.json
bash tools/gpu_variants.sh > gpurun_out/variants3.log 2>&1; grep -E "^==|passed|failed|cfg2 5" gpurun_out/variants3.log | cut -c1-200
