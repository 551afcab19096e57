#!/bin/bash
mkdir -p gpurun_out
(time timeout 600 python -m pytest tests -m gpu -x -q) > gpurun_out/s5_pytest_gpu.log 2>&1; tail -3 gpurun_out/s5_pytest_gpu.log
timeout 400 python bench.py > gpurun_out/s5_bench_n1.json 2> gpurun_out/s5_bench_n1.err; tail -3 gpurun_out/s5_bench_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/s5_bench_n1.json').readline())
print({k:d[k] for k in ('value','ms_per_step','timed_region_s','parity_checked','gpu_launches') if k in d})
print('burst', d.get('burst'))
print('e2e', d.get('e2e'))
print('roofline', {k:d['roofline'][k] for k in ('frac','kernel_ms_per_launch','kernel_ms_per_launch_overlapped_with_stage_b')})
print('clocks', d.get('clocks'))
print(json.dumps(d.get('extra'))[:2500])
print('cpu', d.get('cpu_baseline'))
PY
timeout 120 python tools/gpu_scan_probe.py 2>&1 | tail -4
