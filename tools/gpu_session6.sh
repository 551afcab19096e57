#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edges.py -m gpu -x -q 2>&1 | tail -1
timeout 100 python tools/gpu_latency.py 2>&1 | grep -E "median_us|scan_kernel|stage_b|results_to|scan_begin|launches_per" 
timeout 100 python tools/gpu_timing.py cfg2 2>&1 | tail -2 | cut -c1-170
timeout 100 python tools/gpu_timing.py cfg5 2>&1 | tail -1 | cut -c1-170
timeout 200 python bench.py --no-cpu --no-e2e --no-extra --steps 5 --warmup 1 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('value', round(d['value']), 'ms/launch', d['ms_per_step']/96, 'burst', round(d['burst']['value']), 'scan', d['roofline']['kernel_ms_per_launch'])"
