#!/bin/bash
# Round 2, second measurement session: tiles shared between warps in small runs (B200_SCAN_SUB), the Mode A/C scan / walk changes.
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 ) > gpurun_out/s8_pytest.log 2>&1
tail -5 gpurun_out/s8_pytest.log
for sub in 0 2 1; do
  B200_SCAN_SUB=$sub timeout 200 python tools/gpu_latency.py > gpurun_out/s8_lat_sub$sub.json 2> gpurun_out/s8_lat_sub$sub.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/s8_lat_sub$sub.json").read().split("\nsolo kernel")[0])
for k in ("mag_handoff", "iq_handoff", "mag_handoff_with_cuda_events"):
    v = d.get(k, {})
    print("sub=$sub", k, {q: v.get(q) for q in ("median_us", "p99_us", "min_us", "device_timeline_of_last_call_us")})
PY
done
timeout 100 python tools/gpu_modeac_timing.py 2>&1 | tail -1 | cut -c1-260 | tee gpurun_out/s8_modeac.txt
timeout 100 python tools/gpu_timing.py 2>&1 | tail -2 | cut -c1-220 | tee gpurun_out/s8_timing_cfg2.txt
timeout 100 python tools/gpu_scan_probe.py 2>&1 | tail -4 | tee gpurun_out/s8_scan_probe.txt
