#!/bin/bash
# Round 2, third measurement session: A/B of the scan kernel against the build before the shared-tile change (same box, alternating),
# Mode A/C with the Mode S scan's magnitudes.
mkdir -p gpurun_out
OLD=${1:?path of the library build to compare with}
for i in 1 2; do
  echo "--- new"; timeout 100 python tools/gpu_scan_probe.py 2>&1 | tail -4
  echo "--- old"; B200_DEMOD_LIB=$OLD timeout 100 python tools/gpu_scan_probe.py 2>&1 | tail -4
done | tee gpurun_out/s9_scan_ab.txt
echo "--- modeac new"; timeout 100 python tools/gpu_modeac_timing.py 2>&1 | tail -1 | cut -c1-260 | tee gpurun_out/s9_modeac_new.txt
echo "--- modeac old"; B200_DEMOD_LIB=$OLD timeout 100 python tools/gpu_modeac_timing.py 2>&1 | tail -1 | cut -c1-260 | tee gpurun_out/s9_modeac_old.txt
echo "--- modeac new"; timeout 100 python tools/gpu_modeac_timing.py 2>&1 | tail -1 | cut -c1-260 | tee -a gpurun_out/s9_modeac_new.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:modeac -c 12 --csv --log-file gpurun_out/s9_modeac_launches.csv python tools/gpu_modeac_timing.py > /dev/null 2>&1
grep -i modeac gpurun_out/s9_modeac_launches.csv | cut -d, -f5,13- | tail -8
( timeout 600 python -m pytest tests -m gpu -x -q -k "modeac or ac or edges or fullsize" 2>&1 | tail -3 ) | tee gpurun_out/s9_pytest.log
timeout 200 python tools/gpu_latency.py > gpurun_out/s9_lat.json 2> gpurun_out/s9_lat.err
python - <<PY
import json
d=json.loads(open("gpurun_out/s9_lat.json").read().split("\nsolo kernel")[0])
for k in ("mag_handoff", "iq_handoff", "mag_handoff_with_cuda_events"):
    v = d.get(k, {})
    print(k, {q: v.get(q) for q in ("median_us", "p99_us", "min_us", "device_timeline_of_last_call_us")})
PY
