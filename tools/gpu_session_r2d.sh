#!/bin/bash
# Round 2: A/B of the working tree against a variant build of the library (default: the build before the shared-tile change), same box, alternating.
mkdir -p gpurun_out
V=${1:?path of the variant build of libb200demod.so (readsb_b200.build.build_demod(defines=..., out=...))}
for i in 1 2; do
  echo "--- head";  timeout 100 python tools/gpu_scan_probe.py 2>&1 | tail -4
  echo "--- variant"; B200_DEMOD_LIB=$V timeout 100 python tools/gpu_scan_probe.py 2>&1 | tail -4
done | tee gpurun_out/s10_scan_ab.txt
for i in 1 2; do
  echo "--- head";  timeout 200 python bench.py --steps 3 --warmup 3 --launches-per-step 48 --no-cpu --no-extra 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline'])"
  echo "--- variant"; B200_DEMOD_LIB=$V timeout 200 python bench.py --steps 3 --warmup 3 --launches-per-step 48 --no-cpu --no-extra 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline'])"
done | tee gpurun_out/s10_bench_ab.txt
echo "--- cfg5 head";  timeout 100 python tools/gpu_timing.py cfg5 2>&1 | tail -1 | cut -c1-200
echo "--- cfg5 variant"; B200_DEMOD_LIB=$V timeout 100 python tools/gpu_timing.py cfg5 2>&1 | tail -1 | cut -c1-200
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_edges.py -m gpu -x -q 2>&1 | tail -2
