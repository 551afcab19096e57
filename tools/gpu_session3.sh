#!/bin/bash
# A/B of the scan-kernel toggles (one variant library each) + a trial of the new bench legs
mkdir -p gpurun_out
for lib in readsb_b200/libb200demod*.so; do
    name=$(basename $lib .so)
    B200_DEMOD_LIB=$PWD/$lib timeout 100 python tools/gpu_timing.py cfg2 2>&1 | tail -2 | sed -e "s/^/$name /" | cut -c1-175 | tee -a gpurun_out/s3_variants.txt
done
B200_DEMOD_LIB=$PWD/readsb_b200/libb200demod.so timeout 100 python tools/gpu_timing.py cfg5 2>&1 | tail -1 | cut -c1-175 | tee -a gpurun_out/s3_variants.txt
timeout 300 python bench.py --steps 2 --warmup 1 --launches-per-step 8 > gpurun_out/s3_bench_trial.json 2> gpurun_out/s3_bench_trial.err; tail -3 gpurun_out/s3_bench_trial.err; python -c "
import json; d=json.loads(open('gpurun_out/s3_bench_trial.json').readline()); print({k:d[k] for k in ('value','ms_per_step','timed_region_s','parity_checked') if k in d}); print(json.dumps(d.get('extra'))[:1500]); print(d.get('burst')); print(d.get('e2e')); print(d.get('roofline'))"
