#!/bin/bash
# GPU session 2 of round 2: parity on the hardware after the filter / scan changes, A/B timing of the library variants
# (readsb_b200/libb200demod*.so, see tools/gpu_variants.sh), a full ncu capture of the new scan kernel.
mkdir -p gpurun_out
(time timeout 600 python -m pytest tests -m gpu -x -q) > gpurun_out/s2_pytest_gpu.log 2>&1; tail -3 gpurun_out/s2_pytest_gpu.log
for lib in readsb_b200/libb200demod*.so; do
    name=$(basename $lib .so)
    B200_DEMOD_LIB=$PWD/$lib timeout 200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edges.py -m gpu -x -q -k "replay or many_streams or device_resident or ragged or async_pipeline_matches" 2>&1 | tail -1
    for wl in cfg2 cfg5; do
        B200_DEMOD_LIB=$PWD/$lib timeout 100 python tools/gpu_timing.py $wl 2>&1 | tail -2 | sed -e "s/^/$name /" | cut -c1-200 | tee -a gpurun_out/s2_variants.txt
    done
done
timeout 200 ncu --set full --clock-control none --import-source on -k regex:^scan_kernel -s 3 -c 1 -f -o gpurun_out/s2_scan_kernel python tools/gpu_timing.py > gpurun_out/s2_ncu_scan.log 2>&1; tail -1 gpurun_out/s2_ncu_scan.log
timeout 200 python bench.py --no-cpu > gpurun_out/s2_bench_n1.json 2> gpurun_out/s2_bench_n1.err; cut -c1-300 gpurun_out/s2_bench_n1.json
