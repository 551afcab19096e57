#!/bin/bash
# A/B timing of experimental builds of the library on one GPU (see readsb_b200/build.py build_demod(defines=..., out=...)):
# for every readsb_b200/libb200demod*.so: a parity subset, then the per-kernel CUDA-event timings of a bench-shaped step.
mkdir -p gpurun_out
for lib in readsb_b200/libb200demod*.so; do
    name=$(basename $lib .so)
    echo "== $name"
    B200_DEMOD_LIB=$PWD/$lib timeout 200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edges.py -m gpu -x -q -k "replay or many_streams or device_resident or ragged or fuzz or async_pipeline_matches" 2>&1 | tail -1
    B200_DEMOD_LIB=$PWD/$lib timeout 100 python tools/gpu_timing.py 2>&1 | tail -3 | sed -e "s/^/$name /" | cut -c1-230
done
