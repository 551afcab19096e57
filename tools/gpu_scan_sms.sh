#!/bin/bash
# A/B on the GPU: persistent scan grid smaller than the chip, so that stage B + finalize of step n run on the SMs the scan of
# step n+1 leaves free instead of alternating with it.  One bench line per setting.
mkdir -p gpurun_out
for n in 148 144 140 136 132 124; do
  B200_SCAN_SMS=$n timeout 120 python bench.py --no-cpu --no-e2e --steps 30 --warmup 6 2>/dev/null \
    | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('scan_sms', $n, 'ms_per_step', round(d['ms_per_step'],4), 'value', round(d['value']), 'scan_ms', round(d['roofline']['kernel_ms_per_launch'],4))" \
    | tee -a gpurun_out/scan_sms.txt
done
