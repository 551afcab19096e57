#!/bin/bash
# A/B on the GPU: persistent scan grid smaller than the chip, so that stage B + finalize of step n run on the SMs the scan of
# step n+1 leaves free instead of alternating with it.  One bench line per setting.
mkdir -p gpurun_out
for n in 148 146 144 140 132; do
  B200_SCAN_SMS=$n timeout 120 python bench.py --no-cpu --no-e2e --no-extra --steps 4 --warmup 1 --launches-per-step 48 2>/dev/null \
    | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('scan_sms', $n, 'ms_per_launch', round(d['ms_per_step']/48,4), 'value', round(d['value']), 'scan_ms', round(d['roofline']['kernel_ms_per_launch'],4))" \
    | tee -a gpurun_out/scan_sms.txt
done
