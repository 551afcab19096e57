#!/bin/bash
# Round 2: the chip partitioned between the scan (a grid smaller than the chip, back to back) and stage B + finalizer of the step
# before (what is left), instead of alternating on all SMs: fixed grids (B200_SCAN_SMS) against the tuned one (default) and none.
mkdir -p gpurun_out
run() { # label, env...
  local label=$1; shift
  echo -n "$label: "
  env "$@" timeout 300 python bench.py --steps 4 --warmup 3 --launches-per-step 48 --no-cpu --no-extra "${EXTRA[@]}" 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), round(d['ms_per_step']/48, 4), 'ms/launch  scan alone', round(d['roofline']['kernel_ms_per_launch'], 4), ' e2e', round(d['e2e']['value']))"
}
EXTRA=()
run "whole chip (B200_SCAN_PART=0)" B200_SCAN_PART=0
run "tuned (default)              " X=1
run "fixed 126                    " B200_SCAN_SMS=126
run "tuned (default)              " X=1
run "whole chip                   " B200_SCAN_PART=0
EXTRA=(--workload config5_dense)
run "dense whole chip             " B200_SCAN_PART=0
run "dense tuned                  " X=1
run "dense fixed 126              " B200_SCAN_SMS=126
