#!/bin/bash
# timing of the level-0 3x3 convolution under the kernel's tuning / debug overrides
for cfg in "$@"; do
  echo "== $cfg"
  env $cfg MFLOW_CONV_TIMING_ONLY=1 timeout 120 python tools/test_conv_tc.py 2>&1 | grep -E "tcgen05|variant \+res|FAILED|rror"
done
