#!/bin/bash
# Development aid (GPU box): FMA contraction on / off
O=gpurun_out/r2ad; mkdir -p $O
L=$PWD/gstpeaq_b200
for v in "" _FMAD; do
  PEAQ_B200_LIBRARY=$L/libpeaq_b200$v.so python scripts/time_modes.py both 2>&1 | grep libpeaq
done > $O/log.txt
cat $O/log.txt
PEAQ_B200_LIBRARY=$L/libpeaq_b200_FMAD.so python -m pytest tests -m gpu -q 2>&1 | tail -15
