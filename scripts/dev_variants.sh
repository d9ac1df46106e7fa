#!/bin/bash
# Development aid (GPU box): occupancy variants of the advanced-mode scan kernels
O=gpurun_out/r2x; mkdir -p $O
L=$PWD/gstpeaq_b200
for v in "" _S7 _S8 _F5 _F6; do
  PEAQ_B200_LIBRARY=$L/libpeaq_b200$v.so python scripts/time_modes.py advanced 2048 2>&1 | tail -1
done > $O/log.txt
cat $O/log.txt
