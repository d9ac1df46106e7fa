#!/bin/bash
# Development aid (GPU box): timings of build variants (basic + advanced resident steps)
L=$PWD/gstpeaq_b200
python scripts/time_modes.py both
for v in ROLLED K2OCC3 ROT; do
  PEAQ_B200_LIBRARY=$L/libpeaq_b200_$v.so python scripts/time_modes.py basic
done
PEAQ_B200_LIBRARY=$L/libpeaq_b200_ROLLED.so python scripts/time_modes.py advanced
PEAQ_B200_FUSED=1 python scripts/time_modes.py basic
PEAQ_B200_FUSED=1 PEAQ_B200_LIBRARY=$L/libpeaq_b200_ROLLED.so python scripts/time_modes.py basic
python scripts/compare_builds.py $L/libpeaq_b200_r1.so $L/libpeaq_b200_libm.so | tail -1
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
