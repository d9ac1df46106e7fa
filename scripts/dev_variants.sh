#!/bin/bash
# Development aid (GPU box): timings of build variants (basic + advanced resident steps)
L=$PWD/gstpeaq_b200
python scripts/time_modes.py both
PEAQ_B200_LIBRARY=$L/libpeaq_b200_ROLLED.so python scripts/time_modes.py both
echo "--- long items 32 x 600 s"
python scripts/time_modes.py basic 32 600
python scripts/compare_builds.py $L/libpeaq_b200_r1.so $L/libpeaq_b200.so
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
