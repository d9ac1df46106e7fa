#!/bin/bash
# Development aid (GPU box): bit-identity / deltas of the current build against the round-1 build,
# resident-step timings of the variants, then the GPU tests.
L=$PWD/gstpeaq_b200
python scripts/compare_builds.py $L/libpeaq_b200_r1.so $L/libpeaq_b200_libm.so
PEAQ_B200_FUSED=1 python scripts/compare_builds.py $L/libpeaq_b200_r1.so $L/libpeaq_b200_libm.so
python scripts/compare_builds.py $L/libpeaq_b200_r1.so $L/libpeaq_b200.so
PEAQ_B200_LIBRARY=$L/libpeaq_b200_r1.so python scripts/time_modes.py both
PEAQ_B200_FUSED=0 PEAQ_B200_LIBRARY=$L/libpeaq_b200_libm.so python scripts/time_modes.py both
PEAQ_B200_FUSED=1 PEAQ_B200_LIBRARY=$L/libpeaq_b200_libm.so python scripts/time_modes.py basic
PEAQ_B200_FUSED=0 python scripts/time_modes.py both
PEAQ_B200_FUSED=1 python scripts/time_modes.py basic
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
