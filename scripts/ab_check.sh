#!/bin/bash
# Development aid (GPU box): the in-tree build against gstpeaq_b200/libpeaq_b200_prev.so (the previous
# commit's kernels, built by hand) -- result bytes, the GPU tests, resident-step timings of both.
O=gpurun_out/${1:-ab}; mkdir -p $O
PREV=gstpeaq_b200/libpeaq_b200_prev.so   # e.g. git worktree add /tmp/prev <commit>; make -C /tmp/prev/gstpeaq_b200/csrc lib LIBNAME=libpeaq_b200_prev.so; cp it here
( PEAQ_B200_FB_SMEM_COEF=1 timeout 300 python scripts/compare_builds.py $PREV ) > $O/compare_smemcoef.txt 2>&1
( timeout 300 python scripts/compare_builds.py $PREV ) > $O/compare_default.txt 2>&1
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > $O/pytest.txt
( PEAQ_B200_LIBRARY=$PWD/$PREV timeout 300 python scripts/time_modes.py both ) > $O/time_prev.txt 2>&1
( timeout 300 python scripts/time_modes.py both ) > $O/time_new.txt 2>&1
( PEAQ_B200_FB_SMEM_COEF=1 timeout 300 python scripts/time_modes.py advanced ) > $O/time_new_smemcoef.txt 2>&1
tail -3 $O/compare_smemcoef.txt; tail -9 $O/compare_default.txt; cat $O/pytest.txt; grep -h "frames/s" $O/time_*.txt
