#!/bin/bash
# Development aid (GPU box)
O=gpurun_out/r2n; mkdir -p $O
python -m pytest tests -m gpu -x -q -k "segment" 2>&1 | tail -40 > $O/pytest_seg.txt
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > $O/pytest.txt
ll() { # name pairs seconds advanced
  PEAQ_PROFILE_PAIRS=$2 PEAQ_PROFILE_SECONDS=$3 PEAQ_PROFILE_ADVANCED=$4 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv \
    --log-file $O/$1.csv python scripts/profile_workload.py > $O/$1.log 2>&1
  python scripts/launch_summary.py $O/$1.csv > $O/$1.txt; cat $O/$1.log $O/$1.txt
}
cat $O/pytest_seg.txt $O/pytest.txt
ll long_adv 32 600 1
ll long_basic 32 600 0
