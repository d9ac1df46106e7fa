#!/bin/bash
# Development aid (GPU box)
O=gpurun_out/r2o; mkdir -p $O
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > $O/pytest.txt
python scripts/time_modes.py both 32 600 > $O/log.txt 2>&1
python scripts/time_modes.py advanced >> $O/log.txt 2>&1
PEAQ_B200_HP_PARALLEL=1 python scripts/time_modes.py advanced >> $O/log.txt 2>&1
cat $O/pytest.txt $O/log.txt
ll() { # name pairs seconds advanced
  PEAQ_PROFILE_PAIRS=$2 PEAQ_PROFILE_SECONDS=$3 PEAQ_PROFILE_ADVANCED=$4 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv \
    --log-file $O/$1.csv python scripts/profile_workload.py > $O/$1.log 2>&1
  python scripts/launch_summary.py $O/$1.csv > $O/$1.txt; cat $O/$1.log $O/$1.txt
}
ll long_adv 32 600 1
