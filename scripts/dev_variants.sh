#!/bin/bash
# Development aid (GPU box)
O=gpurun_out/r2ac; mkdir -p $O
( time python -m pytest tests -m gpu -x -q -k "ten_minute" 2>&1 | tail -15 ) > $O/pytest.txt 2>&1
cat $O/pytest.txt
export PEAQ_PROFILE_PAIRS=592
for spec in "3 fb_hp_par_zero" "5 fb_hp_par_out"; do
  set -- $spec
  PEAQ_PROFILE_ADVANCED=1 ncu --set full --clock-control none --import-source on -k regex:fb_hp_par_block_kernel --launch-skip $1 -c 1 -f -o $O/$2 python scripts/profile_workload.py > $O/$2.log 2>&1
  ncu -i $O/$2.ncu-rep --page raw --csv > $O/$2.raw.csv 2>/dev/null
  ncu -i $O/$2.ncu-rep --page source --csv --print-source cuda,sass > $O/$2.source.csv 2>/dev/null
  ncu -i $O/$2.ncu-rep --page details > $O/$2.details.txt 2>/dev/null
  rm -f $O/$2.ncu-rep
done
ls -la $O
