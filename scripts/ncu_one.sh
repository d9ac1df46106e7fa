#!/bin/bash
# ncu --set full capture of one kernel:  bash scripts/ncu_one.sh <kernel-regex> <advanced 0|1> <out-name> [pairs]
# (environment, e.g. PEAQ_B200_FUSED, is passed through to the workload)
OUT=gpurun_out/$3
mkdir -p $(dirname $OUT)
PEAQ_PROFILE_PAIRS=${4:-592} PEAQ_PROFILE_ADVANCED=$2 ncu --set full --clock-control none --import-source on -k regex:$1 --launch-skip 1 -c 1 \
  -f -o $OUT python scripts/profile_workload.py > $OUT.log 2>&1
ncu -i $OUT.ncu-rep --page raw --csv > $OUT.raw.csv 2>/dev/null
ncu -i $OUT.ncu-rep --page source --csv --print-source cuda,sass > $OUT.source.csv 2>/dev/null
ncu -i $OUT.ncu-rep --page details > $OUT.details.txt 2>/dev/null
rm -f $OUT.ncu-rep
