#!/bin/bash
# Per-kernel launch lists (both modes) and ncu --set full captures of the hot kernels.
# Run on the GPU box from the repo root:  bash scripts/gpu_profile_all.sh [tag]
# A number printed under ncu is never a bench value: these files only give the
# kernels' SHARES of a step and their pipe/stall profiles.
set -u
TAG=${1:-r1}
OUT=gpurun_out/prof_$TAG
mkdir -p $OUT
export PEAQ_PROFILE_PAIRS=${PEAQ_PROFILE_PAIRS:-592}
for adv in 0 1; do
  PEAQ_PROFILE_ADVANCED=$adv ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file $OUT/launches_adv$adv.csv python scripts/profile_workload.py > $OUT/launches_adv$adv.log 2>&1
done
cap() {  # kernel-regex advanced skip name
  PEAQ_PROFILE_ADVANCED=$2 ncu --set full --clock-control none --import-source on -k regex:$1 --launch-skip $3 -c 1 \
    -f -o $OUT/$4 python scripts/profile_workload.py > $OUT/$4.log 2>&1
  ncu -i $OUT/$4.ncu-rep --page raw --csv > $OUT/$4.raw.csv 2>/dev/null
  ncu -i $OUT/$4.ncu-rep --page source --csv --print-source cuda,sass > $OUT/$4.source.csv 2>/dev/null
  ncu -i $OUT/$4.ncu-rep --page details > $OUT/$4.details.txt 2>/dev/null
  rm -f $OUT/$4.ncu-rep
}
cap fft_frames_kernel 0 1 fft_frames
cap scan_basic_kernel 0 1 scan_basic
cap fb_bank_rec_kernel 1 1 fb_bank_rec
cap fb_spread_kernel 1 1 fb_spread
cap fb_scan_kernel 1 1 fb_scan
cap fb_hp_kernel 1 1 fb_hp
ls -la $OUT
