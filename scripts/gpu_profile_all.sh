#!/bin/bash
# Per-kernel launch lists (both modes, and bench.py itself) and ncu --set full captures of the hot
# kernels.  Run on the GPU box from the repo root:  bash scripts/gpu_profile_all.sh [tag]
# A number printed under ncu is never a bench value: these files give the kernels' SHARES of a
# step, their DRAM traffic and their pipe / stall profiles.
set -u
TAG=${1:-r2}
OUT=gpurun_out/prof_$TAG
mkdir -p $OUT
export PEAQ_PROFILE_PAIRS=${PEAQ_PROFILE_PAIRS:-592}
# one filter-bank chunk per pass, so that a captured launch covers all frames of the workload
export PEAQ_B200_FB_BUDGET_MB=${PEAQ_B200_FB_BUDGET_MB:-65536}
cap() {  # kernel-regex advanced skip name
  PEAQ_PROFILE_ADVANCED=$2 ncu --set full --clock-control none --import-source on -k regex:$1 --launch-skip $3 -c 1 \
    -f -o $OUT/$4 python scripts/profile_workload.py > $OUT/$4.log 2>&1
  ncu -i $OUT/$4.ncu-rep --page raw --csv > $OUT/$4.raw.csv 2>/dev/null
  ncu -i $OUT/$4.ncu-rep --page source --csv --print-source cuda,sass > $OUT/$4.source.csv 2>/dev/null
  ncu -i $OUT/$4.ncu-rep --page details > $OUT/$4.details.txt 2>/dev/null
  rm -f $OUT/$4.ncu-rep
}
# most important first (a call may be cut short by the GPU budget)
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches_bench.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --headline-only > $OUT/bench_under_ncu.json 2> $OUT/bench_under_ncu.err
cap fft_frames_kernel 0 1 fft_frames
cap fb_bank_rec_kernel 1 1 fb_bank_rec
cap scan_basic_kernel 0 1 scan_basic
cap fft_frames_kernel 1 1 fft_frames_adv
cap fb_spread_kernel 1 1 fb_spread
cap fb_scan_kernel 1 1 fb_scan
# the block passes of the DC-reject scan: zero-state, from-s0 and output pass per run (skip the first run)
cap fb_hp_par_block_kernel 1 3 fb_hp_par_zero
cap fb_hp_par_block_kernel 1 5 fb_hp_par_out
for adv in 0 1; do
  PEAQ_PROFILE_ADVANCED=$adv ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file $OUT/launches_adv$adv.csv python scripts/profile_workload.py > $OUT/launches_adv$adv.log 2>&1
done
if [ "${PEAQ_PROFILE_FUSED:-0}" = 1 ]; then PEAQ_B200_FUSED=1 cap peaq_fused_basic_kernel 0 1 fused_basic; fi
ls -la $OUT
