#!/bin/bash
# Development aid (GPU box): compute-sanitizer passes
O=gpurun_out/r2san; mkdir -p $O
S=/usr/local/cuda/bin/compute-sanitizer
$S --tool memcheck --error-exitcode 9 python scripts/sanitize_workload.py short session fused > $O/memcheck_short.log 2>&1; echo "memcheck short rc $?" | tee $O/summary.txt
$S --tool memcheck --error-exitcode 9 python scripts/sanitize_workload.py segments > $O/memcheck_segments.log 2>&1; echo "memcheck segments rc $?" | tee -a $O/summary.txt
$S --tool racecheck --error-exitcode 9 python scripts/sanitize_workload.py short session > $O/racecheck_short.log 2>&1; echo "racecheck short rc $?" | tee -a $O/summary.txt
$S --tool synccheck --error-exitcode 9 python scripts/sanitize_workload.py short > $O/synccheck_short.log 2>&1; echo "synccheck short rc $?" | tee -a $O/summary.txt
PEAQ_B200_HP_PARALLEL=0 $S --tool memcheck --error-exitcode 9 python scripts/sanitize_workload.py short > $O/memcheck_hpseq.log 2>&1; echo "memcheck hp sequential rc $?" | tee -a $O/summary.txt
for f in $O/*.log; do echo "== $f"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|odg|adv" $f | tail -12; done
