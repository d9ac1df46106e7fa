#!/bin/bash
# Development aid (GPU box)
O=gpurun_out/r2ab; mkdir -p $O
for i in 1 2 3; do PEAQ_B200_DEBUG=1 PEAQ_PROFILE_ADVANCED=1 python scripts/e2e_diag.py 2>&1 | grep -E "host|sub-batches" | tail -3; done > $O/log.txt 2>&1
cat $O/log.txt
