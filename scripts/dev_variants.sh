#!/bin/bash
# Development aid (GPU box)
O=gpurun_out/r2s; mkdir -p $O
(PEAQ_PROFILE_ADVANCED=1 python scripts/e2e_diag.py
PEAQ_PROFILE_ADVANCED=1 PEAQ_PROFILE_PAIRS=32 PEAQ_PROFILE_SECONDS=600 python scripts/e2e_diag.py
PEAQ_PROFILE_ADVANCED=0 PEAQ_PROFILE_PAIRS=32 PEAQ_PROFILE_SECONDS=600 python scripts/e2e_diag.py
PEAQ_PROFILE_ADVANCED=0 python scripts/e2e_diag.py) > $O/log.txt 2>&1
grep host $O/log.txt
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
