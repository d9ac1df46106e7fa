#!/bin/bash
# Development aid (GPU box): what the driver runs at round end -- smoke, the GPU tests, the default
# bench line and the reference arm -- into gpurun_out/<tag>/.   bash scripts/round_end_check.sh [tag]
O=gpurun_out/${1:-roundend}; mkdir -p $O
( time python -c "import __graft_entry__ as g; g.smoke()" ) > $O/smoke.txt 2>&1
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > $O/pytest.txt
( time python bench.py --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err ) 2> $O/bench.time
if [ "${PEAQ_CHECK_REFERENCE_ARM:-1}" = 1 ]; then ( time python bench.py --impl reference --steps 5 --warmup 2 > $O/bench_ref.json 2> $O/bench_ref.err ) 2> $O/bench_ref.time; fi
tail -3 $O/smoke.txt; cat $O/pytest.txt; tail -3 $O/bench.err; cat $O/bench.time
