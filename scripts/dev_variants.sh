#!/bin/bash
# Development aid (GPU box): what the driver runs at round end, then the ncu capture set
O=gpurun_out/r2y; mkdir -p $O
( time python -c "import __graft_entry__ as g; g.smoke()" ) > $O/smoke.txt 2>&1
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > $O/pytest.txt
( time python bench.py --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err ) 2> $O/bench.time
( time python bench.py --impl reference --steps 5 --warmup 2 > $O/bench_ref.json 2> $O/bench_ref.err ) 2> $O/bench_ref.time
tail -3 $O/smoke.txt; cat $O/pytest.txt; tail -3 $O/bench.err; cat $O/bench.time
bash scripts/gpu_profile_all.sh r2b > $O/profile_all.log 2>&1
ls gpurun_out/prof_r2b | wc -l
