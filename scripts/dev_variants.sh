#!/bin/bash
# Development aid (GPU box, 8 GPUs): copy ceiling at 1/2/4/8 GPUs and the headline of BASELINE configs[3]
O=gpurun_out/r2u; mkdir -p $O
nvidia-smi -L > $O/gpus.txt; nproc >> $O/gpus.txt; free -g >> $O/gpus.txt
python scripts/h2d_bench.py --gpus 1,2,4,8 > $O/h2d.txt 2>&1
cat $O/h2d.txt
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 5 --warmup 3 --headline-only > $O/bench_8gpu.json 2> $O/bench_8gpu.err ) 2> $O/time.txt
tail -5 $O/bench_8gpu.err; cat $O/time.txt
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2u/bench_8gpu.json').read().strip().splitlines()[-1])
print('value',d['value'],d['ms_per_step'],'e2e',d['e2e'], 'gather', d['gather_ms_per_step'])
print(d['config'])
print(d['parity'], d['parity_failed'])
PY
