#!/bin/bash
# Development aid (GPU box, 2 GPUs): the distributed bench with the 8-GPU batch size, copy ceiling
O=gpurun_out/r2t; mkdir -p $O
nvidia-smi -L > $O/gpus.txt; nproc >> $O/gpus.txt; free -g >> $O/gpus.txt
python scripts/h2d_bench.py --gpus 1,2 > $O/h2d.txt 2>&1
cat $O/h2d.txt
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 3 --pairs-per-gpu 8192 > $O/bench_2gpu.json 2> $O/bench_2gpu.err ) 2> $O/time.txt
tail -5 $O/bench_2gpu.err; cat $O/time.txt
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2t/bench_2gpu.json').read().strip().splitlines()[-1])
print('value',d['value'],d['ms_per_step'],'e2e',d['e2e'], 'gather', d['gather_ms_per_step'])
print('adv',d['modes']['advanced']['value'],d['modes']['advanced']['e2e']['value'])
for m in ('basic','advanced'): print('long',m,d['long_items'][m]['value'],d['long_items'][m]['e2e']['value'])
print(d['parity'], d['parity_failed'])
PY
