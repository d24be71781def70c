#!/bin/bash
# chunk length vs launch strategy at multi-GPU shard sizes (131072 = 1/8, 262144 = 1/4 of 2^20 columns)
cd "$(dirname "$0")/.."
run() { python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline "$@" | python -c "import json,sys; d=json.load(sys.stdin); print('%.4g  %.4f ms/step  launches %d' % (d['value'], d['ms_per_step'], d['gpu_launches']))"; }
for n in 16384 65536 131072 262144 1048576; do echo -n "ncol $n auto: "; run --ncol $n; done
