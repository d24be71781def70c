#!/bin/bash
# per-stage vs persistent launches at several problem sizes (1 GPU)
cd "$(dirname "$0")/.."
run() { python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline "$@" | python -c "import json,sys; d=json.load(sys.stdin); print('%.4g cell-steps/s  %.4f ms/step  frac %.3f  launches %d' % (d['value'], d['ms_per_step'], d['roofline']['frac'], d['gpu_launches']))"; }
for args in "--ncol 1048576" "--ncol 131072" "--ncol 16384" "--ncol 1024" "--model richards --nlayer 100 --ncol 1024" "--model richards --nlayer 150 --ncol 1" "--general-vg --ncol 131072" "--model richards --nlayer 100 --ncol 81920"; do
  for l in auto; do echo -n "$args --launch $l: "; run $args --launch $l; done
done
