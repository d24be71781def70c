#!/bin/bash
# run the short bench for the default build and every build_variants/*.so
cd "$(dirname "$0")/.."
run() { python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline "$@" | python -c "import json,sys; d=json.load(sys.stdin); print('%.4g cell-steps/s  %.3f ms/step  frac %.3f' % (d['value'], d['ms_per_step'], d['roofline']['frac']))"; }
echo "default:"; run "$@"
for so in build_variants/*.so; do echo "$so:"; LH_SOIL_LIBRARY=$PWD/$so run "$@"; done
