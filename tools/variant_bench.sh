#!/bin/bash
# run the short bench for the default build and every build_variants/*.so, for the three headline workloads
cd "$(dirname "$0")/.."
run() { python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline "$@" | python -c "import json,sys; d=json.load(sys.stdin); print('%.4g cell-steps/s  %.3f ms/step  frac %.3f' % (d['value'], d['ms_per_step'], d['roofline']['frac']))"; }
all3() {
  echo -n "  coupled n=2     : "; run
  echo -n "  coupled general : "; run --general-vg
  echo -n "  richards 100    : "; run --model richards --nlayer 100 --ncol 655360
}
echo "default:"; all3
for so in build_variants/*.so; do [ -f "$so" ] || continue; echo "$so:"; export LH_SOIL_LIBRARY=$PWD/$so; all3; done
