#!/bin/bash
# round-2 baseline: GPU parity tests, the variant table of the round-1 build, and a sustained (>= 3 s) timing probe
cd "$(dirname "$0")/.."
o=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $o/r02a_smi.txt
timeout 900 python -m pytest tests -m gpu -x -q > $o/r02a_pytest_gpu.log 2>&1; echo "pytest rc $?" >> $o/r02a_pytest_gpu.log
tail -3 $o/r02a_pytest_gpu.log
timeout 900 python tools/variants.py > $o/r02a_variants.json 2> $o/r02a_variants.err; tail -20 $o/r02a_variants.err
# sustained: 2000 steps (~3.3 s) vs 20 steps, clocks sampled by bench.py
timeout 300 python bench.py --steps 2000 --warmup 3 --no-e2e --no-cpu-baseline > $o/r02a_bench_2000.json 2> $o/r02a_bench_2000.err; cat $o/r02a_bench_2000.json
timeout 300 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline > $o/r02a_bench_20.json 2> $o/r02a_bench_20.err; cat $o/r02a_bench_20.json
