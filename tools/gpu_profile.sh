#!/bin/bash
# tools/gpu_profile.sh <tag>: plain bench, ncu launch list, and ncu --set full captures of the three stage
# kernels for the coupled n=2 / coupled general-n / Richards workloads.  Run under gpurun.
tag=${1:-x}
cd "$(dirname "$0")/.."
o=gpurun_out
python bench.py > $o/bench_$tag.json 2> $o/bench_$tag.err || exit 1
cat $o/bench_$tag.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $o/launches_$tag.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $o/ncu_launches_$tag.log 2>&1
B="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline"
N="ncu --set full --clock-control none --import-source on -k regex:lh_soil_stage -s 9 -c 3 -f"
$N -o $o/prof_${tag}_coupled $B > $o/ncu_${tag}_coupled.log 2>&1
$N -o $o/prof_${tag}_general $B --general-vg > $o/ncu_${tag}_general.log 2>&1
$N -o $o/prof_${tag}_richards $B --model richards --nlayer 100 --ncol 655360 > $o/ncu_${tag}_richards.log 2>&1
ls -la $o | tail -8
