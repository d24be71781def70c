#!/bin/bash
# tools/gpu_profile.sh <tag>: plain bench, ncu launch list, and ncu --set full captures of the three stage
# kernels for the coupled n=2 / coupled general-n / Richards workloads.  Run under gpurun.  The .ncu-rep files
# are summarised on the box (tools/ncu_stalls.py, tools/ncu_mix.py) and only the coupled one is kept: gpurun_out/
# is limited to 64 MiB.
tag=${1:-x}
cd "$(dirname "$0")/.."
o=gpurun_out
python bench.py > $o/bench_$tag.json 2> $o/bench_$tag.err || exit 1
cat $o/bench_$tag.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $o/launches_$tag.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $o/ncu_launches_$tag.log 2>&1
B="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline"
N="ncu --set full --clock-control none -k regex:lh_soil_stage -s 9 -c 3 -f"
sum=$o/ncu_summary_$tag.txt
: > $sum
for m in coupled general richards; do
  case $m in coupled) extra=""; cells=67108864;; general) extra="--general-vg"; cells=67108864;; richards) extra="--model richards --nlayer 100 --ncol 655360"; cells=65536000;; esac
  $N -o /tmp/prof_${tag}_$m $B $extra > $o/ncu_${tag}_$m.log 2>&1
  echo "=== $m (ncu --set full --clock-control none; bench.py --steps 1 --warmup 3 $extra; launches 10-12)" >> $sum
  python tools/ncu_stalls.py /tmp/prof_${tag}_$m.ncu-rep >> $sum 2>&1
  python tools/ncu_mix.py /tmp/prof_${tag}_$m.ncu-rep $cells 2>/dev/null | grep -A30 "warp-instructions per cell" >> $sum
done
ls -la /tmp/prof_${tag}_*.ncu-rep
cp /tmp/prof_${tag}_coupled.ncu-rep $o/ 2>/dev/null
du -sh $o
