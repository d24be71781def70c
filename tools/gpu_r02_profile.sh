#!/bin/bash
# round 2 ncu evidence: launch list of the default bench command, --set full captures of the stage kernels for the coupled n = 2,
# coupled general-n, coupled ice, Richards workloads; per-variant DRAM traffic -> profiles/dram_traffic.json format.
tag=${1:-r02p}
cd "$(dirname "$0")/.."
o=gpurun_out
B="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-variants --min-seconds 0"
$B > $o/${tag}_plain.json 2> $o/${tag}_plain.err || { tail -5 $o/${tag}_plain.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $o/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-variants --min-seconds 0 > $o/${tag}_ncu_launches.log 2>&1
N="ncu --set full --clock-control none --import-source on -k regex:lh_soil_stage -s 9 -c 3 -f"
sum=$o/${tag}_ncu_full_summary.txt
: > $sum
for m in coupled general ice richards; do
  case $m in coupled) extra=""; cells=67108864;; general) extra="--general-vg"; cells=67108864;; ice) extra="--ice"; cells=67108864;;
             richards) extra="--model richards --nlayer 100 --ncol 655360"; cells=65536000;; esac
  $N -o /tmp/prof_${tag}_$m $B $extra > $o/${tag}_ncu_$m.log 2>&1
  echo "=== $m (ncu --set full --clock-control none; bench.py --steps 1 --warmup 3 $extra; launches 10-12 = stages 1, 2, 3 of the first timed step)" >> $sum
  python tools/ncu_stalls.py /tmp/prof_${tag}_$m.ncu-rep >> $sum 2>&1
  python tools/ncu_mix.py /tmp/prof_${tag}_$m.ncu-rep $cells 2>/dev/null | grep -A30 "warp-instructions per cell" >> $sum
done
ls -la /tmp/prof_${tag}_*.ncu-rep
cp /tmp/prof_${tag}_coupled.ncu-rep $o/ 2>/dev/null
du -sh $o
