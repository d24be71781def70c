#!/bin/bash
# tools/gpu_scaling.sh <tag> <ngpus>: bench.py at N = 1, 2, 4, 8 (up to <ngpus>) back to back, like the driver's
# scaling run; one JSON line per N into gpurun_out/scale_<tag>.jsonl.  Run under `gpurun --gpus <ngpus>`.
tag=${1:-x}; nmax=${2:-8}
cd "$(dirname "$0")/.."
out=gpurun_out/scale_$tag.jsonl
: > $out
for n in 1 2 4 8; do
  [ $n -gt $nmax ] && break
  if [ $n -eq 1 ]; then
    python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu-baseline >> $out 2>> gpurun_out/scale_$tag.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) \
      bench.py --gpus $n --steps 20 --warmup 3 >> $out 2>> gpurun_out/scale_$tag.err
  fi
done
python - <<PY
import json
for l in open("$out"):
    d = json.loads(l)
    print(d["n_gpus"], "%.4g" % d["value"], "ms/step %.3f" % d["ms_per_step"], "frac %.3f" % d["roofline"]["frac"],
          "e2e %.4g" % d["e2e"]["value"], "allreduce_ms %.3f" % d["budgets"]["allreduce_ms"])
PY
