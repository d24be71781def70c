#!/bin/bash
# round 2 scaling: 2-GPU bitwise shard test (kept log), then bench.py at N = 1, 2, 4, 8 like the driver's scaling run.
tag=${1:-r02s}; nmax=${2:-8}
cd "$(dirname "$0")/.."
o=gpurun_out
timeout 600 python -m pytest tests/test_multi_gpu.py -x -q -m gpu -rA > $o/${tag}_multi_gpu_test.log 2>&1; echo "multi-gpu test rc $?" >> $o/${tag}_multi_gpu_test.log; tail -8 $o/${tag}_multi_gpu_test.log
out=$o/${tag}_scale.jsonl
: > $out
for n in 1 2 4 8; do
  [ $n -gt $nmax ] && break
  if [ $n -eq 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu-baseline --no-variants >> $out 2>> $o/${tag}_scale.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) \
      bench.py --gpus $n --steps 20 --warmup 3 >> $out 2>> $o/${tag}_scale.err
  fi
done
python - <<PY
import json
base = None
for l in open("$out"):
    d = json.loads(l)
    base = base or d
    print(d["n_gpus"], "%.4g" % d["value"], "eff %.3f" % (d["value"] / base["value"] / d["n_gpus"]), "ms/step %.4f" % d["ms_per_step"],
          "frac %.3f" % d["roofline"]["frac"], "e2e %.4g" % d["e2e"]["value"], "e2e eff %.3f" % (d["e2e"]["value"] / base["e2e"]["value"] / d["n_gpus"]),
          "shards", d["e2e"]["shards_per_gpu"], "numa", d["e2e"].get("numa_bound"), "clk", d["clocks"]["sm_mhz"], d["roofline"]["variant"][-60:])
PY
tail -3 $o/${tag}_scale.err
