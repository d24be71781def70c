#!/bin/bash
# round 2, step e: straight-line ICE closures, per-column heat parameters, Integrator on lh_soil_run, e2e on lh_soil_run
cd "$(dirname "$0")/.."
o=gpurun_out; tag=r02e
timeout 1800 python -m pytest tests -m gpu -x -q > $o/${tag}_pytest_gpu.log 2>&1; echo "pytest rc $?" >> $o/${tag}_pytest_gpu.log; tail -8 $o/${tag}_pytest_gpu.log
timeout 900 python bench.py --no-cpu-baseline > $o/${tag}_bench.json 2> $o/${tag}_bench.err; echo "bench rc $?"; tail -5 $o/${tag}_bench.err
python - <<'P'
import json
d=json.load(open('gpurun_out/r02e_bench.json'))
print('value', d['value'], 'first', d['sustained']['value_first_block'], 'clk', d['clocks']['sm_mhz'], 'frac', d['roofline']['frac'], d['roofline']['frac_on_wire'])
print('e2e', d['e2e']['value'], d['e2e']['seconds'], d['e2e'].get('numa_bound'))
for k,v in d['extra']['variants'].items(): print(k, '%.3e'%v.get('cell_steps_per_s',0), round(v.get('frac_contract',0),3), round(v.get('frac_on_wire',0),3))
P
