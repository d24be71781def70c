#!/bin/bash
# round 2, step f: PrescribedAtmosForcing, per-column fluxes and heat parameters, Integrator on lh_soil_run
cd "$(dirname "$0")/.."
o=gpurun_out; tag=r02f
timeout 1800 python -m pytest tests -m gpu -x -q > $o/${tag}_pytest_gpu.log 2>&1; echo "pytest rc $?" >> $o/${tag}_pytest_gpu.log; tail -8 $o/${tag}_pytest_gpu.log
