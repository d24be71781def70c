#!/bin/bash
cd "$(dirname "$0")/.."
o=gpurun_out; tag=r02g
timeout 1800 python -m pytest tests -m gpu -x -q > $o/${tag}_pytest_gpu.log 2>&1; echo "pytest rc $?" >> $o/${tag}_pytest_gpu.log; tail -8 $o/${tag}_pytest_gpu.log
bash tools/gpu_r02_profile.sh r02p
