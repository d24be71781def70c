#!/bin/bash
cd "$(dirname "$0")/.."
o=gpurun_out; tag=${1:-r02b}
timeout 1200 python -m pytest tests -m gpu -x -q > $o/${tag}_pytest_gpu.log 2>&1; echo "pytest rc $?" >> $o/${tag}_pytest_gpu.log
tail -15 $o/${tag}_pytest_gpu.log
timeout 900 python tools/variants.py ${2:-} > $o/${tag}_variants.json 2> $o/${tag}_variants.err; cut -c1-330 $o/${tag}_variants.err
