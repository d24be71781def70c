#!/bin/bash
# round 2, step c: chained stage launches + pointer-increment addressing.  Chain tests first, under their own timeout.
cd "$(dirname "$0")/.."
o=gpurun_out; tag=r02c
timeout 600 python -m pytest tests/test_chain.py -x -q > $o/${tag}_chain.log 2>&1; echo "chain rc $?" >> $o/${tag}_chain.log; tail -5 $o/${tag}_chain.log
grep -q "chain rc 0" $o/${tag}_chain.log || exit 1
timeout 300 python tools/chain_bench.py 40 > $o/${tag}_chain_bench.json 2> $o/${tag}_chain_bench.err; cat $o/${tag}_chain_bench.err | cut -c1-400
timeout 1500 python -m pytest tests -m gpu -x -q > $o/${tag}_pytest_gpu.log 2>&1; echo "pytest rc $?" >> $o/${tag}_pytest_gpu.log; tail -5 $o/${tag}_pytest_gpu.log
timeout 900 python tools/variants.py > $o/${tag}_variants.json 2> $o/${tag}_variants.err; cut -c1-250 $o/${tag}_variants.err
