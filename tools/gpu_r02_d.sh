#!/bin/bash
# round 2, step d: lh_soil_run / checkpoint / aux tables / async budgets / C client on the GPU, then the new bench line
cd "$(dirname "$0")/.."
o=gpurun_out; tag=r02d
timeout 900 python -m pytest tests/test_run_api.py tests/test_abi_client.py tests/test_chain.py -x -q -m gpu > $o/${tag}_new.log 2>&1; echo "new rc $?" >> $o/${tag}_new.log; tail -15 $o/${tag}_new.log
timeout 1500 python -m pytest tests -m gpu -x -q > $o/${tag}_pytest_gpu.log 2>&1; echo "pytest rc $?" >> $o/${tag}_pytest_gpu.log; tail -5 $o/${tag}_pytest_gpu.log
timeout 900 python bench.py > $o/${tag}_bench.json 2> $o/${tag}_bench.err; echo "bench rc $?"; tail -5 $o/${tag}_bench.err; cut -c1-1500 $o/${tag}_bench.json
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $o/${tag}_bench_ref.json 2> $o/${tag}_bench_ref.err; echo "ref rc $?"; cut -c1-900 $o/${tag}_bench_ref.json
