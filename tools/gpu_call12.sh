#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -n 25 > gpurun_out/t2_pytest.log
timeout 600 python bench.py > gpurun_out/bench_own.json 2> gpurun_out/bench_own.err
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
tail -n 6 gpurun_out/t2_pytest.log; tail -n 2 gpurun_out/smoke.log; tail -n 3 gpurun_out/bench_own.err; cut -c1-240 gpurun_out/bench_own.json
