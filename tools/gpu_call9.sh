#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 8 > gpurun_out/t2_pytest.log
timeout 200 python tools/attn_bench.py > gpurun_out/t6_attn_bench.log 2>&1
timeout 400 python bench.py --no-cpu-baseline > gpurun_out/t5_bench_pair.json 2> gpurun_out/t5.err
tail -n 3 gpurun_out/t2_pytest.log; cat gpurun_out/t6_attn_bench.log; cut -c1-260 gpurun_out/t5_bench_pair.json
