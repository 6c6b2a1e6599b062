#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 15 > gpurun_out/t2_pytest.log
timeout 300 python tools/gemm_bench.py > gpurun_out/t3_gemm_bench.log 2>&1
timeout 400 python bench.py --no-cpu-baseline > gpurun_out/t5_bench_pair.json 2> gpurun_out/t5.err
tail -n 4 gpurun_out/t2_pytest.log; cat gpurun_out/t3_gemm_bench.log; cat gpurun_out/t5_bench_pair.json | cut -c1-300
