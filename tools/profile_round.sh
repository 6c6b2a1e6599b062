#!/bin/bash
# Round profiling recipe (run under gpurun, 1 GPU): tests, bench, ncu launch list, ncu full captures, SFTS isolation.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -5 > gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench_own.json 2> gpurun_out/bench_own.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_kernel -s 420 -c 6 -o gpurun_out/prof_gemm \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_tc -s 40 -c 2 -o gpurun_out/prof_attn_fwd \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_tc_bwd -s 12 -c 1 -o gpurun_out/prof_attn_bwd \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"rollout|freq_counts|ln_bwd_kernel|ln_fwd_kernel" -s 60 -c 4 -o gpurun_out/prof_sfts_ln \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
timeout 600 python tools/sfts_bench.py > gpurun_out/sfts_bench.json 2> gpurun_out/sfts_bench.err
ls -la gpurun_out
