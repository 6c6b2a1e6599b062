#!/bin/bash
# round 2, call A: tests, own bench (with the reference GPU legs + SFTS isolation), reference arm, SFTS ncu captures
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/a_gpu.txt
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -n 40 > gpurun_out/a_pytest_gpu.log
timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/a_bench_own.json 2> gpurun_out/a_bench_own.err
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/a_bench_ref.json 2> gpurun_out/a_bench_ref.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"rollout_topk|freq_counts|sfts_pack_fwd" -s 3 -c 3 -f -o gpurun_out/prof_sfts \
    python tools/sfts_bench.py --batch 256 > /dev/null 2> gpurun_out/a_ncu_sfts.err
ls -la gpurun_out
