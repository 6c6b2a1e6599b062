#!/bin/bash
# round 2, call C: failing tests with full logs, CUDA-graph step, aux-prefetch A/B, ncu of the persistent attention kernels
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_dropin_do_train_gpu.py tests/test_bf16_spread_gpu.py tests/test_model_gpu.py -m gpu -q --tb=short -k "dropin or do_train or grad or graphed or gradscaler or bf16 or full_size_eval" > gpurun_out/c_pytest_sel.log 2>&1
timeout 300 python tools/gemm_bench.py > gpurun_out/c_gemm_bench_pf3.txt 2>&1
EDB_LIB=$PWD/editor_b200/lib_x/libeditor_b200.so timeout 300 python tools/gemm_bench.py > gpurun_out/c_gemm_bench_pf1.txt 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 --no-ref-gpu --no-sfts --no-cpu-baseline > gpurun_out/c_bench_graph.json 2> gpurun_out/c_bench_graph.err
timeout 600 python bench.py --steps 20 --warmup 5 --no-ref-gpu --no-sfts --no-cpu-baseline --no-graph > gpurun_out/c_bench_eager.json 2> gpurun_out/c_bench_eager.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_tc_fwd -s 2 -c 1 -f -o gpurun_out/prof_attn_fwd_r02 \
    python tools/attn_bench.py > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_tc_bwd -s 2 -c 1 -f -o gpurun_out/prof_attn_bwd_r02 \
    python tools/attn_bench.py > /dev/null 2>&1
ls -la gpurun_out
