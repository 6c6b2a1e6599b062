#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python tools/graph_probe2.py > gpurun_out/g_graph_probe2.txt 2>&1
timeout 600 python -m pytest tests/test_augment_gpu.py tests/test_ops_gpu.py -m gpu -q --tb=short -k "augment or attention" > gpurun_out/g_pytest_sel.log 2>&1
timeout 200 python tools/attn_bench.py > gpurun_out/g_attn_bench.txt 2>&1
ls -la gpurun_out
