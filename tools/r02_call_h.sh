#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python tools/graph_probe3.py > gpurun_out/h_graph_probe3.txt 2>&1
timeout 600 python -m pytest tests/test_augment_gpu.py -m gpu -q --tb=short > gpurun_out/h_pytest_aug.log 2>&1
ls -la gpurun_out
