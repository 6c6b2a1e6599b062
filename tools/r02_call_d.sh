#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python tools/graph_probe.py > gpurun_out/d_graph_probe.txt 2>&1
timeout 300 python tools/scaler_probe.py > gpurun_out/d_scaler_probe.txt 2>&1
timeout 1200 python -m pytest tests/test_dropin_do_train_gpu.py -m gpu -q --tb=short > gpurun_out/d_pytest_dropin.log 2>&1
ls -la gpurun_out
