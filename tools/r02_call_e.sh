#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python tools/graph_probe.py > gpurun_out/e_graph_probe.txt 2>&1
EDB_LIB=$PWD/editor_b200/lib_dbg/libeditor_b200.so timeout 600 python -m pytest tests/test_ops_gpu.py -q -x -k "attention" 2>&1 | tail -n 30 > gpurun_out/e_attn_dbg.log
if grep -q "passed" gpurun_out/e_attn_dbg.log && ! grep -q "failed" gpurun_out/e_attn_dbg.log; then
  timeout 300 python tools/attn_bench.py > gpurun_out/e_attn_bench.txt 2>&1
  timeout 900 python -m pytest tests/test_dropin_do_train_gpu.py tests/test_model_gpu.py -m gpu -q --tb=short > gpurun_out/e_pytest_sel.log 2>&1
fi
ls -la gpurun_out
