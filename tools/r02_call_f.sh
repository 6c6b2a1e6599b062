#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python tools/graph_probe.py > gpurun_out/f_graph_probe.txt 2>&1
for rep in 1 2; do
for v in lib lib_y lib_z; do
  echo "== $v" >> gpurun_out/f_attn_ab.txt
  EDB_LIB=$PWD/editor_b200/$v/libeditor_b200.so timeout 200 python tools/attn_bench.py >> gpurun_out/f_attn_ab.txt 2>&1
done
done
for v in lib lib_x; do
  echo "== $v" >> gpurun_out/f_gemm_ab.txt
  EDB_LIB=$PWD/editor_b200/$v/libeditor_b200.so timeout 300 python tools/gemm_bench.py >> gpurun_out/f_gemm_ab.txt 2>&1
done
timeout 300 python -m pytest tests/test_gemm_gpu.py -q -x 2>&1 | tail -5 > gpurun_out/f_gemm_tests.log
ls -la gpurun_out
