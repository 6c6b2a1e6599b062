#!/bin/bash
set -x
mkdir -p gpurun_out
for v in lib lib_y lib lib_y; do
  echo "== $v" >> gpurun_out/l_ln_ab.txt
  EDB_LIB=$PWD/editor_b200/$v/libeditor_b200.so timeout 200 python tools/ln_bench.py >> gpurun_out/l_ln_ab.txt 2>&1
done
cat gpurun_out/l_ln_ab.txt
timeout 2400 python -m pytest tests -m gpu -q --tb=short > gpurun_out/l_pytest_gpu.log 2>&1
tail -15 gpurun_out/l_pytest_gpu.log
ls -la gpurun_out
