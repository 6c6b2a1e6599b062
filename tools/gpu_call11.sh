#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 6 > gpurun_out/t2_pytest.log
for v in pf0 pf2 ""; do
  L=$PWD/editor_b200/lib${v:+_$v}/libeditor_b200.so
  echo "=== $L" >> gpurun_out/t7_prefetch_ab.log
  EDB_LIB=$L timeout 200 python tools/gemm_bench.py 2>&1 | grep -E "proj fwd|fc2 fwd|fc2 dgrad|sum per" >> gpurun_out/t7_prefetch_ab.log
done
for v in pf0 pf2 ""; do
  L=$PWD/editor_b200/lib${v:+_$v}/libeditor_b200.so
  echo "=== $L" >> gpurun_out/t7_prefetch_ab.log
  EDB_LIB=$L timeout 200 python tools/gemm_bench.py 2>&1 | grep -E "proj fwd|fc2 fwd|fc2 dgrad|sum per" >> gpurun_out/t7_prefetch_ab.log
done
tail -n 3 gpurun_out/t2_pytest.log; cat gpurun_out/t7_prefetch_ab.log
