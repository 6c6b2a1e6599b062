#!/bin/bash
# round 2, call B: persistent attention kernels -- first under the mbarrier-timeout debug build (a wrong barrier traps instead
# of hanging the box), then the release build: full GPU tests, attention micro-bench, short bench
set -x
mkdir -p gpurun_out
EDB_LIB=$PWD/editor_b200/lib_dbg/libeditor_b200.so timeout 600 python -m pytest tests/test_ops_gpu.py -q -x -k "attention" 2>&1 | tail -n 30 > gpurun_out/b_attn_dbg.log
if ! grep -q "passed" gpurun_out/b_attn_dbg.log || grep -q "failed" gpurun_out/b_attn_dbg.log; then
  echo "debug-build attention tests failed: stopping" >> gpurun_out/b_attn_dbg.log
  exit 0
fi
timeout 300 python tools/attn_bench.py > gpurun_out/b_attn_bench.txt 2>&1
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -n 60 > gpurun_out/b_pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-ref-gpu --no-sfts --no-cpu-baseline > gpurun_out/b_bench_own.json 2> gpurun_out/b_bench_own.err
ls -la gpurun_out
