#!/bin/bash
# GPU call: validate the CTA-pair GEMM (debug build first: a stuck mbarrier traps instead of hanging), then the suite + A/B benches
set -x
mkdir -p gpurun_out
EDB_LIB=$PWD/editor_b200/lib_dbg/libeditor_b200.so timeout 240 python -m pytest tests/test_gemm_gpu.py -x -q 2>&1 | tail -40 > gpurun_out/t1_gemm_dbg.log
if grep -q "passed" gpurun_out/t1_gemm_dbg.log && ! grep -q "failed" gpurun_out/t1_gemm_dbg.log; then
  PAIR_OK=1
else
  PAIR_OK=0; export EDB_GEMM_MODE=1
fi
echo "PAIR_OK=$PAIR_OK" > gpurun_out/t0_status.log
if [ $PAIR_OK = 1 ]; then
  timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/t2_pytest.log
else
  timeout 600 python -m pytest tests -m gpu -q -k "not pair" 2>&1 | tail -15 > gpurun_out/t2_pytest.log
fi
timeout 300 python tools/gemm_bench.py > gpurun_out/t3_gemm_bench.log 2>&1
timeout 400 python bench.py --no-cpu-baseline --gemm-mode 1 > gpurun_out/t4_bench_single.json 2> gpurun_out/t4.err
if [ $PAIR_OK = 1 ]; then
  timeout 400 python bench.py --no-cpu-baseline --gemm-mode 0 > gpurun_out/t5_bench_pair.json 2> gpurun_out/t5.err
fi
tail -3 gpurun_out/t1_gemm_dbg.log gpurun_out/t2_pytest.log; cat gpurun_out/t3_gemm_bench.log
