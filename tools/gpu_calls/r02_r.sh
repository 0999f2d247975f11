#!/bin/bash
# round 2 call r (1 GPU): TMA-staged fp64 GEMM with the stage-release proxy fence = product default again.  Control (fence off) vs product
# in the factorisation, stand-alone rates with the fence, full GPU suite, default bench both arms, launch list, ncu captures.
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python tools/diag_determinism.py 16384 32768 2>&1 | cut -c1-200 | tee $O/r02r_diag.log | tail -70
timeout 300 ./tools/micro_dgemm 2>&1 | head -52 | tee $O/r02r_micro_dgemm.log | head -30
timeout 1800 python -m pytest tests -m gpu -q --durations=8 -p no:cacheprovider > $O/r02r_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/r02r_pytest_gpu.log
tail -15 $O/r02r_pytest_gpu.log
timeout 900 python bench.py --steps 5 > $O/r02r_bench_default.log 2>&1; tail -1 $O/r02r_bench_default.log | cut -c1-1500
timeout 600 python bench.py --steps 5 --no-cpu --no-also --opt dgemm_tma=0 2>&1 | tail -1 > $O/r02r_bench_c4_cpasync.log; cut -c1-700 $O/r02r_bench_c4_cpasync.log
