#!/bin/bash
# round 2 call g (1 GPU): persistent TMA GEMM in the factorisation's regime (k = 128, lower tile list), stand-alone exactness checks,
# racecheck of the same, and the factorisation with / without cross-tile prefetch
mkdir -p gpurun_out
O=gpurun_out
timeout 300 ./tools/micro_dgemm check 20 2>&1 | tee $O/r02g_micro_check.log
timeout 600 compute-sanitizer --tool racecheck --racecheck-report all ./tools/micro_dgemm check 1 > $O/r02g_racecheck.log 2>&1; tail -25 $O/r02g_racecheck.log
timeout 600 python tools/diag_determinism.py 16384 2>&1 | cut -c1-200 | tee $O/r02g_diag.log | tail -30
timeout 300 ./tools/micro_kbuild 32768 2>&1 | tee $O/r02g_micro_kbuild.log | tail -32
