#!/bin/bash
# round 2 call e (1 GPU): locate the nondeterminism seen in call d (full-size parity failures with the TMA-staged GEMM on)
mkdir -p gpurun_out
O=gpurun_out
timeout 300 ./tools/micro_dgemm 2>&1 | grep -i "check" | tee $O/r02e_micro_dgemm_checks.log
timeout 900 python tools/diag_determinism.py 2>&1 | tee $O/r02e_diag.log | tail -60
timeout 600 python -m pytest tests/test_fitc.py tests/test_gpu_parity.py -m gpu -q -x -k "fitc or extreme" -p no:cacheprovider 2>&1 | tail -30 | tee $O/r02e_pytest_fitc.log
