#!/bin/bash
# round 2 call f (1 GPU): which ingredient of the TMA-staged persistent GEMM makes the factorisation nondeterministic (call e)
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python tools/diag_determinism.py 16384 32768 2>&1 | tee $O/r02f_diag.log | tail -70
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "extreme or K_entrywise" -p no:cacheprovider 2>&1 | tail -15 | tee $O/r02f_pytest.log
