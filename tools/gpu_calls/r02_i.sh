#!/bin/bash
# round 2 call i (1 GPU): which USE of the TMA-staged GEMM corrupts the factorisation (mask per usage), and does the row stride matter (Np = 16384 vs 16512)
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python tools/diag_determinism.py 16383 16384 2>&1 | cut -c1-200 | tee $O/r02i_diag.log | tail -50
