#!/bin/bash
# round 2 call w (1 GPU): the race control on one more box (product = stage fence on; control = fence off), solve slab timing
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=serial,uuid --format=csv,noheader | tee $O/r02w_gpu_id.log
timeout 600 python tools/diag_determinism.py 16384 32768 2>&1 | cut -c1-200 | tee $O/r02w_diag.log | grep -v "idx=-1" | tail -20
grep -c "idx=-1" $O/r02w_diag.log
timeout 300 python tools/solve_slab_timing.py 2>&1 | tail -5 | tee $O/r02w_solve_slab_timing.log
