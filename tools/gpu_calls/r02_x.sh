#!/bin/bash
# round 2 call x (1 GPU): product kernel (stage fence) vs the true control (same kernel compiled WITHOUT the fence: the original schedule)
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=serial --format=csv,noheader | tee $O/r02x_gpu_id.log
timeout 600 python tools/diag_determinism.py 16384 32768 2>&1 | cut -c1-200 | tee $O/r02x_diag.log | grep -v "idx=-1" | tail -24
echo "bit-reproducible runs: $(grep -c 'idx=-1' $O/r02x_diag.log) of $(grep -c rep $O/r02x_diag.log)"
timeout 300 ./tools/micro_dgemm pipeline 30 2>&1 | tee $O/r02x_pipeline_check.log | tail -8
