#!/bin/bash
# round 2 call q (1 GPU): do generic->async proxy fences around the reuse of a shared-memory stage remove the corruption of the TMA-staged bulk update?
mkdir -p gpurun_out
timeout 600 python tools/diag_determinism.py 16384 2>&1 | cut -c1-200 | tee gpurun_out/r02q_diag.log | tail -40
