#!/bin/bash
# round 2 call h (1 GPU): does a generic->async proxy fence (or dropping the L2 promotion) remove the intermittent corruption?
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python tools/diag_determinism.py 16384 2>&1 | cut -c1-220 | tee $O/r02h_diag.log | tail -40
