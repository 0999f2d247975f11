#!/bin/bash
# round 2 call vf (1 GPU): per-step timeline of the C2 factorisation with the rescheduled diagonal-panel kernel
mkdir -p gpurun_out
timeout 100 python tools/chol_trace.py 8192 8 > gpurun_out/r02vf_chol_trace_c2.log 2>&1; grep '"config"' gpurun_out/r02vf_chol_trace_c2.log | head -2 | cut -c1-700
