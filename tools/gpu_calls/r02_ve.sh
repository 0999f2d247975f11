#!/bin/bash
# round 2 call ve (1 GPU): one find_MAP objective evaluation (factorise + value + gradient) at the headline size, N = 32768 d = 8 Matern-5/2
mkdir -p gpurun_out
timeout 100 python tools/grad_timing.py 32768 8 Matern52 2>&1 | tail -2 | tee gpurun_out/r02ve_grad_timing_c4.txt | cut -c1-500
