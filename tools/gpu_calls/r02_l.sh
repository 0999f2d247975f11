#!/bin/bash
# round 2 call l (1 GPU): what about the kernel that precedes a TMA-staged update makes it go wrong (stand-alone reproduction of call k)
mkdir -p gpurun_out
timeout 900 ./tools/micro_dgemm pipeline 6 2>&1 | tee gpurun_out/r02l_pipeline_check.log
