#!/bin/bash
# round 2 call kd (1 GPU): K-build v6 ablations (no stores, no table lookup, no DMMA, no evaluation): where the time goes
mkdir -p gpurun_out
O=gpurun_out
timeout 300 ./tools/micro_kbuild 32768 > $O/r02kd_micro_kbuild.log 2>&1; echo "micro rc=$?"; cut -c1-200 $O/r02kd_micro_kbuild.log
