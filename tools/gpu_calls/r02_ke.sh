#!/bin/bash
# round 2 call ke (1 GPU): K-build v6 after the instruction trims (two-sided one-pair clip, group-wise flush): G x occupancy
mkdir -p gpurun_out
O=gpurun_out
timeout 300 ./tools/micro_kbuild 32768 > $O/r02ke_micro_kbuild.log 2>&1; echo "micro rc=$?"; cut -c1-200 $O/r02ke_micro_kbuild.log
