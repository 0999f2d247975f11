#!/bin/bash
# round 2 call kb (1 GPU): K-build v6 variants (producer-warp prefetch, CT column tiles per barrier, late use of the looked-up word,
# G entries evaluated side by side, occupancy) against the frozen v4 / v5 kernels, same box
mkdir -p gpurun_out
O=gpurun_out
timeout 300 ./tools/micro_kbuild 32768 > $O/r02kb_micro_kbuild.log 2>&1; echo "micro rc=$?"; cut -c1-200 $O/r02kb_micro_kbuild.log
