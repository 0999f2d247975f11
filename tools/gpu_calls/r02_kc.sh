#!/bin/bash
# round 2 call kc (1 GPU): K-build v6 with the LSU-wavefront knobs (ncu of v5: l1tex data-pipe wavefronts 82 % of peak) -- conflict-free 16-copy exp table + quartic (TB),
# 128-byte store rows through a lane-pair exchange (ST), against the frozen v4 / v5 kernels, same box
mkdir -p gpurun_out
O=gpurun_out
timeout 300 ./tools/micro_kbuild 32768 > $O/r02kc_micro_kbuild.log 2>&1; echo "micro rc=$?"; cut -c1-200 $O/r02kc_micro_kbuild.log
