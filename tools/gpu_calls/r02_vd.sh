#!/bin/bash
# round 2 call vd (1 GPU): find_MAP on the device with the round's final kernels (fit() path: value + gradient per evaluation), N = 8192
mkdir -p gpurun_out
timeout 120 python tools/map_timing.py 8192 2>&1 | tail -2 | tee gpurun_out/r02vd_find_map_timing.txt | cut -c1-400
