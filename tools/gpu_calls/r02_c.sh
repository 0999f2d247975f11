#!/bin/bash
# round 2 call c (1 GPU): K-build write ceilings + fixed exp constant, full GPU suite incl. the C3/C4 full-size parity tests, new default bench (c4 fp64) + reference arm, ncu of the K-build.
mkdir -p gpurun_out
timeout 600 ./tools/micro_kbuild 32768 2>&1 | tee gpurun_out/micro_kbuild_r02c.log
timeout 1500 python -m pytest tests -m gpu -q -x --durations=8 2>&1 | tail -30 | tee gpurun_out/pytest_gpu_r02c.log
timeout 900 python bench.py --steps 5 2>&1 | tail -1 | tee gpurun_out/bench_r02c_default.log
( time timeout 900 python bench.py --impl reference --steps 20 --warmup 5 ) 2>&1 | tail -5 | tee gpurun_out/bench_r02c_reference.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kbuild_persist -c 4 -f -o gpurun_out/ncu_kbuild_r02c ./tools/micro_kbuild 32768 one > gpurun_out/ncu_kbuild_r02c.log 2>&1
nproc; free -g | head -2
