#!/bin/bash
# round 2 call b (1 GPU): persistent K-build kernel -- micro-benchmark (variants, accuracy), full GPU suite on it, bench lines.
mkdir -p gpurun_out
timeout 600 ./tools/micro_kbuild 32768 2>&1 | tee gpurun_out/micro_kbuild_r02b.log
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -25 | tee gpurun_out/pytest_gpu_r02b.log
for w in c2 c4 c3; do timeout 600 python bench.py --no-cpu --workload $w --steps 5 2>&1 | tail -1 | tee gpurun_out/bench_r02b_$w.log; done
timeout 600 python bench.py --no-cpu --workload c4 --steps 5 --opt fp64_panel=8 2>&1 | tail -1 | tee gpurun_out/bench_r02b_c4_panel8.log
timeout 600 python bench.py --no-cpu --workload c4 --steps 5 --opt fp64_panel=4 --opt solve_streams=4 2>&1 | tail -1 | tee gpurun_out/bench_r02b_c4_panel4_streams4.log
