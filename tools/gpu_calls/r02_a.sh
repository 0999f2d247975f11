#!/bin/bash
# round 2 call a (1 GPU): full GPU suite without the blanket xfail (tracebacks kept), then the queued option measurements.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x --deselect tests/test_zz_late_gpu.py 2>&1 | tail -15 | tee gpurun_out/pytest_gpu_r02a_main.log
timeout 900 python -m pytest tests/test_zz_late_gpu.py -m gpu -q 2>&1 | tail -150 | tee gpurun_out/pytest_gpu_r02a_late.log
timeout 600 python bench.py 2>&1 | tail -2 | tee gpurun_out/bench_r02a.log
for s in 2 4; do timeout 300 python bench.py --no-cpu --opt solve_streams=$s 2>&1 | tail -1 | tee gpurun_out/bench_r02a_streams$s.log; done
timeout 300 ./tools/micro_dgemm 2>&1 | tee gpurun_out/micro_dgemm_r02a.log
timeout 300 python tools/fused_timing.py 2>&1 | tail -8 | tee gpurun_out/fused_timing_r02a.log
timeout 300 python tools/chol_trace.py 2>&1 | tail -40 | tee gpurun_out/chol_trace_r02a.log
timeout 120 ./tools/micro_potrf 2>&1 | tail -10 | tee gpurun_out/micro_potrf_r02a.log
for o in small_diag=1 green_sms=8; do timeout 300 python bench.py --no-cpu --opt $o 2>&1 | tail -1 | tee gpurun_out/bench_r02a_$o.log; done
for w in c2 c4; do for o in fp64_panel=0 fp64_panel=2 fp64_panel=4; do timeout 400 python bench.py --no-cpu --workload $w --steps 5 --opt $o 2>&1 | tail -1 | tee gpurun_out/bench_r02a_${w}_$o.log; done; done
timeout 600 python tools/kron_timing.py 2>&1 | tail -8 | tee gpurun_out/kron_timing_r02a.log
