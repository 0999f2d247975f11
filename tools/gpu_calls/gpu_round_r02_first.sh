#!/bin/bash
# first GPU call of round 2 (1 GPU, ~12 min; give gpurun --timeout 1500): full GPU suite (includes the Kronecker-solve device tests that round 1 could not
# run), Kronecker vs dense cold-predict timing on BASELINE config 3, default bench line.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu_r02a.log
timeout 600 python tools/kron_timing.py 2>&1 | tail -8 | tee gpurun_out/kron_timing_r02a.log
timeout 600 python bench.py 2>&1 | tail -2 | tee gpurun_out/bench_r02a.log
# wave-tail filling of the predict solve: prediction rows as 2 / 4 concurrent slabs (default 1)
for s in 2 4; do timeout 300 python bench.py --no-cpu --opt solve_streams=$s 2>&1 | tail -1 | tee gpurun_out/bench_r02a_streams$s.log; done
# kernel-only rates of the DMMA GEMM per launch shape (full waves vs the recursion's partial waves)
timeout 300 ./tools/micro_dgemm 2>&1 | tee gpurun_out/micro_dgemm_r02a.log
# cold predict: factorise-then-solve (1/2/4 solve streams) against the fused one-pass entry point
timeout 300 python tools/fused_timing.py 2>&1 | tail -8 | tee gpurun_out/fused_timing_r02a.log
# (2+ GPUs, separate call)  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/replicate_timing.py
#                           python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 tools/c5_kron_demo.py
# timeline of the blocked Cholesky: is the diagonal kernel starved of an SM by the bulk update (222 KB of shared memory)?
timeout 300 python tools/chol_trace.py 2>&1 | tail -30 | tee gpurun_out/chol_trace_r02a.log
# diagonal-panel kernel: phase clocks of both variants + bit-identity of the small-footprint one
timeout 120 ./tools/micro_potrf 2>&1 | tail -10 | tee gpurun_out/micro_potrf_r02a.log
for o in small_diag=1 green_sms=8; do timeout 300 python bench.py --no-cpu --opt $o 2>&1 | tail -1 | tee gpurun_out/bench_r02a_$o.log; done
# two-level blocking of the fp64 factorisation (panels of 2 / 4 column blocks), C2 and C4
for w in c2 c4; do for o in fp64_panel=2 fp64_panel=4; do timeout 400 python bench.py --no-cpu --workload $w --steps 5 --opt $o 2>&1 | tail -1 | tee gpurun_out/bench_r02a_${w}_$o.log; done; done
