#!/bin/bash
# round 2 call n (1 GPU): full GPU suite + default bench (both arms) on the current defaults (cp.async GEMM, two-level blocking auto,
# solve_streams 4, persistent K-build with the range / clip fixes), launch list of one c4 step, ncu of the K-build and the bulk GEMM
mkdir -p gpurun_out
O=gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --durations=8 -p no:cacheprovider > $O/r02n_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/r02n_pytest_gpu.log
tail -40 $O/r02n_pytest_gpu.log
timeout 900 python bench.py --steps 5 > $O/r02n_bench_default.log 2>&1; tail -1 $O/r02n_bench_default.log | cut -c1-3000
( time timeout 900 python bench.py --impl reference --steps 20 --warmup 5 ) > $O/r02n_bench_reference.log 2>&1; tail -5 $O/r02n_bench_reference.log | cut -c1-1500
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $O/r02n_smoke.log
