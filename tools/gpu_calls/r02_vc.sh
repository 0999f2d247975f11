#!/bin/bash
# round 2 call vc (1 GPU): the round's final library (K-build v6 + rescheduled diagonal-panel kernel) -- full GPU suite, smoke, default bench
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > $O/r02vc_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/r02vc_pytest_gpu.log
tail -4 $O/r02vc_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee $O/r02vc_smoke.log | cut -c1-200
timeout 600 python bench.py --steps 5 > $O/r02vc_bench_default.log 2>&1; tail -1 $O/r02vc_bench_default.log | cut -c1-300
