#!/bin/bash
# round 2 call va (1 GPU): validation of the round's final library (K-build v6) -- full GPU suite, smoke, default bench (c4 fp64 + c2 / c3 / c4-tf32
# under "also"), per-step timeline of the C2 and C4 factorisations, launch list of one c4 step
mkdir -p gpurun_out
O=gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --durations=5 -p no:cacheprovider > $O/r02va_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/r02va_pytest_gpu.log
tail -10 $O/r02va_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $O/r02va_smoke.log | cut -c1-300
timeout 900 python bench.py --steps 5 > $O/r02va_bench_default.log 2>&1; tail -1 $O/r02va_bench_default.log | cut -c1-700
timeout 200 python tools/chol_trace.py 8192 8 > $O/r02va_chol_trace_c2.log 2>&1; grep '"config"' $O/r02va_chol_trace_c2.log | cut -c1-600
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 20000 --csv --log-file $O/r02va_launches_c4.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-also > $O/r02va_launches_c4.log 2>&1
python tools/launch_summary.py $O/r02va_launches_c4.csv 2>&1 | tail -25 | tee $O/r02va_launches_c4_summary.txt
rm -f $O/r02va_launches_c4.csv
ls -la $O | grep r02va
