#!/bin/bash
# round 2 call t (1 GPU): final validation of the round's product state -- full GPU suite, smoke, default bench both arms, launch list of one
# c4 step, DRAM traffic of the predict solve's GEMM launches, ncu --set full of the K-build (Matern-5/2, product instantiation) and of the
# TMA-staged GEMM, Kronecker timing
mkdir -p gpurun_out
O=gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --durations=8 -p no:cacheprovider > $O/r02t_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/r02t_pytest_gpu.log
tail -12 $O/r02t_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $O/r02t_smoke.log | cut -c1-300
timeout 900 python bench.py --steps 5 > $O/r02t_bench_default.log 2>&1; tail -1 $O/r02t_bench_default.log | cut -c1-600
( time timeout 900 python bench.py --impl reference --steps 20 --warmup 5 ) > $O/r02t_bench_reference.log 2>&1; tail -5 $O/r02t_bench_reference.log | cut -c1-400
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 20000 --csv --log-file $O/r02t_launches_c4.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-also > $O/r02t_launches_c4.log 2>&1
python tools/launch_summary.py $O/r02t_launches_c4.csv 2>&1 | tail -25 | tee $O/r02t_launches_c4_summary.txt
rm -f $O/r02t_launches_c4.csv
timeout 900 ncu --profile-from-start off --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $O/r02t_solve_traffic_c4.csv -k regex:dgemm python tools/prof_factorize.py 32768 1 Matern52 fp64 predict > $O/r02t_solve_traffic_c4.log 2>&1
tail -2 $O/r02t_solve_traffic_c4.log | cut -c1-300; wc -l $O/r02t_solve_traffic_c4.csv
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kbuild_persist -c 1 -f -o $O/r02t_ncu_kbuild_matern python tools/prof_factorize.py 32768 1 Matern52 > $O/r02t_ncu_kbuild.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dgemm_tma -s 2 -c 1 -f -o $O/r02t_ncu_dgemm_tma ./tools/micro_dgemm > $O/r02t_ncu_dgemm.log 2>&1
timeout 600 python tools/kron_timing.py 2>&1 | tail -8 | tee $O/r02t_kron_timing.log | cut -c1-300
ls -la $O | grep r02t
