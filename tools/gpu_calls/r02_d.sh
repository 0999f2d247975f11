#!/bin/bash
# round 2 call d (1 GPU): first device run of the TMA-staged fp64 GEMM and the persistent K-build, full GPU suite
# (no -x: every failure with its traceback), default bench (c4 fp64) + reference arm, launch list, ncu captures.
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv | tee $O/r02d_gpu.log
nproc | tee -a $O/r02d_gpu.log; free -g | head -2 | tee -a $O/r02d_gpu.log
timeout 240 ./tools/micro_dgemm > $O/r02d_micro_dgemm.log 2>&1; echo "micro_dgemm rc=$?" | tee -a $O/r02d_micro_dgemm.log
tail -45 $O/r02d_micro_dgemm.log
# if the TMA kernel is wrong or hangs, the rest of the call runs on the cp.async kernel (and says so)
if ! grep -q "check TMA-staged" $O/r02d_micro_dgemm.log || ! python - <<'E'
import re, sys
s = open("gpurun_out/r02d_micro_dgemm.log").read()
m = re.search(r"check TMA-staged.*max \|diff\| ([0-9.e+-]+) at scale ([0-9.e+-]+)", s)
sys.exit(0 if m and float(m.group(1)) <= 1e-10 * max(1.0, float(m.group(2))) else 1)
E
then export GB2_OPTS="dgemm_tma=0"; echo "TMA GEMM check FAILED -> GB2_OPTS=$GB2_OPTS" | tee $O/r02d_tma_disabled.log; fi
timeout 400 ./tools/micro_kbuild 32768 > $O/r02d_micro_kbuild.log 2>&1; echo "micro_kbuild rc=$?" >> $O/r02d_micro_kbuild.log; tail -40 $O/r02d_micro_kbuild.log
timeout 1800 python -m pytest tests -m gpu -q --durations=10 -p no:cacheprovider > $O/r02d_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/r02d_pytest_gpu.log
tail -60 $O/r02d_pytest_gpu.log
timeout 900 python bench.py --steps 5 > $O/r02d_bench_default.log 2>&1; tail -1 $O/r02d_bench_default.log
GB2_OPTS="dgemm_tma=0" timeout 600 python bench.py --steps 5 --no-cpu --no-also > $O/r02d_bench_c4_notma.log 2>&1; tail -1 $O/r02d_bench_c4_notma.log | cut -c1-1500
( time timeout 900 python bench.py --impl reference --steps 20 --warmup 5 ) > $O/r02d_bench_reference.log 2>&1; tail -5 $O/r02d_bench_reference.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file $O/r02d_launches_c4.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-also > $O/r02d_launches_c4.log 2>&1
python tools/launch_summary.py $O/r02d_launches_c4.csv 2>&1 | tail -25 | tee $O/r02d_launches_c4_summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kbuild_persist -c 2 -f -o $O/r02d_ncu_kbuild ./tools/micro_kbuild 32768 one > $O/r02d_ncu_kbuild.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dgemm_tma -s 4 -c 2 -f -o $O/r02d_ncu_dgemm ./tools/micro_dgemm > $O/r02d_ncu_dgemm.log 2>&1
ls -la $O | tail -30
