#!/bin/bash
# round 2 call ka (1 GPU): K-build v5 (lean persistent kernel) -- fp64 pipe ceilings, micro_kbuild (v4 frozen vs v5, strips, occupancies, accuracy),
# ncu --set full of the v5 Matern-5/2 and ExpQuad product instantiations, full GPU suite on the rebuilt library
mkdir -p gpurun_out
O=gpurun_out
timeout 120 ./tools/micro_fp64 > $O/r02ka_micro_fp64.log 2>&1; tail -12 $O/r02ka_micro_fp64.log | cut -c1-300
timeout 300 ./tools/micro_kbuild 32768 > $O/r02ka_micro_kbuild.log 2>&1; echo "micro rc=$?"; cut -c1-330 $O/r02ka_micro_kbuild.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:kbuild_persist_kernel --launch-skip 5 --launch-count 1 -f -o $O/r02ka_ncu_kbuild_matern ./tools/micro_kbuild 32768 one > $O/r02ka_ncu_kbuild_matern.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:kbuild_persist_kernel --launch-skip 2 --launch-count 1 -f -o $O/r02ka_ncu_kbuild_expquad ./tools/micro_kbuild 32768 one > $O/r02ka_ncu_kbuild_expquad.log 2>&1
timeout 1200 python -m pytest tests -m gpu -q -x --durations=5 -p no:cacheprovider > $O/r02ka_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/r02ka_pytest_gpu.log
tail -12 $O/r02ka_pytest_gpu.log
ls -la $O | grep r02ka
