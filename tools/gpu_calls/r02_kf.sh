#!/bin/bash
# round 2 call kf (1 GPU): K-build v6 final form (knobs removed) -- micro_kbuild (v4 frozen vs v6, accuracy), parity tests of the rebuilt library,
# ncu --set full of the Matern-5/2 and ExpQuad product instantiations
mkdir -p gpurun_out
O=gpurun_out
timeout 300 ./tools/micro_kbuild 32768 > $O/r02kf_micro_kbuild.log 2>&1; echo "micro rc=$?"; grep "kind\|v4\|v6\|max rel" $O/r02kf_micro_kbuild.log | cut -c1-200
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_zz_late_gpu.py -m gpu -q -x -p no:cacheprovider > $O/r02kf_pytest_parity.log 2>&1; echo "pytest rc=$?" >> $O/r02kf_pytest_parity.log
tail -5 $O/r02kf_pytest_parity.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:kbuild_persist_kernel --launch-skip 5 --launch-count 1 -f -o $O/r02kf_ncu_kbuild_matern ./tools/micro_kbuild 32768 one > $O/r02kf_ncu_kbuild_matern.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:kbuild_persist_kernel --launch-skip 2 --launch-count 1 -f -o $O/r02kf_ncu_kbuild_expquad ./tools/micro_kbuild 32768 one > $O/r02kf_ncu_kbuild_expquad.log 2>&1
ls -la $O | grep r02kf
