#!/bin/bash
# round 2 call pa (1 GPU): diagonal-panel kernel with the inverse assembly and the factor write-back moved onto the idle warps of the
# factorisation phases -- phase clocks and residuals old vs new (same box), parity tests of the rebuilt library, C2 bench
mkdir -p gpurun_out
O=gpurun_out
( echo "== old (round-2 call t)"; timeout 60 ./tools/micro_potrf_old; echo "== new"; timeout 60 ./tools/micro_potrf ) > $O/r02pa_micro_potrf.log 2>&1; cat $O/r02pa_micro_potrf.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_zz_late_gpu.py -m gpu -q -x -p no:cacheprovider > $O/r02pa_pytest_parity.log 2>&1; echo "pytest rc=$?" >> $O/r02pa_pytest_parity.log
tail -4 $O/r02pa_pytest_parity.log
timeout 300 python bench.py --workload c2 --steps 10 --no-cpu --no-also 2>&1 | tail -1 > $O/r02pa_bench_c2.log
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02pa_bench_c2.log').read().strip().splitlines()[-1])
print("c2 ms/step %.2f"%d["ms_per_step"], {k: round(v,3) for k,v in d["phases_ms"].items()})
PY
