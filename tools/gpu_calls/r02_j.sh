#!/bin/bash
# round 2 call j (1 GPU): (1) which USE of the TMA-staged GEMM corrupts the factorisation, power-of-two row stride vs not;
# (2) Cholesky / solve options on the cp.async product path: two-level blocking, small diagonal kernel, solve streams, fused cold predict
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python tools/diag_determinism.py 16383 16384 2>&1 | cut -c1-200 | tee $O/r02j_diag.log | tail -50
for o in "" "fp64_panel=2" "fp64_panel=4" "fp64_panel=8" "small_diag=1" "solve_streams=2" "fp64_panel=4 --opt small_diag=1"; do
  tag=$(echo "$o" | tr -d ' ' | tr '=' '_' | sed 's/--opt/+/g'); [ -z "$tag" ] && tag=default
  if [ -z "$o" ]; then optarg=""; else optarg="--opt $o"; fi
  timeout 300 python bench.py --workload c4 --steps 4 --no-cpu --no-also $optarg 2>&1 | tail -1 > $O/r02j_bench_c4_$tag.log
  python - "$O/r02j_bench_c4_$tag.log" "c4 $tag" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read())
    print(sys.argv[2], "ms/step %.1f" % d["ms_per_step"], {k: round(v,2) for k,v in d["phases_ms"].items()}, "e2e ms %.1f" % d["e2e"]["ms_per_step"])
except Exception as e:
    print(sys.argv[2], "FAILED", open(sys.argv[1]).read()[-300:])
PY
done | tee $O/r02j_bench_c4_summary.txt
for o in "" "fp64_panel=2" "fp64_panel=4" "small_diag=1" "green_sms=8" "solve_streams=2" "solve_streams=4"; do
  tag=$(echo "$o" | tr -d ' ' | tr '=' '_'); [ -z "$tag" ] && tag=default
  if [ -z "$o" ]; then optarg=""; else optarg="--opt $o"; fi
  timeout 300 python bench.py --workload c2 --steps 10 --no-cpu --no-also $optarg 2>&1 | tail -1 > $O/r02j_bench_c2_$tag.log
  python - "$O/r02j_bench_c2_$tag.log" "c2 $tag" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read())
    print(sys.argv[2], "ms/step %.2f" % d["ms_per_step"], {k: round(v,3) for k,v in d["phases_ms"].items()}, "e2e ms %.2f" % d["e2e"]["ms_per_step"])
except Exception as e:
    print(sys.argv[2], "FAILED", open(sys.argv[1]).read()[-300:])
PY
done | tee $O/r02j_bench_c2_summary.txt
timeout 300 python tools/fused_timing.py 2>&1 | tail -10 | tee $O/r02j_fused_timing_c2.log
timeout 400 python tools/fused_timing.py 32768 8 3 2>&1 | tail -10 | tee $O/r02j_fused_timing_c4.log
