#!/bin/bash
# round 2 call k (1 GPU): bulk-update call pattern of the factorisation reproduced stand-alone (static / freshly written / concurrent
# neighbour writes / back-to-back), TMA-staged vs cp.async; plus two more Cholesky / solve option points at c4
mkdir -p gpurun_out
O=gpurun_out
timeout 600 ./tools/micro_dgemm pipeline 6 2>&1 | tee $O/r02k_pipeline_check.log
for o in "fp64_panel=16" "fp64_panel=8 --opt solve_streams=4"; do
  tag=$(echo "$o" | tr -d ' ' | tr '=' '_' | sed 's/--opt/+/g')
  timeout 300 python bench.py --workload c4 --steps 4 --no-cpu --no-also --opt $o 2>&1 | tail -1 > $O/r02k_bench_c4_$tag.log
  python - "$O/r02k_bench_c4_$tag.log" "c4 $tag" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read())
    print(sys.argv[2], "ms/step %.1f" % d["ms_per_step"], {k: round(v,2) for k,v in d["phases_ms"].items()}, "e2e ms %.1f" % d["e2e"]["ms_per_step"])
except Exception as e:
    print(sys.argv[2], "FAILED", open(sys.argv[1]).read()[-300:])
PY
done | tee $O/r02k_bench_c4_summary.txt
