#!/bin/bash
# round 2 call s (1 GPU): Cholesky with the TMA-staged GEMM: bulk updates one CTA per tile (default now) vs persistent, panel widths; C2 too
mkdir -p gpurun_out
O=gpurun_out
summ() { python - "$1" "$2" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], "ms/step %.2f" % d["ms_per_step"], {k: round(v,2) for k,v in d["phases_ms"].items()})
except Exception as e:
    print(sys.argv[2], "FAILED", open(sys.argv[1]).read()[-300:])
PY
}
for o in "" "bulk_persistent=1" "fp64_panel=8" "fp64_panel=12" "fp64_panel=8 --opt bulk_persistent=1" "fp64_panel=0"; do
  tag=$(echo "$o" | tr -d ' ' | tr '=' '_' | sed 's/--opt/+/g'); [ -z "$tag" ] && tag=default
  if [ -z "$o" ]; then optarg=""; else optarg="--opt $o"; fi
  timeout 300 python bench.py --workload c4 --steps 4 --no-cpu --no-also $optarg 2>&1 | tail -1 > $O/r02s_bench_c4_$tag.log
  summ $O/r02s_bench_c4_$tag.log "c4 $tag"
done | tee $O/r02s_bench_c4_summary.txt
for o in "" "bulk_persistent=1" "fp64_panel=4" "dgemm_tma=0"; do
  tag=$(echo "$o" | tr -d ' ' | tr '=' '_'); [ -z "$tag" ] && tag=default
  if [ -z "$o" ]; then optarg=""; else optarg="--opt $o"; fi
  timeout 300 python bench.py --workload c2 --steps 10 --no-cpu --no-also $optarg 2>&1 | tail -1 > $O/r02s_bench_c2_$tag.log
  summ $O/r02s_bench_c2_$tag.log "c2 $tag"
done | tee $O/r02s_bench_c2_summary.txt
