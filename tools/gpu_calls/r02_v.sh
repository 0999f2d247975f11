#!/bin/bash
# round 2 call v (1 GPU): the public-API drop-in tests with the CUDA engine (real Regressor from baseline/_ref), and persistent bulk
# updates capped at fewer CTAs (SMs left to the chain) vs one CTA per tile
mkdir -p gpurun_out
O=gpurun_out
ls baseline/_ref | head -3
timeout 600 python -m pytest tests/test_dropin_gpu.py -m gpu -q -p no:cacheprovider 2>&1 | tail -8 | tee $O/r02v_pytest_dropin_gpu.log
summ() { python - "$1" "$2" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], "ms/step %.2f" % d["ms_per_step"], {k: round(v,2) for k,v in d["phases_ms"].items()})
except Exception as e:
    print(sys.argv[2], "FAILED", open(sys.argv[1]).read()[-300:])
PY
}
for o in "" "bulk_persistent=280" "bulk_persistent=264" "bulk_persistent=232"; do
  tag=$(echo "$o" | tr -d ' ' | tr '=' '_'); [ -z "$tag" ] && tag=default
  if [ -z "$o" ]; then optarg=""; else optarg="--opt $o"; fi
  timeout 300 python bench.py --workload c4 --steps 4 --no-cpu --no-also $optarg 2>&1 | tail -1 > $O/r02v_bench_c4_$tag.log
  summ $O/r02v_bench_c4_$tag.log "c4 $tag"
done | tee $O/r02v_bench_c4_summary.txt
for o in "" "bulk_persistent=264"; do
  tag=$(echo "$o" | tr -d ' ' | tr '=' '_'); [ -z "$tag" ] && tag=default
  if [ -z "$o" ]; then optarg=""; else optarg="--opt $o"; fi
  timeout 300 python bench.py --workload c2 --steps 10 --no-cpu --no-also $optarg 2>&1 | tail -1 > $O/r02v_bench_c2_$tag.log
  summ $O/r02v_bench_c2_$tag.log "c2 $tag"
done | tee $O/r02v_bench_c2_summary.txt
