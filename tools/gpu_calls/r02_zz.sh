#!/bin/bash
# round 2 call zz (1 GPU): split-TF32 option sweep on c4 with the round-2 kernels; final GPU suite + smoke + default bench on the final library
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
for o in "" "tf32_nb=4" "tf32_nb=16" "tf32_leaf=2" "tf32_leaf=8"; do
  tag=$(echo "$o" | tr -d ' ' | tr '=' '_'); [ -z "$tag" ] && tag=default
  if [ -z "$o" ]; then optarg=""; else optarg="--opt $o"; fi
  timeout 300 python bench.py --workload c4 --precision tf32 --steps 5 --no-cpu --no-also $optarg 2>&1 | tail -1 > $O/r02zz_bench_c4_tf32_$tag.log
  summ $O/r02zz_bench_c4_tf32_$tag.log "c4 tf32 $tag"
done | tee $O/r02zz_bench_c4_tf32_summary.txt
timeout 1800 python -m pytest tests -m gpu -q -p no:cacheprovider > $O/r02zz_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/r02zz_pytest_gpu.log
tail -6 $O/r02zz_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $O/r02zz_smoke.log | cut -c1-200
timeout 900 python bench.py --steps 5 > $O/r02zz_bench_default.log 2>&1; tail -1 $O/r02zz_bench_default.log | cut -c1-400
