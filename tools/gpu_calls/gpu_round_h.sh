#!/bin/bash
# GPU call h (1 GPU): K-build v3 (strip + cp.async prefetch) parity + timing + ncu.
TAG=${1:-r01h}
O=gpurun_out
mkdir -p $O
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $O/pytest_gpu_$TAG.log
run_bench() { name=$1; shift; timeout 900 python bench.py "$@" > $O/bench_${name}_$TAG.json 2> $O/bench_${name}_$TAG.err; python - <<PY
import json
try:
    d = json.loads([l for l in open("$O/bench_${name}_$TAG.json") if l.startswith("{")][-1])
    print("$name", {k: d[k] for k in ("value", "ms_per_step", "phases_ms")}, "kb", d["roofline_kbuild"]["achieved"], d["roofline_kbuild"]["frac"], "e2e", d["e2e"]["value"], d["e2e"].get("host_call_ms_per_step"), "err", (d["cpu_baseline"] or {}).get("max_rel_err_mean_vs_gpu"))
except Exception as e:
    print("bench $name failed", e); print(open("$O/bench_${name}_$TAG.err").read()[-3000:])
PY
}
run_bench c2_fp64 --workload c2 --steps 10 --warmup 3
run_bench c4_fp64 --workload c4 --steps 2 --warmup 3 --no-cpu
run_bench c4_tf32 --workload c4 --precision tf32 --steps 3 --warmup 3 --no-cpu
echo "== ncu kbuild_strip"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:kbuild_strip -c 2 -o $O/prof_kbuild4_$TAG -f python tools/prof_factorize.py 8192 1 > $O/ncu_kbuild4_$TAG.log 2>&1
