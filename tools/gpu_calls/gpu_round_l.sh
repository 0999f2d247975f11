#!/bin/bash
# GPU call l (1 GPU): Cholesky critical-path fast path (row fix) vs the old schedule; full parity suite; default bench.
TAG=${1:-r01l}
O=gpurun_out
mkdir -p $O
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee $O/pytest_gpu_$TAG.log
run_bench() { name=$1; shift; timeout 900 python bench.py "$@" > $O/bench_${name}_$TAG.json 2> $O/bench_${name}_$TAG.err; python - <<PY
import json
try:
    d = json.loads([l for l in open("$O/bench_${name}_$TAG.json") if l.startswith("{")][-1])
    print("$name", {k: d[k] for k in ("value", "ms_per_step", "phases_ms", "cholesky_tflops")}, "e2e", d["e2e"]["value"], (d["cpu_baseline"] or {}).get("max_rel_err_mean_vs_gpu"))
except Exception as e:
    print("bench $name failed", e); print(open("$O/bench_${name}_$TAG.err").read()[-3000:])
PY
}
run_bench c2_fp64 --workload c2 --steps 10 --warmup 3
run_bench c2_fp64_nofast --workload c2 --steps 10 --warmup 3 --no-cpu --opt fastdiag=0
run_bench c2_tf32 --workload c2 --precision tf32 --steps 5 --warmup 3 --no-cpu
run_bench c4_fp64 --workload c4 --steps 2 --warmup 3 --no-cpu
run_bench c4_tf32 --workload c4 --precision tf32 --steps 3 --warmup 3 --no-cpu
run_bench c1_fp64 --workload c1 --steps 20 --warmup 3
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 | tee $O/bench_ref_c2_$TAG.json | cut -c1-400
