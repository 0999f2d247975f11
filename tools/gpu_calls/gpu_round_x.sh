#!/bin/bash
# GPU call x (1 GPU): Cholesky chain kept on the panel stream vs the old choreography; parity suite.
TAG=${1:-r01x}
O=gpurun_out
mkdir -p $O
run_bench() { name=$1; shift; timeout 900 python bench.py "$@" > $O/bench_${name}_$TAG.json 2> $O/bench_${name}_$TAG.err; python - <<PY
import json
try:
    d = json.loads([l for l in open("$O/bench_${name}_$TAG.json") if l.startswith("{")][-1])
    print("$name", {k: d[k] for k in ("value", "ms_per_step")}, "chol", d["phases_ms"]["cholesky_ms"], "solve", d["phases_ms"]["solve_ms"], "e2e", d["e2e"]["value"])
except Exception as e:
    print("bench $name failed", e); print(open("$O/bench_${name}_$TAG.err").read()[-3000:])
PY
}
run_bench c2_chain1 --workload c2 --steps 10 --warmup 3 --no-cpu
run_bench c2_chain0 --workload c2 --steps 10 --warmup 3 --no-cpu --opt chain_on_panel=0
run_bench c2_tf32_chain1 --workload c2 --precision tf32 --steps 5 --warmup 3 --no-cpu
run_bench c2_tf32_chain0 --workload c2 --precision tf32 --steps 5 --warmup 3 --no-cpu --opt chain_on_panel=0
run_bench c4_tf32_chain1 --workload c4 --precision tf32 --steps 3 --warmup 3 --no-cpu
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee $O/pytest_gpu_$TAG.log
