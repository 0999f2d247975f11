#!/bin/bash
# GPU call w (1 GPU): strip K-build occupancy bound 3 vs 4; final parity suite; final default bench + reference arm.
TAG=${1:-r01w}
O=gpurun_out
mkdir -p $O
run_bench() { name=$1; shift; timeout 900 python bench.py "$@" > $O/bench_${name}_$TAG.json 2> $O/bench_${name}_$TAG.err; python - <<PY
import json
try:
    d = json.loads([l for l in open("$O/bench_${name}_$TAG.json") if l.startswith("{")][-1])
    print("$name", {k: d[k] for k in ("value", "ms_per_step", "phases_ms")}, "kb", d["roofline_kbuild"]["achieved"], d["roofline_kbuild"]["frac"], "e2e", d["e2e"]["value"])
except Exception as e:
    print("bench $name failed", e); print(open("$O/bench_${name}_$TAG.err").read()[-3000:])
PY
}
run_bench c2_occ4 --workload c2 --steps 10 --warmup 3 --no-cpu --opt kbuild_occ=4
run_bench c2_occ3 --workload c2 --steps 10 --warmup 3 --no-cpu --opt kbuild_occ=3
run_bench c4_occ4 --workload c4 --steps 2 --warmup 3 --no-cpu --opt kbuild_occ=4
run_bench c4_occ3 --workload c4 --steps 2 --warmup 3 --no-cpu --opt kbuild_occ=3
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee $O/pytest_gpu_$TAG.log
echo "== bench default"; timeout 600 python bench.py > $O/bench_default_$TAG.json 2> $O/bench_default_$TAG.err; python - <<PY
import json
d = json.loads([l for l in open("$O/bench_default_$TAG.json") if l.startswith("{")][-1])
print({k: d[k] for k in ("value", "ms_per_step", "phases_ms", "cholesky_tflops")}, "kb", d["roofline_kbuild"]["frac"], "e2e", d["e2e"]["value"], d["cpu_baseline"]["value"])
PY
