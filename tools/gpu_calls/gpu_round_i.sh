#!/bin/bash
# GPU call i (1 GPU): tcgen05 GEMM with 128x256 CTA tiles (NB=2) vs 128x128; tf32 benches; parity.
TAG=${1:-r01i}
O=gpurun_out
mkdir -p $O
echo "== test_tf32"; timeout 240 tools/test_tf32 2>&1 | tee $O/test_tf32_$TAG.log
echo "== pytest -m gpu (tf32 + dist-free subset)"; timeout 1200 python -m pytest tests -m gpu -x -q -k "tf32 or golden or oracle" 2>&1 | tail -8 | tee $O/pytest_gpu_$TAG.log
run_bench() { name=$1; shift; timeout 900 python bench.py "$@" > $O/bench_${name}_$TAG.json 2> $O/bench_${name}_$TAG.err; python - <<PY
import json
try:
    d = json.loads([l for l in open("$O/bench_${name}_$TAG.json") if l.startswith("{")][-1])
    print("$name", {k: d[k] for k in ("value", "ms_per_step", "phases_ms", "cholesky_tflops")}, "roof", d["roofline"]["achieved"], d["roofline"]["frac"], "e2e", d["e2e"]["value"])
except Exception as e:
    print("bench $name failed", e); print(open("$O/bench_${name}_$TAG.err").read()[-3000:])
PY
}
run_bench c2_tf32 --workload c2 --precision tf32 --steps 5 --warmup 3 --no-cpu
run_bench c4_tf32 --workload c4 --precision tf32 --steps 3 --warmup 3 --no-cpu
run_bench c4_tf32_nb8 --workload c4 --precision tf32 --steps 3 --warmup 3 --no-cpu --opt tf32_nb=8
