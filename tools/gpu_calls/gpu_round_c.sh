#!/bin/bash
# GPU call c (1 GPU): tcgen05 GEMM re-check with banded raster, all parity tests (incl. tf32 mode), benches fp64/tf32.
TAG=${1:-r01c}
O=gpurun_out
mkdir -p $O
echo "== test_tf32"; timeout 180 tools/test_tf32 2>&1 | tee $O/test_tf32_$TAG.log
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee $O/pytest_gpu_$TAG.log
for cfg in "c2 fp64" "c2 tf32" "c4 tf32"; do set -- $cfg; echo "== bench $1 $2"; timeout 900 python bench.py --workload $1 --precision $2 --steps 5 --warmup 3 > $O/bench_$1_$2_$TAG.json 2> $O/bench_$1_$2_$TAG.err; python - <<PY
import json
try:
    d = json.load(open("$O/bench_$1_$2_$TAG.json"))
    print({k: d[k] for k in ("value", "ms_per_step", "phases_ms", "cholesky_tflops")}, d["roofline"]["achieved"], d["roofline"]["frac"], d["e2e"], d["cpu_baseline"])
except Exception as e:
    print("bench failed", e); print(open("$O/bench_$1_$2_$TAG.err").read()[-3000:])
PY
done
