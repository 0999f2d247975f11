#!/bin/bash
# GPU call u (1 GPU): potrf 3rd-order rsqrt + K-build plain mode: clocks, parity, default bench.
TAG=${1:-r01u}
O=gpurun_out
mkdir -p $O
echo "== micro_potrf"; timeout 60 tools/micro_potrf 2>&1 | tee $O/micro_potrf_$TAG.log
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee $O/pytest_gpu_$TAG.log
echo "== bench default"; timeout 600 python bench.py > $O/bench_default_$TAG.json 2> $O/bench_default_$TAG.err
python - <<PY
import json
d = json.loads([l for l in open("$O/bench_default_$TAG.json") if l.startswith("{")][-1])
print({k: d[k] for k in ("value", "ms_per_step", "phases_ms", "cholesky_tflops")}, "kb", d["roofline_kbuild"]["frac"], "e2e", d["e2e"]["value"], d["cpu_baseline"]["max_rel_err_mean_vs_gpu"], d["cpu_baseline"]["max_rel_err_var_vs_gpu"])
PY
echo "== bench reference"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref_$TAG.json 2>/dev/null; cut -c1-300 $O/bench_ref_$TAG.json
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-200
