#!/bin/bash
# GPU call e (1 GPU): K-build v2 + coalesced tcgen05 epilogue: parity tests, GEMM check, benches, ncu of the new kernels.
TAG=${1:-r01e}
O=gpurun_out
mkdir -p $O
echo "== test_tf32"; timeout 180 tools/test_tf32 2>&1 | tee $O/test_tf32_$TAG.log
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee $O/pytest_gpu_$TAG.log
run_bench() { name=$1; shift; timeout 900 python bench.py "$@" > $O/bench_${name}_$TAG.json 2> $O/bench_${name}_$TAG.err; python - <<PY
import json
try:
    d = json.loads([l for l in open("$O/bench_${name}_$TAG.json") if l.startswith("{")][-1])
    print("$name", {k: d[k] for k in ("value", "ms_per_step", "phases_ms", "cholesky_tflops")}, "roof", d["roofline"]["achieved"], d["roofline"]["frac"], "kb", d["roofline_kbuild"]["frac"], "e2e", d["e2e"]["value"], "err", (d["cpu_baseline"] or {}).get("max_rel_err_mean_vs_gpu"), (d["cpu_baseline"] or {}).get("max_rel_err_var_vs_gpu"))
except Exception as e:
    print("bench $name failed", e); print(open("$O/bench_${name}_$TAG.err").read()[-3000:])
PY
}
run_bench c2_fp64 --workload c2 --steps 10 --warmup 3
run_bench c2_fp64_kbv1 --workload c2 --steps 5 --warmup 3 --no-cpu --opt kbuild_v1=1
run_bench c2_tf32 --workload c2 --precision tf32 --steps 5 --warmup 3 --no-cpu
run_bench c4_tf32 --workload c4 --precision tf32 --steps 3 --warmup 3 --no-cpu
run_bench c4_tf32_nb8 --workload c4 --precision tf32 --steps 3 --warmup 3 --no-cpu --opt tf32_nb=8
run_bench c3_fp64 --workload c3 --steps 2 --warmup 3 --no-cpu
echo "== ncu kbuild_dmma"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:kbuild_dmma -c 2 -o $O/prof_kbuild2_$TAG -f python tools/prof_factorize.py 8192 1 > $O/ncu_kbuild2_$TAG.log 2>&1
echo "== ncu tf32 gemm"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32x3 -s 4 -c 2 -o $O/prof_tf32_$TAG -f python tools/prof_factorize.py 16384 1 ExpQuad tf32 > $O/ncu_tf32_$TAG.log 2>&1
ls -la $O | tail -8
