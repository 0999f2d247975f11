#!/bin/bash
# GPU call f (1 GPU): regression after the peer-push / K-build specialisation changes; K-build timing; e2e phase breakdown.
TAG=${1:-r01f}
O=gpurun_out
mkdir -p $O
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee $O/pytest_gpu_$TAG.log
run_bench() { name=$1; shift; timeout 900 python bench.py "$@" > $O/bench_${name}_$TAG.json 2> $O/bench_${name}_$TAG.err; python - <<PY
import json
try:
    d = json.loads([l for l in open("$O/bench_${name}_$TAG.json") if l.startswith("{")][-1])
    print("$name", {k: d[k] for k in ("value", "ms_per_step", "phases_ms", "cholesky_tflops")}, "roof", d["roofline"]["achieved"], d["roofline"]["frac"], "kb", d["roofline_kbuild"]["frac"], "e2e", d["e2e"], "err", (d["cpu_baseline"] or {}).get("max_rel_err_mean_vs_gpu"), (d["cpu_baseline"] or {}).get("max_rel_err_var_vs_gpu"))
except Exception as e:
    print("bench $name failed", e); print(open("$O/bench_${name}_$TAG.err").read()[-3000:])
PY
}
run_bench c2_fp64 --workload c2 --steps 10 --warmup 3
run_bench c4_matern_fp64 --workload c4 --steps 2 --warmup 3 --no-cpu
run_bench c4_tf32 --workload c4 --precision tf32 --steps 3 --warmup 3 --no-cpu
echo "== ncu kbuild_dmma"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:kbuild_dmma -c 2 -o $O/prof_kbuild3_$TAG -f python tools/prof_factorize.py 8192 1 > $O/ncu_kbuild3_$TAG.log 2>&1
echo "== ncu dram traffic of the predict solve (dgemm launches of one predict)"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:dgemm_nt --csv --log-file $O/solve_traffic_$TAG.csv python tools/prof_factorize.py 8192 1 > /dev/null 2>&1
python - <<PY
import csv, io
rows = list(csv.DictReader(io.StringIO("".join(l for l in open("$O/solve_traffic_$TAG.csv") if not l.startswith("==")))))
per = {}
for r in rows:
    per.setdefault(r["ID"], {})[r["Metric Name"]] = (float(r["Metric Value"].replace(",", "")), r["Metric Unit"])
ids = sorted(per, key=int)
print("dgemm launches profiled:", len(ids))
# the last 129 launches are the predict solve (2*65-1)
def tob(v, u):
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
sel = ids[-129:]
tot = sum(tob(*per[i]["dram__bytes_read.sum"]) + tob(*per[i]["dram__bytes_write.sum"]) for i in sel)
print("predict-solve launches:", len(sel), "dram bytes total %.3e  per launch %.3e" % (tot, tot / len(sel)))
PY
