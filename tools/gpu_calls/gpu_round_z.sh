#!/bin/bash
# GPU call z (1 GPU): final state check of the driver entry points: smoke, default bench, reference arm.
TAG=${1:-r01z}
O=gpurun_out
mkdir -p $O
timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -1 | cut -c1-160
timeout 600 python bench.py > $O/bench_default_$TAG.json 2> $O/bench_default_$TAG.err; python - <<PY
import json
d = json.loads([l for l in open("$O/bench_default_$TAG.json") if l.startswith("{")][-1])
print({k: d[k] for k in ("value", "ms_per_step", "phases_ms")}, d["roofline"]["frac"], d["roofline_kbuild"], "e2e", d["e2e"]["value"], d["cpu_baseline"]["value"], d["clocks"])
PY
tail -3 $O/bench_default_$TAG.err
