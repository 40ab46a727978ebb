#!/bin/bash
# bench A/B over full environment assignments: VARIANTS="A=1 ATTRICI_TMA_PREP=4 ..." (one assignment per run)
mkdir -p gpurun_out
TAG=${TAG:-abenv}
for v in ${VARIANTS:-A=1}; do
    env $v timeout ${BENCH_TIMEOUT:-200} python bench.py --steps 10 --warmup 3 --no-cpu-baseline ${BENCH_ARGS:---no-workloads} > gpurun_out/bench_${TAG}_$v.json 2> gpurun_out/bench_${TAG}_$v.err
    tail -2 gpurun_out/bench_${TAG}_$v.err
    python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_${TAG}_$v.json"))
    print("$v", "value", round(d["value"]), "ms", round(d["ms_per_step"], 3), {k: round(x["ms"], 3) for k, x in d["segments"].items()}, "e2e", round(d["e2e"]["value"]))
except Exception as e:
    print("$v", "failed", e)
PY
done
