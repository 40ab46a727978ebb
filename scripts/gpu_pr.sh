#!/bin/bash
# pr (Bernoulli-Gamma) on the GPU box: parity tests that involve pr, then the pr workload of the bench with and without
# the phase-ordered compacting quantile map
mkdir -p gpurun_out
TAG=${TAG:-pr}
timeout 420 python -m pytest tests -m gpu -q -x -p no:cacheprovider ${PYTEST_ARGS:--k "pr or golden or families or modes or bootstrap"} > gpurun_out/pytest_gpu_$TAG.log 2>&1
echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu_$TAG.log
tail -6 gpurun_out/pytest_gpu_$TAG.log
for v in A ${VARIANTS:-ATTRICI_NO_PR_FOLD}; do
    env $v=1 timeout 300 python bench.py --workload pr --steps 3 --warmup 3 --no-cpu-baseline --no-workloads > gpurun_out/bench_${TAG}_$v.json 2> gpurun_out/bench_${TAG}_$v.err
    tail -2 gpurun_out/bench_${TAG}_$v.err
    python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_${TAG}_$v.json"))
    print("$v", "value", round(d["value"]), "ms", round(d["ms_per_step"], 3), {k: round(x["ms"], 3) for k, x in d["segments"].items()}, "e2e", round(d["e2e"]["value"]) if d.get("e2e") else None)
except Exception as e:
    print("$v", "failed", e)
PY
done
