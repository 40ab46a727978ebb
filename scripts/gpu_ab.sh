#!/bin/bash
# A/B on the GPU box: parity tests, then the bench with the default path and with env-selected variants.
# Every step runs under its own short timeout (a handshake bug must cost minutes, not the call's whole limit).
mkdir -p gpurun_out
TAG=${TAG:-ab}
if [ "${SKIP_TESTS:-0}" != "1" ]; then
timeout ${TEST_TIMEOUT:-420} python -m pytest tests -m gpu -q -x -p no:cacheprovider ${PYTEST_ARGS:-} > gpurun_out/pytest_gpu_$TAG.log 2>&1
echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu_$TAG.log
tail -8 gpurun_out/pytest_gpu_$TAG.log
fi
run() {  # name, env assignments...
    name=$1; shift
    env "$@" timeout ${BENCH_TIMEOUT:-200} python bench.py --steps 10 --warmup 3 --no-cpu-baseline ${BENCH_ARGS:---no-workloads} > gpurun_out/bench_${TAG}_$name.json 2> gpurun_out/bench_${TAG}_$name.err
    tail -2 gpurun_out/bench_${TAG}_$name.err
    python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_${TAG}_$name.json"))
    print("$name", "value", round(d["value"]), "ms", round(d["ms_per_step"], 3), {k: round(v["ms"], 3) for k, v in d["segments"].items()}, "e2e", round(d["e2e"]["value"]), d.get("fit_stats",{}).get("passes_per_cell_mean"))
except Exception as e:
    print("$name", "failed", e)
PY
}
run new A=1
for v in ${VARIANTS:-}; do run $v $v=1; done
