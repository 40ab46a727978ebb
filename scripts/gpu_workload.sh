#!/bin/bash
# one workload of the bench on the GPU box after the parity tests that involve it: WORKLOAD=hurs|pr|sfcWind|..., KEXPR = pytest -k
mkdir -p gpurun_out
TAG=${TAG:-wl}
W=${WORKLOAD:-hurs}
timeout 500 python -m pytest tests -m gpu -q -x -p no:cacheprovider -k "${KEXPR:-$W or golden or families or bootstrap or rank}" > gpurun_out/pytest_gpu_$TAG.log 2>&1
echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu_$TAG.log
tail -4 gpurun_out/pytest_gpu_$TAG.log
timeout 400 python bench.py --workload $W --steps 3 --warmup 3 --no-cpu-baseline --no-workloads > gpurun_out/bench_${TAG}_$W.json 2> gpurun_out/bench_${TAG}_$W.err
tail -2 gpurun_out/bench_${TAG}_$W.err
python - <<PY
import json
d = json.load(open("gpurun_out/bench_${TAG}_$W.json"))
print("$W", "value", round(d["value"]), "ms", round(d["ms_per_step"], 3), {k: round(x["ms"], 3) for k, x in d["segments"].items()}, "e2e", round(d["e2e"]["value"]) if d.get("e2e") else None)
PY
