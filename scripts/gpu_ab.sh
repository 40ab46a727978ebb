#!/bin/bash
# A/B on the GPU box: parity tests, then the bench with and without the fused fit
mkdir -p gpurun_out
TAG=${TAG:-ab}
timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider ${PYTEST_ARGS:-} > gpurun_out/pytest_gpu_$TAG.log 2>&1
echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu_$TAG.log
tail -15 gpurun_out/pytest_gpu_$TAG.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}_new.json 2> gpurun_out/bench_${TAG}_new.err
tail -3 gpurun_out/bench_${TAG}_new.err
python - <<PY
import json
for n in ("new",):
    try:
        d = json.load(open("gpurun_out/bench_${TAG}_%s.json" % n))
        print(n, "value", round(d["value"]), "ms", round(d["ms_per_step"], 3), {k: round(v["ms"], 3) for k, v in d["segments"].items()}, "e2e", round(d["e2e"]["value"]), d["config"].get("passes_per_cell_mean"))
    except Exception as e:
        print(n, "failed", e)
PY
