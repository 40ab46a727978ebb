#!/bin/bash
# Runs on the GPU box (via gpurun): parity tests, smoke, the default bench line; everything lands in gpurun_out/.
mkdir -p gpurun_out
TAG=${TAG:-x}
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider ${PYTEST_ARGS:-} > gpurun_out/pytest_gpu_$TAG.log 2>&1
echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu_$TAG.log
tail -5 gpurun_out/pytest_gpu_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1
echo "smoke exit $?" | tee -a gpurun_out/smoke_$TAG.log
tail -3 gpurun_out/smoke_$TAG.log
timeout 900 python bench.py ${BENCH_ARGS:-} > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
echo "bench exit $?"; tail -3 gpurun_out/bench_$TAG.err; head -c 300 gpurun_out/bench_$TAG.json; echo
