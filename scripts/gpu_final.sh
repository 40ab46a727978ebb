#!/bin/bash
# Evidence run of a round on one B200 (via gpurun): parity tests, smoke, the default bench line, ncu launch list and
# full captures of the three kernels of the tas step and of the precipitation quantile map.  Everything lands in
# gpurun_out/; the summaries that are judged are copied to profiles/ afterwards (scripts/README.md).
mkdir -p gpurun_out
TAG=${TAG:-r2final}
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu_$TAG.txt 2>&1
timeout 600 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/pytest_gpu_$TAG.log 2>&1
echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu_$TAG.log
tail -4 gpurun_out/pytest_gpu_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1
echo "smoke exit $?" | tee -a gpurun_out/smoke_$TAG.log
timeout 900 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
echo "bench exit $?"; tail -2 gpurun_out/bench_$TAG.err; head -c 600 gpurun_out/bench_$TAG.json; echo
TAG=${TAG}_tas KREGEX='qmap_fold_bulk|prep_onepass_kernel|fit_fused' COUNT=3 timeout 600 bash scripts/gpu_ncu.sh > gpurun_out/ncu_tas_$TAG.log 2>&1
tail -3 gpurun_out/ncu_tas_$TAG.log
TAG=${TAG}_pr CELLS=2048 BENCH_ARGS='--workload pr --no-workloads' KREGEX='qmap_fold_pr|mt19937' COUNT=2 timeout 600 bash scripts/gpu_ncu.sh > gpurun_out/ncu_pr_$TAG.log 2>&1
tail -3 gpurun_out/ncu_pr_$TAG.log
