#!/bin/bash
# round-2 first probe: FP64 tensor vs scalar rates, per-family segment timings, baseline bench, ncu of the pr / hurs quantile map
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/r2_host.txt; free -g >> gpurun_out/r2_host.txt
./scripts/micro/dmma_peak > gpurun_out/r2_dmma_peak.txt 2>&1
CELLS=4096 timeout 600 python scripts/bench_families.py > gpurun_out/r2_families_base.txt 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_base.json 2> gpurun_out/r2_bench_base.err
CELLS=1024 VARS=pr,hurs timeout 300 python scripts/bench_families.py > gpurun_out/r2_fam_plain.txt 2>&1 &&
CELLS=1024 VARS=pr,hurs timeout 900 ncu --set full --clock-control none --import-source on -k regex:qmap_kernel -c 2 \
    -o gpurun_out/r2_prof_qmap_prhurs python scripts/bench_families.py > gpurun_out/r2_ncu_qmap.log 2>&1
tail -3 gpurun_out/r2_ncu_qmap.log
cat gpurun_out/r2_dmma_peak.txt gpurun_out/r2_families_base.txt gpurun_out/r2_bench_base.json
