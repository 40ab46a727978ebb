#!/bin/bash
# Runs on the GPU box (via gpurun): ncu launch list and one full capture of the chosen kernels, exported to
# CSV on the box (the .ncu-rep is kept only if small; gpurun_out/ is capped at 64 MiB).
mkdir -p gpurun_out
CELLS=${CELLS:-10000}
KREGEX=${KREGEX:-pass_kernel}
SKIP=${SKIP:-0}
COUNT=${COUNT:-3}
TAG=${TAG:-r1}
CMD="python bench.py --cells $CELLS --steps 1 --warmup 3 --no-cpu-baseline ${BENCH_ARGS:---no-workloads}"
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_list_$TAG.log 2>&1
echo "ncu list exit $?"
$CMD > gpurun_out/plain2_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"$KREGEX" -s $SKIP -c $COUNT -o /tmp/prof_$TAG $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
echo "ncu full exit $?"
ncu -i /tmp/prof_$TAG.ncu-rep --page raw --csv > gpurun_out/prof_${TAG}_raw.csv 2>/dev/null
ncu -i /tmp/prof_$TAG.ncu-rep --page details --csv > gpurun_out/prof_${TAG}_details.csv 2>/dev/null
if [ "${NOSRC:-0}" != "1" ]; then ncu -i /tmp/prof_$TAG.ncu-rep --page source --csv > gpurun_out/prof_${TAG}_source.csv 2>/dev/null; fi
SZ=$(stat -c %s /tmp/prof_$TAG.ncu-rep)
echo "rep size $SZ"
if [ "$SZ" -lt 30000000 ] && [ "${NOSRC:-0}" != "1" ]; then cp /tmp/prof_$TAG.ncu-rep gpurun_out/; fi
ls -la gpurun_out
