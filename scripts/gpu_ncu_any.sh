#!/bin/bash
# ncu --set full of chosen kernels of an arbitrary command (CMD), exported to CSV.  KREGEX, SKIP, COUNT, TAG as in gpu_ncu.sh
mkdir -p gpurun_out
TAG=${TAG:-any}
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"$KREGEX" -s ${SKIP:-0} -c ${COUNT:-1} -o /tmp/prof_$TAG $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
echo "ncu full exit $?"
ncu -i /tmp/prof_$TAG.ncu-rep --page raw --csv > gpurun_out/prof_${TAG}_raw.csv 2>/dev/null
ncu -i /tmp/prof_$TAG.ncu-rep --page details --csv > gpurun_out/prof_${TAG}_details.csv 2>/dev/null
ncu -i /tmp/prof_$TAG.ncu-rep --page source --csv > gpurun_out/prof_${TAG}_source.csv 2>/dev/null
