#!/bin/bash
# Runs on the GPU box (via gpurun): the default bench line, then (optionally) ncu captures of a smaller tile.
mkdir -p gpurun_out
STEPS=${STEPS:-5}
timeout 1200 python bench.py --steps $STEPS --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?"; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
if [ "${NCU:-0}" = "1" ]; then
  CMD="python bench.py --cells 4096 --steps 1 --warmup 3 --no-cpu-baseline"
  $CMD > gpurun_out/plain.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
  echo "ncu list exit $?"
  $CMD > gpurun_out/plain2.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:'pass_kernel|qmap_kernel|prep_scale|fit_step' -s 60 -c 24 -o gpurun_out/prof_r1 $CMD > gpurun_out/ncu_full.log 2>&1
  echo "ncu full exit $?"
fi
