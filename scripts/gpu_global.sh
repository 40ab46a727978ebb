#!/bin/bash
# Runs on the GPU box: BASELINE configs[4] (scripts/run_global.py) on N GPUs + the box's PCIe ceiling + the bench at N.
# N=${N:-1}; extra arguments of run_global.py in GLOBAL_ARGS
mkdir -p gpurun_out
N=${N:-1}
TAG=${TAG:-r2_global_${N}gpu}
LAUNCH="python"
if [ "$N" -gt 1 ]; then LAUNCH="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"; fi
nproc > gpurun_out/${TAG}_host.txt; free -g >> gpurun_out/${TAG}_host.txt; df -h /tmp >> gpurun_out/${TAG}_host.txt
nvidia-smi topo -m >> gpurun_out/${TAG}_host.txt 2>&1
timeout 300 $LAUNCH scripts/micro/pcie_ranks.py > gpurun_out/${TAG}_pcie.json 2> gpurun_out/${TAG}_pcie.err
cat gpurun_out/${TAG}_pcie.json
rm -f gpurun_out/${TAG}.jsonl
timeout ${GLOBAL_TIMEOUT:-900} $LAUNCH scripts/run_global.py --json gpurun_out/${TAG}.jsonl ${GLOBAL_ARGS:-} > gpurun_out/${TAG}.log 2> gpurun_out/${TAG}.err
echo "global exit $?"; tail -3 gpurun_out/${TAG}.err; cut -c1-600 gpurun_out/${TAG}.log
if [ "${BENCH:-1}" = "1" ]; then
timeout 600 $LAUNCH bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline ${BENCH_ARGS:-} > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench exit $?"; tail -3 gpurun_out/${TAG}_bench.err; cut -c1-1500 gpurun_out/${TAG}_bench.json
fi
