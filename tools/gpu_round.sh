#!/bin/bash
# One gpurun call: GPU tests, bench, ncu launch list and one full capture of the scan kernel.
# usage: tools/gpu_round.sh [tag]
TAG=${1:-r1}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt; nproc >> gpurun_out/gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_${TAG}.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_${TAG}.log
tail -5 gpurun_out/pytest_${TAG}.log
timeout 600 python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/bench_${TAG}.json; tail -5 gpurun_out/bench_${TAG}.err
SMALL="python bench.py --steps 3 --warmup 3 --no-sweep --no-cpu-baseline"
timeout 300 $SMALL > gpurun_out/plain_${TAG}.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv $SMALL > gpurun_out/ncu_l_${TAG}.log 2>&1
echo "ncu list rc=$?"
timeout 300 $SMALL > gpurun_out/plain2_${TAG}.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:flat_scan -s 4 -c 2 -f -o gpurun_out/prof_${TAG} $SMALL > gpurun_out/ncu_f_${TAG}.log 2>&1
echo "ncu full rc=$?"
