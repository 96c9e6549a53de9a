#!/bin/bash
# usage: tools/prof_kernel.sh <kernel-regex> <tag> [bench args...]   -- one ncu --set full capture of 2 launches
K=$1; TAG=$2; shift 2
SMALL="python bench.py --steps 3 --warmup 3 --no-sweep --no-cpu-baseline $@"
timeout 300 $SMALL > gpurun_out/plain_${TAG}.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s 4 -c 2 -f -o gpurun_out/prof_${TAG} $SMALL > gpurun_out/ncu_f_${TAG}.log 2>&1
echo "ncu full rc=$?"
