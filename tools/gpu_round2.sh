#!/bin/bash
# One gpurun call (round 2): GPU tests, full bench, reference arm, ncu launch list, full captures of the scan kernel at B = 64 / 256 / 1024
# and of the pooling kernel.   usage: tools/gpu_round2.sh [tag]
TAG=${1:-r5}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt; nproc >> gpurun_out/gpu.txt
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_${TAG}.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_${TAG}.log
tail -4 gpurun_out/pytest_${TAG}.log
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/${TAG}_bench.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err; echo "reference arm rc=$?"
SMALL="python bench.py --steps 3 --warmup 3 --extras none --capacity-rows 0 --no-sweep --no-cpu-baseline"
timeout 300 $SMALL > gpurun_out/plain_${TAG}.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv $SMALL > gpurun_out/ncu_l_${TAG}.log 2>&1
echo "ncu list rc=$?"
for B in 64 256 1024; do
  timeout 300 $SMALL --batch $B > gpurun_out/plain_${TAG}_$B.log 2>&1 &&
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:flat_scan_umma -s 6 -c 1 -f -o gpurun_out/${TAG}_scan_b$B $SMALL --batch $B > gpurun_out/ncu_f_${TAG}_$B.log 2>&1
  echo "ncu full B=$B rc=$?"
done
timeout 300 python tools/prof_pool.py 256 512 768 float16 full > gpurun_out/${TAG}_pool.log 2>&1 &&
timeout 600 ncu --set full --import-source on --clock-control none -k regex:pool_norm_cluster -s 3 -c 1 -f -o gpurun_out/${TAG}_pool python tools/prof_pool.py 256 512 768 float16 full > /dev/null 2>&1
echo "ncu pool rc=$?"; cat gpurun_out/${TAG}_pool.log
