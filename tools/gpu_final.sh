#!/bin/bash
# final verification of a round: all GPU tests, the full bench line, the reference arm.   usage: tools/gpu_final.sh [tag]
TAG=${1:-final}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_${TAG}.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_${TAG}.log
tail -4 gpurun_out/pytest_${TAG}.log
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
tail -2 gpurun_out/${TAG}_bench.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err; echo "reference arm rc=$?"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
