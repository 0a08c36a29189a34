#!/bin/bash
# Overlapped single-GPU sweep (two neuron groups on two streams, dynamic item queue in the Gram kernel): parity tests and
# A/B bench lines.   bash profiles/run_r02x.sh [tag]
TAG=${1:-r02x}
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q -k "overlapped or gram_tc or pipelined or peer_kernels or device_moments" ) > gpurun_out/${TAG}_pytest_overlap.log 2>&1
tail -6 gpurun_out/${TAG}_pytest_overlap.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_1gpu.json 2> gpurun_out/${TAG}_bench_1gpu.err
cut -c1-250 gpurun_out/${TAG}_bench_1gpu.json; tail -3 gpurun_out/${TAG}_bench_1gpu.err
PYGLM_OVERLAP=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_1gpu_no_overlap.json 2> gpurun_out/${TAG}_bench_1gpu_no_overlap.err
cut -c1-250 gpurun_out/${TAG}_bench_1gpu_no_overlap.json
PYGLM_OVERLAP=0 PYGLM_TC_DYNAMIC=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_1gpu_no_overlap_static.json 2> gpurun_out/${TAG}_bench_1gpu_no_overlap_static.err
cut -c1-250 gpurun_out/${TAG}_bench_1gpu_no_overlap_static.json
