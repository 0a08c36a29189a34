#!/bin/bash
# 2-GPU checks (under gpurun --gpus 2): NCCL / peer-memory parity tests, then the cfg3 bench on 2 ranks.
TAG=${1:-r02ae}
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_multigpu.py -m gpu -x -q ) 2>&1 | tail -15 | tee gpurun_out/${TAG}_pytest_multigpu.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 \
    bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_2gpu.json 2> gpurun_out/${TAG}_bench_2gpu.err
cut -c1-250 gpurun_out/${TAG}_bench_2gpu.json; grep -o '"kernels_ms": {[^}]*}' gpurun_out/${TAG}_bench_2gpu.json; grep -i "warn\|error" gpurun_out/${TAG}_bench_2gpu.err | head -5
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 \
    bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline --shard neuron > gpurun_out/${TAG}_bench_2gpu_neuron.json 2> gpurun_out/${TAG}_bench_2gpu_neuron.err
cut -c1-250 gpurun_out/${TAG}_bench_2gpu_neuron.json; grep -o '"kernels_ms": {[^}]*}' gpurun_out/${TAG}_bench_2gpu_neuron.json
