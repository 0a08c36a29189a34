#!/bin/bash
# 2-GPU checks (under gpurun --gpus 2): NCCL / peer-memory parity tests, then the cfg3 bench on 2 ranks with and without
# the peer-memory exchange.
TAG=${1:-r02n}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multigpu.py -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/${TAG}_pytest_multigpu.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 \
    bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_2gpu.json 2> gpurun_out/${TAG}_bench_2gpu.err
cut -c1-250 gpurun_out/${TAG}_bench_2gpu.json; grep -o '"kernels_ms": {[^}]*}' gpurun_out/${TAG}_bench_2gpu.json; grep -i "warn\|error" gpurun_out/${TAG}_bench_2gpu.err | head -5
PYGLM_PEER_EXCHANGE=0 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 \
    bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_2gpu_nccl.json 2> gpurun_out/${TAG}_bench_2gpu_nccl.err
cut -c1-250 gpurun_out/${TAG}_bench_2gpu_nccl.json; grep -o '"kernels_ms": {[^}]*}' gpurun_out/${TAG}_bench_2gpu_nccl.json
