#!/bin/bash
# cfg5 (N=1000, B=1, T=1e6, latent-distance prior) neuron-sharded over 8 ranks (under gpurun --gpus 8).
TAG=${1:-r02i}
mkdir -p gpurun_out
python -W always -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29655 \
    bench.py --gpus 8 --config cfg5 --steps 4 --warmup 3 --no-cpu-baseline \
    > gpurun_out/${TAG}_bench_cfg5_8gpu.json 2> gpurun_out/${TAG}_bench_cfg5_8gpu.err
cat gpurun_out/${TAG}_bench_cfg5_8gpu.json
grep -i "warn\|deviat\|error" gpurun_out/${TAG}_bench_cfg5_8gpu.err | head
tail -3 gpurun_out/${TAG}_bench_cfg5_8gpu.err
