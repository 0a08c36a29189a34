TAG=r01m
mkdir -p gpurun_out
python -m pytest tests/test_multigpu.py -m gpu -q 2>&1 | tail -5 | tee gpurun_out/${TAG}_pytest_multigpu.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29602 \
    bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_2gpu.json 2> gpurun_out/${TAG}_bench_2gpu.err
cat gpurun_out/${TAG}_bench_2gpu.json; tail -2 gpurun_out/${TAG}_bench_2gpu.err
