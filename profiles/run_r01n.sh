TAG=r01n
mkdir -p gpurun_out
python examples/block_network.py 2>&1 | tail -6 | tee gpurun_out/${TAG}_example_block_network.log
python -m pytest tests -m gpu -q 2>&1 | tail -5 | tee gpurun_out/${TAG}_pytest_gpu.log
python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_1gpu.json 2> gpurun_out/${TAG}_bench_1gpu.err
cat gpurun_out/${TAG}_bench_1gpu.json
