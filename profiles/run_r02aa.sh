TAG=r02aa
( time timeout 900 python -m pytest tests -m gpu -x -q -k "overlapped or pipelined or device_moments" ) > gpurun_out/${TAG}_pytest_overlap.log 2>&1
grep -E "passed|failed|error" gpurun_out/${TAG}_pytest_overlap.log | tail -3
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_1gpu.json 2> gpurun_out/${TAG}_bench_1gpu.err
cut -c1-250 gpurun_out/${TAG}_bench_1gpu.json; tail -3 gpurun_out/${TAG}_bench_1gpu.err
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_1gpu_20.json 2> gpurun_out/${TAG}_bench_1gpu_20.err
cut -c1-250 gpurun_out/${TAG}_bench_1gpu_20.json
