( time timeout 900 python -m pytest tests -m gpu -x -q -k "overlapped or activation or pipelined or loglik or means" ) > gpurun_out/r02y_pytest_overlap.log 2>&1
tail -4 gpurun_out/r02y_pytest_overlap.log
python profiles/probe_overlap_timeline.py > gpurun_out/r02y_timeline.log 2>&1; tail -42 gpurun_out/r02y_timeline.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02y_bench_1gpu.json 2> gpurun_out/r02y_bench_1gpu.err
cut -c1-250 gpurun_out/r02y_bench_1gpu.json; tail -3 gpurun_out/r02y_bench_1gpu.err
