#!/bin/bash
# Strong-scaling run of the bench on one box (under gpurun --gpus 8): N = 8, 4, 2 ranks.  Outputs -> gpurun_out/.
TAG=${1:-r01}
mkdir -p gpurun_out
for n in 8 4 2; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + n)) \
      bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_${n}gpu.json 2> gpurun_out/${TAG}_bench_${n}gpu.err
  cat gpurun_out/${TAG}_bench_${n}gpu.json
  tail -2 gpurun_out/${TAG}_bench_${n}gpu.err
done
python -m pytest tests/test_multigpu.py -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/${TAG}_pytest_multigpu.log
