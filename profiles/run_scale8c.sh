#!/bin/bash
# 8-GPU box (gpurun --gpus 8) at the end of round 2: cfg3 on 8 ranks (default time-sharded; neuron-sharded), on 4 ranks, cfg4 on 8.
TAG=${1:-r02af}
mkdir -p gpurun_out
run() {
  local name=$1 n=$2; shift 2
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29700 + RANDOM % 200)) \
      bench.py --gpus $n --warmup 3 --no-cpu-baseline "$@" > gpurun_out/${TAG}_${name}.json 2> gpurun_out/${TAG}_${name}.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_${name}.json").read().strip().splitlines()[-1])
    print("${name}", "value %.1f e2e %.1f ms %.3f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]), {k: round(v,3) for k,v in d["kernels_ms"].items()}, d["clocks"])
except Exception as e:
    print("${name} FAILED", e); print(open("gpurun_out/${TAG}_${name}.err").read()[-1500:])
PY
}
run bench_8gpu 8 --steps 20
run bench_8gpu_neuron 8 --steps 20 --shard neuron
run bench_4gpu 4 --steps 20
run bench_cfg4_8gpu 8 --config cfg4 --steps 5
