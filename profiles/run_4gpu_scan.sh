#!/bin/bash
# 4-GPU box: cfg3 with the scan kernel chosen by default and with the 4-CTA triangular cluster kernel forced.
TAG=${1:-r02ah}
mkdir -p gpurun_out
run() {
  local name=$1; shift
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port $((29700 + RANDOM % 200)) \
      bench.py --gpus 4 --warmup 3 --steps 30 --no-cpu-baseline > gpurun_out/${TAG}_${name}.json 2> gpurun_out/${TAG}_${name}.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_${name}.json").read().strip().splitlines()[-1])
    print("${name}", "value %.1f e2e %.1f ms %.3f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]), {k: round(v,3) for k,v in d["kernels_ms"].items()})
except Exception as e:
    print("${name} FAILED", e); print(open("gpurun_out/${TAG}_${name}.err").read()[-1500:])
PY
}
run bench_4gpu_default
PYGLM_SS_VARIANT=84 run bench_4gpu_tri4
