"""Time the one-pass and the branch-compacted two-pass Polya-gamma kernels on the cfg3 shape (T=1e5, 200 neurons,
psi ~ N(-2, 1)), CUDA events on the launching stream, L2 flushed by the 160 MB operands themselves."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyglm_b200.kernels import CudaKernels, pad_ldn  # noqa: E402

K = CudaKernels()
T, n = 100000, 200
torch.manual_seed(0)
psi = torch.zeros(T, pad_ldn(n), dtype=torch.float64, device=K.device)
psi[:, :n] = torch.randn(T, n, dtype=torch.float64, device=K.device) - 2.0
om = torch.zeros_like(psi)
res = {}
for variant in ("1", "2"):
    os.environ["PYGLM_PG_VARIANT"] = variant
    for _ in range(3):
        K.pg_draw(psi, n, om, 1, 1, 0, 0, n)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(10):
        K.pg_draw(psi, n, om, 1, 2 + i, 0, 0, n)
    e1.record()
    torch.cuda.synchronize()
    res["variant_%s_ms" % variant] = e0.elapsed_time(e1) / 10
    res["checksum_%s" % variant] = float(om[:, :n].sum())
print(json.dumps(res))
