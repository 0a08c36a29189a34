"""One launch of the streaming tensor-core Gram at a named shape (for ncu captures):
    ncu --set full -k regex:gram_tc python profiles/probe_one_gram.py cfg4r [resident]"""
import sys

import torch

sys.path.insert(0, ".")
from pyglm_b200.kernels import CudaKernels, pad_ldn  # noqa: E402
from pyglm_b200.utils.basis import cosine_basis  # noqa: E402
from profiles.probe_gram_stream import SHAPES  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
stream = not (len(sys.argv) > 2 and sys.argv[2] == "resident")
sh = SHAPES[name]
N, B, T, n = sh["N"], sh["B"], sh["T"], sh["n"]
D = N * B + 1
K = CudaKernels()
g = torch.Generator(device=K.device)
g.manual_seed(0)
Y = (torch.rand(T, N, generator=g, device=K.device, dtype=torch.float32) < 0.05).to(torch.float64)
Xp = K.filter_spikes(Y, K.to_device(cosine_basis(B=B, L=100) / 100), True)
del Y
om = K.zeros(T, pad_ldn(n))
om[:, :n] = 0.02 + 0.4 * torch.rand(T, n, generator=g, device=K.device, dtype=torch.float64) ** 3
plan = K.gram_tc_plan(Xp, D, n, 4, stream=stream)
plan.slice_omega(om)
for _ in range(2):
    plan.mma()
torch.cuda.synchronize()
print("ok", name, "stream" if stream else "resident")
