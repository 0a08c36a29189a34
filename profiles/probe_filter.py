import sys, time, numpy as np, torch
sys.path.insert(0, ".")
from pyglm_b200.kernels import CudaKernels
from pyglm_b200.utils.basis import cosine_basis
K = CudaKernels(torch.device("cuda", 0))
for (T, N, B, L) in [(100000, 200, 2, 100), (100000, 27, 3, 100), (1000000, 100, 1, 100)]:
    Y = (np.random.default_rng(0).random((T, N)) < 0.05).astype(np.float64)
    Yd, bd = K.to_device(Y), K.to_device(cosine_basis(B=B, L=L) / L)
    for _ in range(2): Xp = K.filter_spikes(Yd, bd, True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): Xp = K.filter_spikes(Yd, bd, True)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    byts = 8.0 * T * N + 8.0 * T * Xp.shape[1]
    print("filter T=%d N=%d B=%d: %.3f ms, %.0f GB/s algorithmic, %.2f TFMA/s" % (T, N, B, ms, byts / ms / 1e6, T * N * B * L / ms / 1e9))
