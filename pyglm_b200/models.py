"""Model container and Gibbs driver: the drop-in surface of pyglm/models.py.

    SparseBernoulliGLM(N, basis=...)   add_data(Y)   generate(T=...)   resample_model()   log_likelihood()
    weights / adjacency / biases / means / regressions / network / data_list

The public attributes and call signatures are the reference's (models.py:28-236); what changes is where the
work happens: add_data filters on the GPU, resample_model runs one GibbsEngine.sweep for all N regressions at
once (the reference loops over them, models.py:169-171), and log_likelihood / means are single fused kernels.
Extra keyword-only arguments (`seed`, `shard`, `gram` / `precision`, `device`) default to reference behaviour.
"""
import numpy as np

from . import networks as _networks
from . import regression as _regression


class DeviceDesign(object):
    """Stand-in for the host copy of X in `data_list` when add_data(..., host_X=False): the (T, N, B) regressors
    live in HBM only and are copied to the host on demand (np.asarray(x), x[...])."""

    def __init__(self, dataset, N, B):
        self._ds, self._N, self._B = dataset, N, B
        self.shape = (dataset.T, N, B)
        self.ndim = 3
        self.dtype = np.dtype(np.float64)

    def __array__(self, dtype=None, copy=None):
        from .engine import default_kernels
        X = default_kernels().unpack_design(self._ds.Xp, self._N * self._B).cpu().numpy().reshape(self.shape)
        return X if dtype is None else X.astype(dtype)

    def __getitem__(self, idx):
        return np.asarray(self)[idx]


class NonlinearAutoregressiveModel(object):
    """The neuroscience "GLM": a nonlinear vector autoregression, one regression per observed dimension
    (models.py:8-201)."""

    def __init__(self, N, regressions, basis=None, B=10, seed=None, shard="neuron", comm=None, gram="auto",
                 gram_stream="auto", precision=None, device=None):
        # precision: SURVEY 5's name for the Gram arithmetic -- "fp64" (FP64 DMMA kernel), "int8" (tcgen05 integer-digit
        # kernel, exact integer sums, <= 1e-9 of FP64) or "auto"; an alias of `gram`.  device: CUDA device (index or
        # torch.device) of this process' engine; default: the current device (one process per GPU).
        if precision is not None:
            assert gram == "auto", "give either gram= or precision=, not both"
            gram = {"fp64": "fp64", "int8": "tc", "tc": "tc", "auto": "auto"}[precision]
        self._device = device
        self.N = N
        assert len(regressions) == N
        self.regressions = regressions
        if basis is None:
            basis = np.eye(B)
        else:
            assert basis.ndim == 2
        self.basis = basis
        self.B = self.basis.shape[1]
        self.data_list = []
        # device side (created lazily so that constructing / inspecting a model needs no GPU)
        self._seed = int(np.random.randint(2 ** 31 - 1)) if seed is None else int(seed)
        self._shard = shard
        self._comm = comm
        self._gram = gram
        self._gram_stream = gram_stream
        self._engine = None
        self._dev = {}           # (id(X), id(Y)) -> (X, Y, DeviceDataset): keeps the keys alive

    # ------------------------------------------------------------------ state views (models.py:54-64)
    @property
    def weights(self):
        return np.array([r.W for r in self.regressions])

    @property
    def adjacency(self):
        return np.array([r.a for r in self.regressions])

    @property
    def biases(self):
        return np.array([r.b for r in self.regressions]).ravel()

    # ------------------------------------------------------------------ device plumbing
    @property
    def engine(self):
        if self._engine is None:
            from .engine import GibbsEngine, default_kernels
            kernels = None
            if self._device is not None:
                import torch
                dev = self._device if isinstance(self._device, torch.device) else torch.device("cuda", int(self._device))
                kernels = default_kernels(dev)
            self._engine = GibbsEngine(self.N, self.B, kernels=kernels, seed=self._seed, comm=self._comm,
                                       shard=self._shard, gram=self._gram, gram_stream=self._gram_stream)
        return self._engine

    def _sync_ranks(self):
        """Multi-GPU runs: every rank continues from RANK 0's randomness and state -- the engine's Philox seed, numpy's
        global RNG state (the host network step and the eta draws use it), the regressions' (a, W, b, eta) and the
        network's latent state -- so that generate(), the first partial Grams of a time-sharded sweep and every host
        draw agree across ranks whatever each rank's own seeding was.  Collective; runs once, before the first sweep
        or simulation.  The chain is the one a single process with rank 0's seeds would produce."""
        if getattr(self, "_ranks_synced", False):
            return
        self._ranks_synced = True
        eng = self.engine
        comm = eng.comm
        if comm.world == 1:
            return
        regs = self.regressions
        mine = None
        if comm.rank == 0:
            mine = dict(seed=eng.seed, rng=np.random.get_state(), A=self.adjacency, W=self.weights, b=self.biases,
                        eta=[getattr(r, "eta", None) for r in regs],
                        net=self.network.get_state() if hasattr(getattr(self, "network", None), "get_state") else None)
        st = comm.broadcast_object(mine)
        eng.seed = st["seed"]
        np.random.set_state(st["rng"])
        for n, r in enumerate(regs):
            r.a, r.W, r.b = st["A"][n].copy(), st["W"][n].copy(), st["b"][n:n + 1].copy()
            if st["eta"][n] is not None:
                r.eta = st["eta"][n]
        if st["net"] is not None:
            self.network.set_state(st["net"])
        self._state_src = None

    def _time_slab(self, T):
        eng = self.engine
        if eng.shard == "time" and eng.comm.world > 1:
            from .distributed import time_partition
            return time_partition(T, eng.comm.world, eng.comm.rank)
        return 0, T

    def _device_dataset(self, X, Y):
        """Device copy of one data_list entry, uploaded on first use and cached by object identity (users may
        replace data_list entries, test/test_generate.py:27)."""
        key = (id(X), id(Y))
        hit = self._dev.get(key)
        if hit is not None and hit[0] is X and hit[1] is Y:
            return hit[2]
        if isinstance(X, DeviceDesign):
            ds = X._ds
        else:
            lo, hi = self._time_slab(Y.shape[0])
            Xs = np.reshape(X, (Y.shape[0], -1))[lo:hi]
            ds = self.engine.make_dataset(Y[lo:hi], X=Xs, t_off=lo, T_global=Y.shape[0])
        self._dev[key] = (X, Y, ds)
        return ds

    def _device_datasets(self):
        live = set()
        out = []
        for X, Y in self.data_list:
            out.append(self._device_dataset(X, Y))
            live.add((id(X), id(Y)))
        for key in [k for k in self._dev if k not in live]:
            del self._dev[key]
        return out

    @property
    def _gaussian(self):
        """True when the regressions have Gaussian observations (regression.py:380-456)."""
        return isinstance(self.regressions[0], _regression.SparseGaussianRegression)

    def _etas(self):
        return np.array([float(r.eta) for r in self.regressions])

    def _host_state(self):
        """(A, W, b) stacked over the regressions.  After a sweep the regressions hold row VIEWS of the stacked arrays
        the engine returned, so as long as nobody rebound regressions[n].a / .W / .b the stacked arrays themselves are
        the state (in-place edits through the views included) and need not be rebuilt."""
        src = getattr(self, "_state_src", None)
        if src is not None:
            A, W, b, views = src
            if all(r._a is v[0] and r._W is v[1] and r._b is v[2] for r, v in zip(self.regressions, views)):
                return A, W, b
        return self.adjacency, self.weights, self.biases

    # ------------------------------------------------------------------ checkpoint / resume (optional)
    def state_dict(self):
        """Everything a chain needs to continue bit for bit: the sampler state (A, W, b, eta), the hyper-parameters the
        regressions currently hold, the network's latent state, the engine's Philox seed and sweep counter and numpy's
        global RNG state (the host network step draws from it).  O(N^2 B) numbers; the data are not included.  The
        reference has no checkpointing (its state is the same plain attributes)."""
        eng = self.engine
        A, W, b = self._host_state()
        regs = self.regressions
        net = getattr(self, "network", None)
        hyp = self._stacked_hypers()
        return dict(version=1, N=self.N, B=self.B, seed=int(eng.seed), calls=int(eng.calls),
                    rng=np.random.get_state(), A=np.array(A), W=np.array(W), b=np.array(b),
                    eta=[getattr(r, "eta", None) for r in regs],
                    hypers={k: np.array(v) for k, v in hyp.items()},
                    network=net.get_state() if hasattr(net, "get_state") else None)

    def load_state_dict(self, sd):
        """Continue the chain a state_dict() was taken from: same model structure (N, B, network class) and the same
        data added in the same order.  Restores numpy's GLOBAL RNG state."""
        assert sd["version"] == 1 and sd["N"] == self.N and sd["B"] == self.B, "state_dict of a different model"
        eng = self.engine
        eng.seed, eng.calls = int(sd["seed"]), int(sd["calls"])
        eng._pending = None                         # a pre-launched augmentation belongs to the old sweep counter
        for n, r in enumerate(self.regressions):
            r.a, r.W, r.b = sd["A"][n].copy(), sd["W"][n].copy(), sd["b"][n:n + 1].copy()
            if sd["eta"][n] is not None:
                r.eta = sd["eta"][n]
            h = sd["hypers"]
            r.rho, r.mu_w, r.S_w = h["rho"][n].copy(), h["mu_w"][n].copy(), h["S_w"][n].copy()
            r.mu_b, r.S_b = h["mu_b"][n], h["S_b"][n]
        self._state_src = self._hyper_src = None
        net = getattr(self, "network", None)
        if sd["network"] is not None and hasattr(net, "set_state"):
            net.set_state(sd["network"])
        self._ranks_synced = True                   # every rank loads the same dictionary
        np.random.set_state(sd["rng"])

    # ------------------------------------------------------------------ data (models.py:66-80)
    def add_data(self, data, X=None, host_X=True, slab=None):
        """Append a (T, N) spike matrix.  X (T, N, B) may be supplied; otherwise it is the causal convolution of
        `data` with the basis, computed on the GPU.  host_X=False keeps X in HBM only (data_list then holds a
        DeviceDesign handle): use it for recordings whose X should not be mirrored in host RAM.

        slab=(t0, T_total), time-sharded runs only: `data` holds the global bins [t0, t0 + len(data)) of a recording of
        T_total bins -- this rank's slab plus the L bins before it (the filter's history) -- instead of the whole
        recording, so that a 1e7-bin recording never has to exist in one process."""
        N, B = self.N, self.B
        assert isinstance(data, np.ndarray) and data.ndim == 2 and data.shape[1] == self.N
        t0, T = (0, data.shape[0]) if slab is None else (int(slab[0]), int(slab[1]))
        if X is None:
            lo, hi = self._time_slab(T)
            if host_X and (lo, hi) != (0, T):
                raise ValueError("time-sharded add_data needs host_X=False (each rank holds only its slab)")
            # filter halo: the L bins before the slab (zero history only at the true start)
            L = self.basis.shape[0]
            h0 = max(0, lo - L)
            if not (t0 <= h0 and hi <= t0 + data.shape[0]):
                raise ValueError("add_data(slab=...): bins [%d, %d) needed, [%d, %d) given"
                                 % (h0, hi, t0, t0 + data.shape[0]))
            ds = self.engine.make_dataset(data[h0 - t0:hi - t0], basis=self.basis, t_off=h0, T_global=T)
            if lo > h0:
                ds.Xp = ds.Xp[lo - h0:].contiguous()
                ds.Y = ds.Y[lo - h0:].contiguous()
                ds.T, ds.t_off = hi - lo, lo
            if host_X:
                X = np.asarray(DeviceDesign(ds, N, B))
            else:
                X = DeviceDesign(ds, N, B)
            self._dev[(id(X), id(data))] = (X, data, ds)
        else:
            assert slab is None and X.shape == (T, N, B)
        self.data_list.append((X, data))

    # ------------------------------------------------------------------ scoring (models.py:82-96)
    def log_likelihood(self, datas=None):
        A, W, b = self._host_state()
        if datas is None:
            dsets = self._device_datasets()
        else:
            dsets = []
            for data in datas:
                if isinstance(data, tuple):
                    X, Y = data
                    dsets.append(self._device_dataset(X, Y))
                else:
                    dsets.append(self.engine.make_dataset(data, basis=self.basis))
        if self._gaussian:
            # sum over bins of -1/2 log(2 pi eta) - 1/2 (y - psi)^2 / eta (regression.py:400-404)
            eta = self._etas()
            T = sum(ds.T for ds in dsets)
            rss = self.engine.residual_ss(dsets, A, W, b)
            return float(np.sum(-0.5 * T * np.log(2 * np.pi * eta) - 0.5 * rss / eta))
        return self.engine.log_likelihood(dsets, A, W, b)

    # ------------------------------------------------------------------ simulation (models.py:98-151)
    def generate(self, keep=True, T=100, verbose=False, intvl=10, return_uniforms=False):
        """Simulate T time bins forward from the model (models.py:98-151): x_t = window of the last L bins projected
        on the flipped basis, psi_t = W x_t + b, y_t ~ Bern(logistic(psi_t)), sequential in time by construction.

        Runs as one persistent thread-block cluster on the GPU (csrc/generate.cu).  W is `self.weights` as in the
        reference (models.py:124 -- not masked by the adjacency).  The spikes come from the model's Philox stream,
        not from numpy's global state, so they are statistically, not bitwise, the reference's; X equals
        convolve_with_basis(Y) exactly.  `verbose` / `intvl` are accepted for signature compatibility (there is no
        per-step host loop to report from).  With keep=True the simulated recording is appended to data_list and
        its device copy is reused as is (no re-upload, no re-filtering)."""
        if T == 0:
            return np.zeros((0, self.N))
        assert isinstance(T, int), "Size must be an integer number of time bins"
        N, basis = self.N, self.basis
        L, B = basis.shape
        assert not np.allclose(np.flipud(basis), self.basis)
        from .engine import DeviceDataset
        self._sync_ranks()
        eng = self.engine
        K = eng.K
        self._generated = getattr(self, "_generated", 0) + 1
        Xp, Yd, U = K.generate(K.to_device(self.weights.reshape((N, N * B))), K.to_device(self.biases),
                               K.to_device(np.ascontiguousarray(basis, dtype=np.float64)), T, eng.seed,
                               0x40000000 + self._generated, want_uniforms=return_uniforms,
                               # Gaussian observations: regressions[0].rvs draws for every neuron (models.py:146)
                               gauss_sd=float(np.sqrt(self.regressions[0].eta)) if self._gaussian else -1.0)
        Y = Yd.cpu().numpy()
        sharded_time = eng.shard == "time" and eng.comm.world > 1
        ds = DeviceDataset(Xp, Yd)
        X = np.asarray(DeviceDesign(ds, N, B))
        if keep:
            if sharded_time:
                self.add_data(Y, host_X=False)       # every rank simulated the same recording; keep the local slab
            else:
                self._dev[(id(X), id(Y))] = (X, Y, ds)
                self.data_list.append((X, Y))
        if return_uniforms:
            return X, Y, U.cpu().numpy()
        return X, Y

    # ------------------------------------------------------------------ rates (models.py:153-163)
    @property
    def means(self):
        A, W, b = self._host_state()
        if self._gaussian:
            return [self.engine.activations(ds, A, W, b) for ds in self._device_datasets()]
        return [self.engine.means(ds, A, W, b) for ds in self._device_datasets()]

    # ------------------------------------------------------------------ sample statistics on the device
    def start_collecting(self, rates=False):
        """From the next sweep on, accumulate running moments of the samples in HBM (engine.DeviceMoments): the state
        (adjacency, weights, biases) and, with rates=True, the firing rates logistic(psi) of every data set.  The
        reference's example loop collects the same things on the host, one (T, N) D2H copy per sweep
        (examples/synthetic.py:51-59)."""
        from .engine import DeviceMoments
        self.engine.moments = DeviceMoments(rates=rates)

    def stop_collecting(self):
        self.engine.moments = None

    def posterior_moments(self):
        """dict(n, A_mean (N,N), W_mean / W_var (N,N,B), b_mean / b_var (N,), rate_mean / rate_var: per data set the
        (T_local, n_local) block this rank computes -- all of it on a single GPU -- or None)."""
        mom = self.engine.moments
        self.engine._overlap_drain()         # the sums may have been updated on the overlapped sweep's side streams
        assert mom is not None and mom.n > 0, "call start_collecting() and run at least one sweep first"
        N, B, n = self.N, self.B, float(mom.n)
        m1 = (mom.s1 / n).cpu().numpy()
        m2 = (mom.s2 / n).cpu().numpy()
        var = np.maximum(m2 - m1 * m1, 0.0)
        out = dict(n=mom.n, A_mean=m1[:, :N], W_mean=m1[:, N:N + N * B].reshape(N, N, B),
                   W_var=var[:, N:N + N * B].reshape(N, N, B), b_mean=m1[:, N + N * B], b_var=var[:, N + N * B],
                   rate_mean=None, rate_var=None)
        if mom.rates:
            out["rate_mean"] = [(mom.r1[di] / n).cpu().numpy() for di in sorted(mom.r1)]
            out["rate_var"] = [np.maximum((mom.r2[di] / n).cpu().numpy() - rm * rm, 0.0)
                               for di, rm in zip(sorted(mom.r1), out["rate_mean"])]
        return out

    # ------------------------------------------------------------------ Gibbs sampling (models.py:166-171)
    def resample_model(self):
        self.resample_regressions()

    def _stacked_hypers(self):
        regs = self.regressions
        src = getattr(self, "_hyper_src", None)
        if src is not None:
            # fast path: the hyper-parameters are still the row views resample_network() handed out
            sigma_W, mu_W, rho, views = src
            if all(r._S_w is v[0] and r._mu_w is v[1] and r._rho is v[2] for r, v in zip(regs, views)):
                return dict(rho=rho, mu_w=mu_W, S_w=sigma_W, mu_b=np.array([r.mu_b[0] for r in regs]),
                            S_b=np.array([r.S_b[0, 0] for r in regs]))
        return dict(rho=np.stack([r.rho for r in regs]), mu_w=np.stack([r.mu_w for r in regs]),
                    S_w=np.stack([r.S_w for r in regs]), mu_b=np.array([r.mu_b[0] for r in regs]),
                    S_b=np.array([r.S_b[0, 0] for r in regs]))

    def resample_regressions(self):
        """All N regressions in one device sweep (they are conditionally independent given the data)."""
        self._sync_ranks()
        A, W, b = self._host_state()
        if self._gaussian:
            dsets = self._device_datasets()
            A, W, b, rss = self.engine.sweep_gaussian(dsets, A, W, b, self._stacked_hypers(), self._etas())
            # eta_n ~ InvGamma(a_0 + T/2, b_0 + RSS_n): the host step of regression.py:432-445
            T = sum(ds.T for ds in dsets)
            for n, reg in enumerate(self.regressions):
                reg.eta = float(_regression.sample_invgamma(reg.a_0 + T / 2.0, reg.b_0 + rss[n]))
        else:
            A, W, b = self.engine.sweep(self._device_datasets(), A, W, b, self._stacked_hypers())
        views = []
        for n, reg in enumerate(self.regressions):
            v = (A[n], W[n], b[n:n + 1])
            reg._a, reg._W, reg._b = v
            views.append(v)
        self._state_src = (A, W, b, views)

    # ------------------------------------------------------------------ plotting (models.py:174-201)
    def plot(self, fig=None, axs=None, handles=None, title=None, figsize=(6, 3), W_lim=3,
             pltslice=slice(0, 500), N_to_plot=2, data_index=0):
        from .plotting import plot_glm
        return plot_glm(self.data_list[data_index][1], self.weights, self.adjacency, self.means[0], fig=fig,
                        axs=axs, handles=handles, title=title, figsize=figsize, W_lim=W_lim, pltslice=pltslice,
                        N_to_plot=N_to_plot)


class HierarchicalNonlinearAutoregressiveModel(NonlinearAutoregressiveModel):
    """Network GLM: the regressions' priors are tied together by a network object (models.py:204-236)."""

    def __init__(self, N, network, regressions, basis=None, B=10, **kwargs):
        super(HierarchicalNonlinearAutoregressiveModel, self).__init__(N, regressions, basis=basis, B=B, **kwargs)
        self.network = network

    def resample_model(self):
        super(HierarchicalNonlinearAutoregressiveModel, self).resample_model()
        self.resample_network()

    def resample_network(self):
        """Host step (models.py:228-236).  sigma_W / mu_W / rho are built once, not once per neuron as the
        reference's property accesses do."""
        net = self.network
        # Every rank draws the (tiny) network step itself, from the SAME stream: _sync_ranks() handed rank 0's numpy
        # state and network state to all ranks once, after which identical inputs (the all-gathered A, W) give
        # identical draws on every rank -- no per-sweep host collective on the critical path between two scans.
        if self._engine is not None:
            self._sync_ranks()
        A, W, _ = self._host_state()                 # the stacked arrays of the last sweep, not rebuilt row by row
        net.resample((np.asarray(A, dtype=bool), np.asarray(W)))
        sigma_W, mu_W, rho = net.sigma_W, net.mu_W, net.rho
        N, B = self.N, self.B
        fast = (sigma_W.shape == (N, N, B, B) and mu_W.shape == (N, N, B) and rho.shape == (N, N)
                and all(hasattr(reg, "_S_w") for reg in self.regressions))
        views = []
        for n, reg in enumerate(self.regressions):
            if fast:
                # what the property setters do for full-shape arrays (pass through uncopied), minus 3N shape checks
                v = (sigma_W[n], mu_W[n], rho[n])
                reg._S_w, reg._mu_w, reg._rho = v
                views.append(v)
            else:
                reg.S_w = sigma_W[n]
                reg.mu_w = mu_W[n]
                reg.rho = rho[n]
        self._hyper_src = (sigma_W, mu_W, rho, views) if fast else None


GLM = NonlinearAutoregressiveModel
NetworkGLM = HierarchicalNonlinearAutoregressiveModel


class _DefaultMixin(object):
    _network_class = None
    _regression_class = None

    def __init__(self, N, B=10, basis=None, network=None, network_kwargs=None, regressions=None,
                 regression_kwargs=None, **kwargs):
        B = B if basis is None else basis.shape[1]
        if network is None:
            network_kwargs = dict() if network_kwargs is None else network_kwargs
            network = self._network_class(N, B, **network_kwargs)
        if regressions is None:
            regression_kwargs = dict() if regression_kwargs is None else regression_kwargs
            regressions = [self._regression_class(N, B, **regression_kwargs) for _ in range(N)]
        super(_DefaultMixin, self).__init__(N, network, regressions, B=B, basis=basis, **kwargs)


class GaussianGLM(_DefaultMixin, NetworkGLM):
    _network_class = _networks.NIWDenseNetwork
    _regression_class = _regression.GaussianRegression


class SparseGaussianGLM(_DefaultMixin, NetworkGLM):
    _network_class = _networks.NIWSparseNetwork
    _regression_class = _regression.SparseGaussianRegression


class BernoulliGLM(_DefaultMixin, NetworkGLM):
    _network_class = _networks.NIWDenseNetwork
    _regression_class = _regression.BernoulliRegression


class SparseBernoulliGLM(_DefaultMixin, NetworkGLM):
    _network_class = _networks.NIWSparseNetwork
    _regression_class = _regression.SparseBernoulliRegression
