"""Display helper behind `model.plot()` (pyglm/plotting.py:3-113): weights image, adjacency image, spikes and firing
rates of the first few neurons.  Host-only convenience outside the Gibbs hot path (out of scope as a
feature); it exists so that the drop-in call does not fail.  matplotlib is imported on first use -- when it is not
installed the call raises an ImportError that says so.  Returns (fig, axs, handles); passing `handles` back updates
the artists in place, as the reference's example loops do (examples/synthetic.py:62-80)."""
import numpy as np


def _pyplot():
    try:
        import matplotlib.pyplot as plt
    except ImportError as exc:
        raise ImportError("model.plot() needs matplotlib, which is not installed; the samplers do not") from exc
    return plt


def _matrix_panel(plt, fig, ax, M, N, title, **imshow_kw):
    h = ax.imshow(M, interpolation="nearest", **imshow_kw)
    ax.set(xlabel="pre", ylabel="post", title=title, xticks=np.arange(N), yticks=np.arange(N))
    ax.set_xticklabels(np.arange(N) + 1)
    ax.set_yticklabels(np.arange(N) + 1)
    fig.colorbar(h, ax=ax, fraction=0.046, pad=0.04)
    return h


def plot_glm(data, weights, adjacency, firingrates, std_firingrates=None, fig=None, axs=None, handles=None,
             title=None, figsize=(6, 3), W_lim=3, pltslice=slice(0, 500), data_index=0, N_to_plot=2):
    plt = _pyplot()
    W, A = np.asarray(weights), np.asarray(adjacency)
    N = W.shape[0]
    shown = min(N, N_to_plot)
    ts = np.arange(pltslice.start, pltslice.stop)
    if handles is not None:
        handles[0].set_data(W[:, :, 0])
        handles[1].set_data(A)
        for n in range(shown):
            handles[2 + n].set_data(ts, firingrates[pltslice, n])
        if title is not None:
            handles[-1].set_text(title)
        plt.pause(0.001)
        return fig, axs, handles

    fig = plt.figure(figsize=figsize)
    grid = fig.add_gridspec(N_to_plot, 3)
    W_ax, A_ax = fig.add_subplot(grid[:, 0]), fig.add_subplot(grid[:, 1])
    lam_axs = [fig.add_subplot(grid[i, 2]) for i in range(N_to_plot)]
    handles = [_matrix_panel(plt, fig, W_ax, W[:, :, 0], N, "Weights", vmin=-W_lim, vmax=W_lim, cmap="RdBu"),
               _matrix_panel(plt, fig, A_ax, A, N, "Adjacency", vmin=0, vmax=1, cmap="Greys")]
    for n in range(shown):
        ax = lam_axs[n]
        spike_bins = np.flatnonzero(data[pltslice, n])
        ax.plot(spike_bins, np.ones_like(spike_bins), "ko", markersize=4)
        if std_firingrates is not None:
            sausage_plot(ts, firingrates[pltslice, n], std_firingrates[pltslice, n], sgax=ax, alpha=0.5)
        handles.append(ax.plot(firingrates[pltslice, n])[0])
        ax.set_ylim(-0.05, 1.1)
        ax.set_ylabel("$\\lambda_{%d}(t)$" % (n + 1))
    if shown:
        lam_axs[0].set_title("Firing Rates")
        lam_axs[shown - 1].set_xlabel("Time")
    if title is not None:
        handles.append(fig.suptitle(title))
    fig.tight_layout()
    return fig, (W_ax, A_ax, lam_axs), handles


def sausage_plot(x, y, yerr, sgax=None, **kwargs):
    """Shaded band y +- yerr (pyglm/plotting.py:116-137)."""
    plt = _pyplot()
    x, y, yerr = np.asarray(x), np.asarray(y), np.asarray(yerr)
    assert x.shape == y.shape == yerr.shape and x.ndim == 1
    ax = plt.gca() if sgax is None else sgax
    return ax.fill_between(x, y - yerr, y + yerr, **kwargs)
