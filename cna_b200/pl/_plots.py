"""Plotting helpers (reference: ``src/cna/plotting/_umap.py``, ``_strat.py``)."""
import numpy as np


def _need(module, what):
    try:
        return __import__(module, fromlist=["_"])
    except ImportError as exc:  # the plotting stack is optional
        raise ImportError(f"cna_b200.pl.{what} needs the optional dependency '{module.split('.')[0]}'") from exc


def umap_overlay(data, mask, key, scatter0={}, scatter1={}, ax=None, noframe=True):
    """``_umap.py:17-36``: all cells in the background, the cells selected by the boolean ``mask`` on top,
    coloured by ``data.obs[key]`` on a symmetric diverging scale.  Returns the axes."""
    plt = _need("matplotlib.pyplot", "umap_overlay")
    sc = _need("scanpy", "umap_overlay")
    if ax is None:
        ax = plt.gca()
    shown = data.obs[mask][key]
    top = float(np.abs(shown).max()) if len(shown) > 0 else None
    background = {"alpha": 0.8, "s": 2, **scatter0}
    foreground = {"alpha": 0.9, "s": 8, "cmap": "seismic", "vmin": -top if top is not None else 0,
                  "vmax": top if top is not None else 1, **scatter1}
    sc.pl.umap(data, ax=ax, show=False, **background)
    sc.pl.umap(data[mask], color=key, ax=ax, show=False, title="", **foreground)
    return ax


def umap_ncorr(data, fdr_thresh=None, key="coef", **kwargs):
    """``_umap.py:6-15``: overlay of the neighbourhood coefficients that pass ``fdr_thresh`` (default 0.1)
    on the UMAP."""
    if fdr_thresh is None:
        fdr_thresh = 0.1
    passed = data.obs[f"{key}_fdr"] <= fdr_thresh
    if len(passed) == 0:
        print("no neighborhoods were significant at FDR <", fdr_thresh)
    umap_overlay(data, passed, key, **kwargs)


def violinplot(data, stratification, key="coef", ax=None, cmap="seismic", **kwargs):
    """``_strat.py:10-44``: violins of ``data.obs[key]`` per level of ``data.obs[stratification]`` (e.g. a
    clustering), each filled with a vertical colour gradient; extra keyword arguments go to
    ``Axes.violinplot``.  Returns the axes."""
    plt = _need("matplotlib.pyplot", "violinplot")
    path_mod, patches = _need("matplotlib.path", "violinplot"), _need("matplotlib.patches", "violinplot")
    if ax is None:
        ax = plt.gca()
    options = {"widths": 0.9, "showmeans": False, "showextrema": False, "showmedians": False, **kwargs}
    levels = data.obs[stratification].unique()
    groups = [data.obs.loc[data.obs[stratification] == v, key] for v in levels]
    parts = ax.violinplot(groups, np.arange(len(levels)), **options)
    (y0, y1), (x0, x1) = ax.get_ylim(), ax.get_xlim()
    gradient = np.linspace(0, 1, 1000)[:, None]
    for body in parts["bodies"]:
        outline = patches.PathPatch(path_mod.Path(body.get_paths()[0].vertices), facecolor="none", edgecolor="none")
        ax.add_patch(outline)
        ax.imshow(gradient, origin="lower", extent=[x0, x1, y0, y1], aspect="auto", cmap=cmap, clip_path=outline)
    ax.set_xticks(np.arange(len(levels)))
    ax.set_xticklabels(levels)
    ax.set_xlabel(stratification)
    ax.set_ylabel("Neighborhood Coefficient")
    return ax
