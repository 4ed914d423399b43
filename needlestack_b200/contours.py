"""The contour / power grid of plot_rob_nb (/root/reference/bin/plot_rob_nb.r:80-136), without the plotting.

The reference evaluates toQvalue(x, y) (:83-87) on up to 1000 + 2500 grid points per plotted variant, each with
n + 1 tail probabilities and a sort; here every batch of points is one `Caller.qvalue_grid` call (kernel
`k_qgrid`).  What comes back is what the reference would draw: the grid vectors, the q-value matrix and the
"by hands" contour lines (:123-136) for the levels 10, 30, 50, 70, 100.

Deviation: for depths above 1000 the reference samples the zoom search points with rbeta()/runif() (:96); R's
random stream cannot be reproduced, so they are drawn from a numpy Generator (or passed in as `zoom_points`).
Below 1000 the path is deterministic and identical.
"""
import numpy as np

QLEVELS = (10, 30, 50, 70, 100)
MAX_NB_GRID_PTS = 2500
MAX_QVALUE = 100
NB_PTS_ZOOM = 1000


def _rround(x):
    """R's round(): half to even, like numpy"""
    return np.round(x)


def zoom_points(max_dp, rng=None):
    """plot_rob_nb.r:95-99: candidate AO values at the largest depth"""
    if max_dp > NB_PTS_ZOOM:
        rng = rng or np.random.default_rng(0)
        a = _rround(max_dp * rng.beta(1, 100, NB_PTS_ZOOM))
        b = rng.uniform(1, max_dp, 100)
        return np.unique(np.concatenate([a, b]))
    return np.arange(1, max_dp + 1, dtype=float)


def grid_vectors(max_dp, ylim_zoom_cor):
    """plot_rob_nb.r:107-121: xgrid (DP) and ygrid (AO)"""
    if ylim_zoom_cor * max_dp <= MAX_NB_GRID_PTS:
        return np.arange(0, max_dp + 1, dtype=float), np.arange(0, np.floor(ylim_zoom_cor) + 1, dtype=float)
    if ylim_zoom_cor <= 50:
        ygrid = np.arange(0, np.floor(ylim_zoom_cor) + 1, dtype=float)
        xlength = int(_rround(MAX_NB_GRID_PTS / len(ygrid)))
        return _rround(np.linspace(0, max_dp, xlength)), ygrid
    ylength = 50
    xlength = int(_rround(MAX_NB_GRID_PTS / ylength))
    return _rround(np.linspace(0, max_dp, xlength)), _rround(np.linspace(0, ylim_zoom_cor, ylength))


def contour_lines(xgrid, ygrid, matgrid, qlevels=QLEVELS):
    """plot_rob_nb.r:123-136: for every level and every DP of xgrid the smallest AO whose q-value reaches the
    level (rows without one are dropped, except DP == 0 -> AO 0).  Returns {level: (DP[], AO[])}."""
    out = {}
    for qv in qlevels:
        dps, aos = [], []
        for ix, dp in enumerate(xgrid):
            row = matgrid[ix]
            sel = row >= qv
            if not sel.any():
                if dp == 0:
                    dps.append(dp)
                    aos.append(0.0)
                continue
            qval = row[sel].min()
            dps.append(dp)
            aos.append(ygrid[row == qval].min())
        out[qv] = (np.asarray(dps), np.asarray(aos))
    return out


def contour_grid(caller, reg_res, rng=None, zoom=None):
    """reg_res: the dict `Caller.glmrob_nb` returns (coverage, ma_count, coef).  Returns None when the
    reference draws nothing (`is.na(slope)`, :78), else dict(xgrid, ygrid, matgrid, af_min_lim, ylim_zoom,
    ylim_zoom_cor, lines)."""
    sigma, slope = reg_res["coef"]["sigma"], reg_res["coef"]["slope"]
    if not (slope == slope):
        return None
    cov = np.asarray(reg_res["coverage"], dtype=np.int64)
    ma = np.asarray(reg_res["ma_count"], dtype=np.int64)
    max_dp = int(cov.max())
    z = np.asarray(zoom if zoom is not None else zoom_points(max_dp, rng), dtype=float)
    zq = caller.qvalue_grid(sigma, slope, cov, ma, np.full(len(z), max_dp), z)
    hit = np.nonzero(zq >= QLEVELS[0])[0]
    af_min_lim = float(np.log10(z[hit[0]] / max_dp)) if len(hit) else 0.0
    hit = np.nonzero(zq >= MAX_QVALUE)[0]
    ylim_zoom = float(z[hit[0]]) if len(hit) else float("nan")
    ylim_zoom_cor = float(max_dp) if ylim_zoom != ylim_zoom else ylim_zoom
    xgrid, ygrid = grid_vectors(max_dp, ylim_zoom_cor)
    gx, gy = np.meshgrid(xgrid, ygrid, indexing="ij")
    matgrid = caller.qvalue_grid(sigma, slope, cov, ma, gx.ravel(), gy.ravel()).reshape(len(xgrid), len(ygrid))
    return dict(xgrid=xgrid, ygrid=ygrid, matgrid=matgrid, af_min_lim=af_min_lim, ylim_zoom=ylim_zoom,
                ylim_zoom_cor=ylim_zoom_cor, lines=contour_lines(xgrid, ygrid, matgrid))
