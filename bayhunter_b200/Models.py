"""Model parametrisation: the input adapter of the hot path.

`Model.get_vp_vs_h` has the behaviour of BayHunter's (src/Models.py:16-52):
Voronoi nuclei [vs_0..vs_{k-1}, z_0..z_{k-1}] (NaN padded) -> interfaces at the
nucleus mid-points, h (last = 0), vp from vp/vs (crust) or the mantle rule.
`pack_models` is its batched form producing the engine's packed rows
(vs, vp/vs, z_top, h) -- the layout `bh_engine_eval` reads.
"""
import numpy as np


class Model(object):
    @staticmethod
    def split_modelparams(model):
        model = np.asarray(model)
        model = model[~np.isnan(model)]
        n = int(model.size / 2)
        return n, model[:n], model[-n:]

    @staticmethod
    def get_vp(vs, vpvs=1.73, mantle=(4.3, 1.8)):
        vp = vs * vpvs
        ind_m = np.where(vs >= mantle[0])[0]
        if len(ind_m):
            vp[ind_m[0]:] = vs[ind_m[0]:] * mantle[1]
        return vp

    @staticmethod
    def get_vp_vs_h(model, vpvs=1.73, mantle=None):
        n, vs, z_vnoi = Model.split_modelparams(model)
        z_disc = (z_vnoi[:n - 1] + z_vnoi[1:n]) / 2.0
        h_lay = z_disc - np.concatenate(([0], z_disc[:-1]))
        h = np.concatenate((h_lay, [0]))
        vp = Model.get_vp(vs, vpvs, mantle) if mantle is not None else vs * vpvs
        return vp, vs, h


def pack_layers(h, vp, vs, lmax=None):
    """One explicit (h, vp, vs) model -> packed rows [lmax, 4] = (vs, vp/vs, z_top, h).
    z_top is cumsum(h) shifted, exactly what RFminiModRF derives (src/rfmini_modrf.py:122-123).
    vp/vs is chosen so that the device's vs*(vp/vs) reproduces vp (checked)."""
    h = np.asarray(h, dtype=np.float64)
    vp = np.asarray(vp, dtype=np.float64)
    vs = np.asarray(vs, dtype=np.float64)
    n = h.size
    lmax = n if lmax is None else lmax
    rows = np.zeros((lmax, 4))
    ratio = vp / vs
    bad = vs * ratio != vp
    for _ in range(4):     # nudge the ratio by ulps until vs*ratio rounds to vp
        if not bad.any():
            break
        lo = np.nextafter(ratio, -np.inf)
        hi = np.nextafter(ratio, np.inf)
        ratio = np.where(bad & (vs * lo == vp), lo, np.where(bad & (vs * hi == vp), hi, ratio))
        bad = vs * ratio != vp
    rows[:n, 0] = vs
    rows[:n, 1] = ratio
    rows[:n, 2] = np.concatenate(([0.0], np.cumsum(h)[:-1]))
    rows[:n, 3] = h
    return rows


def pack_models(models, vpvs, mantle=None, lmax=None):
    """Batch of Voronoi models [B, 2*kmax] (NaN padded) + vpvs [B] -> (rows [B,lmax,4], nlay [B])."""
    models = np.atleast_2d(np.asarray(models, dtype=np.float64))
    vpvs = np.broadcast_to(np.asarray(vpvs, dtype=np.float64), (models.shape[0],))
    nl = [(~np.isnan(m)).sum() // 2 for m in models]
    lmax = max(nl) if lmax is None else lmax
    rows = np.zeros((models.shape[0], lmax, 4))
    for b, m in enumerate(models):
        vp, vs, h = Model.get_vp_vs_h(m, vpvs[b], mantle)
        rows[b] = pack_layers(h, vp, vs, lmax)
    return rows, np.asarray(nl, dtype=np.int32)
