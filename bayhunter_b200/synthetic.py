"""Seeded synthetic layered-earth workloads for tests and bench.py.

Model draw mirrors SingleChain.draw_initmodel (src/SingleChain.py:94-123) with
the priors of tutorial/config.ini: k nuclei, vs ~ U(2,5) sorted, z ~ U(0,60)
sorted, vpvs ~ U(1.4,2.1), thickmin rejection.  Everything is numpy on the host;
the arrays are what a caller would hand to the engine.
"""
import numpy as np

from .Models import pack_layers

ST3 = dict(h=np.array([5., 23., 8., 0.]), vs=np.array([2.7, 3.6, 3.8, 4.4]), vpvs=1.73)


def draw_model(rng, k, thickmin=0.1):
    while True:
        vs = np.sort(rng.uniform(2.0, 5.0, k))
        z = np.sort(rng.uniform(0.0, 60.0, k))
        zd = (z[:-1] + z[1:]) / 2.0
        h = np.concatenate((zd - np.concatenate(([0.0], zd[:-1])), [0.0]))
        if k == 1 or h[:-1].min() > thickmin:
            return h, vs


def draw_batch(B, nrows, seed=20260101, lmax=None, thickmin=0.1):
    """B packed models.  nrows: int (fixed rows incl. half-space) or (lo, hi) inclusive."""
    rng = np.random.default_rng(seed)
    if np.isscalar(nrows):
        ks = np.full(B, int(nrows))
    else:
        ks = rng.integers(nrows[0], nrows[1] + 1, size=B)
    lmax = int(ks.max()) if lmax is None else lmax
    rows = np.zeros((B, lmax, 4))
    for b in range(B):
        h, vs = draw_model(rng, int(ks[b]), thickmin)
        vpvs = rng.uniform(1.4, 2.1)
        rows[b] = pack_layers(h, vs * vpvs, vs, lmax)
        rows[b, :int(ks[b]), 1] = vpvs     # exact ratio as the sampler holds it
    return rows, ks.astype(np.int32)


def perturb_batch(rows, nlay, rng, dvs=0.015, dz=0.015):
    """One MCMC-like perturbation of every model: vs += N(0,dvs), interfaces += N(0,dz)."""
    out = rows.copy()
    B, L, _ = rows.shape
    mask = np.arange(L)[None, :] < nlay[:, None]
    out[:, :, 0] = np.where(mask, rows[:, :, 0] + rng.normal(0, dvs, (B, L)), 0.0)
    h = np.where(mask, np.abs(rows[:, :, 3] + rng.normal(0, dz, (B, L))), 0.0)
    last = nlay - 1
    h[np.arange(B), last] = 0.0
    out[:, :, 3] = h
    z = np.cumsum(h, axis=1)
    out[:, 1:, 2] = np.where(mask[:, 1:], z[:, :-1], 0.0)
    out[:, 0, 2] = 0.0
    return out


def draw_noise(B, refs, seed=20260102, rf_corr=(0.35, 0.75)):
    """noise [B, 2T]: (corr, sigma) per target; SWD corr 0, RF corr ~ U(rf_corr)."""
    rng = np.random.default_rng(seed)
    cols = []
    for ref in refs:
        if ref in ("prf", "srf"):
            cols += [rng.uniform(rf_corr[0], rf_corr[1], B), rng.uniform(1e-5, 0.05, B)]
        else:
            cols += [np.zeros(B), rng.uniform(1e-5, 0.05, B)]
    return np.ascontiguousarray(np.stack(cols, axis=1))


def unpack(rows_b, n):
    """Packed rows of one model -> (h, vp, vs, rho) as the reference's evaluate sees them."""
    vs = rows_b[:n, 0].copy()
    vp = vs * rows_b[:n, 1]
    h = rows_b[:n, 3].copy()
    rho = vp * 0.32 + 0.77
    return h, vp, vs, rho


CONFIGS = {
    # BASELINE.json configs[1..4] (configs[0] is the CPU plumbing case)
    "swd2": dict(refs=("rdispph", "rdispgr"), periods=np.linspace(1, 40, 20), nrows=6, B=4096, rf=None),
    "joint5": dict(refs=("rdispph", "rdispgr", "ldispph", "ldispgr", "prf"),
                   periods=np.linspace(1, 40, 30), nrows=6, B=8192,
                   rf=dict(n=512, dt=0.1, t0=-5.0)),
    "transd3": dict(refs=("rdispph", "rdispgr", "prf"), periods=np.linspace(1, 40, 30),
                    nrows=(3, 31), B=4096, rf=dict(n=512, dt=0.1, t0=-5.0)),
}


def rf_time_axis(rf):
    return rf["t0"] + rf["dt"] * np.arange(rf["n"])
