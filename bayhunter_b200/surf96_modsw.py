"""Drop-in for BayHunter's SurfDisp plugin (src/surf96_modsw.py:14-126), backed
by the CUDA dispersion kernel through the C ABI entry `bh_surfdisp96` (which has
the argument meaning of the Fortran `surfdisp96`, src/extensions/surfdisp96.f:55).

Same constructor, `set_modelparams`, `run_model(h, vp, vs, rho) -> (x, y)` and
`(nan, nan)` failure convention.  Instances hold no ctypes/CUDA handles, so they
pickle (BayHunter stores plugins in <station>_config.pkl, src/utils.py:102-153)
and the CUDA context is created lazily in whichever process first calls
run_model (fork hazard).
"""
import ctypes

import numpy as np

from . import _lib

_SURFTAGS = {           # ref -> (iwave, igr); iwave 1 Love / 2 Rayleigh, igr 0 phase / 1 group
    "rdispgr": (2, 1), "ldispgr": (1, 1), "rdispph": (2, 0), "ldispph": (1, 0)}


class SurfDisp(object):
    """Forward modeling of dispersion curves on the GPU (SURF96-equivalent)."""

    def __init__(self, obsx, ref):
        self.obsx = np.asarray(obsx, dtype=np.float64)
        self.kmax = self.obsx.size
        self.ref = ref
        self.modelparams = {"mode": 1, "flsph": 0}
        self.wavetype, self.veltype = self.get_surftags(ref)
        if self.kmax > _lib.MAX_PERIODS:
            # the reference resamples to 60 periods and interpolates back
            # (src/surf96_modsw.py:35-43)
            self.obsx_int = np.linspace(self.obsx.min(), self.obsx.max(), _lib.MAX_PERIODS)

    def set_modelparams(self, **mparams):
        self.modelparams.update(mparams)

    def get_surftags(self, ref):
        try:
            return _SURFTAGS[ref]
        except KeyError:
            raise ReferenceError(
                "Reference %r is not available in SurfDisp; available refs are "
                "rdispgr, ldispgr, rdispph, ldispph (r=rayleigh, l=love, gr=group, ph=phase)" % (ref,))

    def run_model(self, h, vp, vs, rho, **params):
        lib = _lib.require_device()
        arrs = [np.ascontiguousarray(a, dtype=np.float32) for a in (h, vp, vs, rho)]  # f2py's REAL*4 cast
        nlayer = arrs[0].size
        if not all(a.size == nlayer for a in arrs):
            raise ValueError("h, vp, vs, rho must have equal length")
        pers = self.obsx_int if self.kmax > _lib.MAX_PERIODS else self.obsx
        pers = np.ascontiguousarray(pers, dtype=np.float64)
        kmax = pers.size
        dispvel = np.zeros(kmax)
        err = ctypes.c_int(0)
        _lib.check(lib.bh_surfdisp96(
            *[a.ctypes.data_as(_lib.c_float_p) for a in arrs], nlayer,
            int(self.modelparams["flsph"]), self.wavetype, int(self.modelparams["mode"]),
            self.veltype, kmax, pers.ctypes.data_as(_lib.c_double_p),
            dispvel.ctypes.data_as(_lib.c_double_p), ctypes.byref(err)))
        if err.value != 0:
            return np.nan, np.nan
        if self.kmax > _lib.MAX_PERIODS:
            return self.obsx, np.interp(self.obsx, pers, dispvel)
        return pers, dispvel
