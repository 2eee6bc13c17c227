"""Drop-in for BayHunter's RFminiModRF plugin (src/rfmini_modrf.py:14-154),
backed by the CUDA receiver-function kernels through `bh_synrf` (argument
meaning of `synrf_cwrap`, src/extensions/rfmini/wrap.cpp:57-80).

Same constructor, `set_modelparams` keys (gauss, p, water, nsv, wtype),
`run_model(h, vp, vs, rho, qp=, qs=) -> (time, rf)`; picklable, lazy CUDA init.
"""
import numpy as np

from . import _lib

_WAVENO = {"P": 0, "SV": 1}


class RFminiModRF(object):
    """Forward modeling of receiver functions on the GPU (rfmini-equivalent)."""

    def __init__(self, obsx, ref):
        self.ref = ref
        self.obsx = np.asarray(obsx, dtype=np.float64)
        self._init_obsparams()
        if self.ref in ("prf", "seis"):
            self.modelparams = {"wtype": "P"}
        elif self.ref in ("srf",):
            self.modelparams = {"wtype": "SV"}
        else:
            raise ReferenceError("Reference %r is not available in RFminiModRF (prf, srf)" % (ref,))
        # `water` is accepted for compatibility; the reference never forwards it
        # to the native code (rfmini.pyx:74-77) and neither do we.
        self.modelparams.update({"gauss": 1.0, "p": 6.4, "water": 0.001, "nsv": None})

    def _init_obsparams(self):
        """fsamp, tshft, nsamp from the observed time axis (src/rfmini_modrf.py:41-62)."""
        deltas = np.round(np.diff(self.obsx), 4)
        if np.unique(deltas).size != 1:
            raise ValueError("Target: %s. Sampling rate must be constant." % self.ref)
        self.fsamp = 1.0 / float(deltas[0])
        self.tshft = -float(self.obsx[0])
        self.nsamp = 2 ** int(np.ceil(np.log2(self.obsx.size * 2)))

    def set_modelparams(self, **mparams):
        self.modelparams.update(mparams)

    def write_startmodel(self, h, vp, vs, rho, modfile, **params):
        """ASCII model table like the reference's (src/rfmini_modrf.py:64-94)."""
        h = np.asarray(h, dtype=float)
        cols = {"z": np.concatenate(([0.0], np.cumsum(h)[:-1])), "vp": vp, "vs": vs, "rho": rho,
                "qp": params.get("qp", np.ones(h.size) * 500.0),
                "qs": params.get("qs", np.ones(h.size) * 225.0)}
        fmts = {"z": "%.2f", "vp": "%.4f", "vs": "%.4f", "rho": "%.4f", "qp": "%.1f", "qs": "%.1f"}
        keys = [k for k in ("z", "vp", "vs", "rho", "qp", "qs") if cols[k] is not None]
        with open(modfile, "w") as f:
            f.write("\t".join(keys) + "\n")
            for i in range(h.size):
                f.write("\t".join(fmts[k] % float(cols[k][i]) for k in keys) + "\n")

    def compute_rf(self, h, vp, vs, rho, **params):
        lib = _lib.require_device()
        n = h.size
        qp = np.ascontiguousarray(params.get("qp", np.ones(n) * 500.0), dtype=np.float64)
        qs = np.ascontiguousarray(params.get("qs", np.ones(n) * 225.0), dtype=np.float64)
        z = np.ascontiguousarray(np.concatenate(([0.0], np.cumsum(h)[:-1])))
        vpvs = float(vp[0]) / float(vs[0])
        poisson = (2 - vpvs ** 2) / (2 - 2 * vpvs ** 2)
        nsv = self.modelparams["nsv"]
        if nsv is None:
            nsv = float(vs[0])
        wtype = self.modelparams["wtype"]
        if wtype not in _WAVENO:
            raise ValueError("wave must be 'P' or 'SV', not %r" % (wtype,))
        rf = np.zeros(self.nsamp)
        ptr = lambda a: a.ctypes.data_as(_lib.c_double_p)
        _lib.check(lib.bh_synrf(
            int(self.nsamp), self.fsamp, self.tshft, float(self.modelparams["p"]),
            float(self.modelparams["gauss"]), float(nsv), poisson, _WAVENO[wtype], n,
            ptr(z), ptr(vp), ptr(vs), ptr(rho), ptr(qp), ptr(qs), None, None, ptr(rf)))
        time = np.arange(self.nsamp) / self.fsamp - self.tshft
        return time[:self.obsx.size], rf[:self.obsx.size]

    def run_model(self, h, vp, vs, rho, **params):
        h, vp, vs, rho = [np.ascontiguousarray(a, dtype=np.float64) for a in (h, vp, vs, rho)]
        assert h.size == vp.size == vs.size == rho.size
        return self.compute_rf(h, vp, vs, rho, **params)
