"""oracle/joint_oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement of the whole hot path for one model at a time, exactly as the
reference strings it together:

  JointTarget.evaluate            src/Targets.py:314-347
  Valuation.get_covariance_*      src/Targets.py:105-173   (dense matrices, as the reference)
  SurfDisp.run_model              src/surf96_modsw.py:84-126   (-> surf96_oracle.c)
  RFminiModRF.compute_rf          src/rfmini_modrf.py:99-142   (-> oracle/_ref rfmini or rf_oracle.c)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs import this module.  The forward numerics live in liboracle.so (C
restatements) and, when present, oracle/_ref/librfmini_ref.so (the reference's
own C++ compiled from /root/reference by oracle/Makefile).
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_F = ctypes.POINTER(ctypes.c_float)
_D = ctypes.POINTER(ctypes.c_double)
_I = ctypes.POINTER(ctypes.c_int)

_state = {}


def build():
    """Compile liboracle.so (and oracle/_ref when /root/reference exists)."""
    subprocess.run(["make", "-s", "-C", HERE, "all"], check=True, capture_output=True)


def lib():
    if "lib" not in _state:
        path = os.path.join(HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        L = ctypes.CDLL(path)
        L.surf96_oracle.argtypes = [_F, _F, _F, _F] + [ctypes.c_int] * 6 + [_D, _D, _I, ctypes.POINTER(ctypes.c_long)]
        L.surf96_oracle.restype = None
        L.rf_oracle.argtypes = [ctypes.c_int] + [ctypes.c_double] * 6 + [ctypes.c_int] * 2 + [_D] * 7
        L.rf_oracle.restype = ctypes.c_int
        L.surf96_oracle_batch.argtypes = [_D, _I] + [ctypes.c_int] * 7 + [_D, _D, _I, ctypes.POINTER(ctypes.c_long)]
        L.surf96_oracle_batch.restype = None
        L.rf_oracle_batch.argtypes = [_D, _I, ctypes.c_int, ctypes.c_int, ctypes.c_int] + \
            [ctypes.c_double] * 5 + [ctypes.c_int, ctypes.c_int, _D, _D, _D]
        L.rf_oracle_batch.restype = None
        _state["lib"] = L
    return _state["lib"]


def ref_rfmini():
    """The reference's own rfmini (oracle/_ref/librfmini_ref.so) or None."""
    if "ref" not in _state:
        path = os.path.join(HERE, "_ref", "librfmini_ref.so")
        R = None
        if os.path.exists(path):
            R = ctypes.CDLL(path)
            R.synrf_cwrap.argtypes = [ctypes.c_int] + [ctypes.c_double] * 6 + [ctypes.c_int] * 2 + [_D] * 9
            R.synrf_cwrap.restype = ctypes.c_int
        _state["ref"] = R
    return _state["ref"]


def ref_surf96():
    """The reference's own surfdisp96.f (oracle/_ref/libsurf96_ref.so, gfortran symbol surfdisp96_) or None.
    It exists only where oracle/Makefile found a Fortran compiler next to the reference tree."""
    if "ref_surf96" not in _state:
        path = os.path.join(HERE, "_ref", "libsurf96_ref.so")
        R = None
        if os.path.exists(path):
            R = ctypes.CDLL(path)
            R.surfdisp96_.argtypes = [_F] * 4 + [_I] * 6 + [_D, _D, _I]
            R.surfdisp96_.restype = None
        _state["ref_surf96"] = R
    return _state["ref_surf96"]


def surfdisp_fortran(h, vp, vs, rho, ref, periods, mode=1, flsph=0):
    """The same call as surfdisp() through the compiled reference Fortran (None when it is not built)."""
    R = ref_surf96()
    if R is None:
        return None
    iwave, igr = SURFTAGS[ref]
    arrs = [np.zeros(100, dtype=np.float32) for _ in range(4)]
    for a, v in zip(arrs, (h, vp, vs, rho)):
        a[:len(v)] = v
    t = np.zeros(60); t[:len(periods)] = periods
    cg = np.zeros(60)
    ints = [ctypes.c_int(v) for v in (len(h), int(flsph), iwave, int(mode), igr, len(periods))]
    err = ctypes.c_int(0)
    R.surfdisp96_(*[a.ctypes.data_as(_F) for a in arrs], *[ctypes.byref(i) for i in ints],
                  t.ctypes.data_as(_D), cg.ctypes.data_as(_D), ctypes.byref(err))
    if err.value != 0:
        return np.nan, np.nan
    return np.asarray(periods, dtype=np.float64), cg[:len(periods)].copy()


SURFTAGS = {"rdispgr": (2, 1), "ldispgr": (1, 1), "rdispph": (2, 0), "ldispph": (1, 0)}


def surfdisp(h, vp, vs, rho, ref, periods, count=None, mode=1, flsph=0):
    """SurfDisp.run_model (src/surf96_modsw.py:84-126): (x, y) or (nan, nan)."""
    periods = np.ascontiguousarray(periods, dtype=np.float64)
    iwave, igr = SURFTAGS[ref]
    arrs = [np.ascontiguousarray(a, dtype=np.float32) for a in (h, vp, vs, rho)]   # f2py cast
    kall = periods.size
    pers = np.linspace(periods.min(), periods.max(), 60) if kall > 60 else periods
    pers = np.ascontiguousarray(pers)
    cg = np.zeros(pers.size)
    err = ctypes.c_int(0)
    ns = (ctypes.c_long * 2)()
    lib().surf96_oracle(*[a.ctypes.data_as(_F) for a in arrs], arrs[0].size, int(flsph), iwave, int(mode), igr,
                        pers.size, pers.ctypes.data_as(_D), cg.ctypes.data_as(_D), ctypes.byref(err), ns)
    if count is not None:
        count[0] += ns[0]
        count[1] += ns[1]
    if err.value != 0:
        return np.nan, np.nan
    if kall > 60:
        return periods, np.interp(periods, pers, cg)
    return pers, cg


def rf_params(obsx):
    """RFminiModRF._init_obsparams (src/rfmini_modrf.py:41-62)."""
    deltas = np.round(np.diff(obsx), 4)
    assert np.unique(deltas).size == 1
    fsamp = 1.0 / float(deltas[0])
    tshft = -float(obsx[0])
    nsamp = int(2. ** int(np.ceil(np.log2(obsx.size * 2))))
    return fsamp, tshft, nsamp


def recfunc(h, vp, vs, rho, obsx, wtype="P", gauss=1.0, p=6.4, nsv=None, qp=None, qs=None,
            use_reference=True):
    """RFminiModRF.compute_rf (src/rfmini_modrf.py:99-142)."""
    h, vp, vs, rho = [np.ascontiguousarray(a, dtype=np.float64) for a in (h, vp, vs, rho)]
    fsamp, tshft, nsamp = rf_params(obsx)
    n = h.size
    qp = np.ones(n) * 500. if qp is None else np.ascontiguousarray(qp, dtype=np.float64)
    qs = np.ones(n) * 225. if qs is None else np.ascontiguousarray(qs, dtype=np.float64)
    z = np.cumsum(h)
    z = np.ascontiguousarray(np.concatenate(([0], z[:-1])))
    vpvs = float(vp[0]) / float(vs[0])
    poisson = (2 - vpvs ** 2) / (2 - 2 * vpvs ** 2)
    if nsv is None:
        nsv = float(vs[0])
    waveno = ["P", "SV"].index(wtype)
    rf = np.zeros(nsamp)
    R = ref_rfmini() if use_reference else None
    ptr = lambda a: a.ctypes.data_as(_D)
    if R is not None and n >= 2:
        fz = np.zeros(nsamp)
        fr = np.zeros(nsamp)
        R.synrf_cwrap(nsamp, fsamp, tshft, p, gauss, nsv, poisson, waveno, n,
                      ptr(z), ptr(vp), ptr(vs), ptr(rho), ptr(qp), ptr(qs), ptr(fz), ptr(fr), ptr(rf))
    else:
        lib().rf_oracle(nsamp, fsamp, tshft, p, gauss, nsv, poisson, waveno, n,
                        ptr(z), ptr(vp), ptr(vs), ptr(rho), ptr(qp), ptr(qs), ptr(rf))
    time = np.arange(nsamp) / fsamp - tshft
    return time[:obsx.size], rf[:obsx.size]


# ---- likelihood: literal restatement of Valuation (dense, like the reference) ----
def cov_nocorr(sigma, size, yerr=None, corr=0):                       # Targets.py:105-115
    return np.diag(np.ones(size)) / (sigma ** 2), (2 * size) * np.log(sigma)


def cov_nocorr_scalederr(sigma, size, yerr, corr=0):                  # Targets.py:117-129
    scaled_err = yerr / yerr.min()
    c_inv = np.diag(np.ones(size)) / (scaled_err * sigma ** 2)
    return c_inv, (2 * size) * np.log(sigma) + np.log(np.prod(scaled_err))


def corr_inv_exp(corr, size):                                         # Targets.py:131-137
    d = np.ones(size) + corr ** 2
    d[0] = d[-1] = 1
    e = np.ones(size - 1) * -corr
    return np.diag(d) + np.diag(e, k=1) + np.diag(e, k=-1)


def cov_exp(corr, sigma, size, yerr=None):                            # Targets.py:139-148
    c_inv = corr_inv_exp(corr, size) / (sigma ** 2 * (1 - corr ** 2))
    return c_inv, (2 * size) * np.log(sigma) + (size - 1) * np.log(1 - corr ** 2)


def gauss_init(corr, size, rcond=None):                               # Targets.py:150-160
    idx = np.fromfunction(lambda i, j: (abs((i + j) - 2 * i)), (size, size))
    rmatrix = corr ** (idx ** 2)
    corr_inv = np.linalg.pinv(rmatrix, rcond=rcond) if rcond is not None else np.linalg.inv(rmatrix)
    _, logdet = np.linalg.slogdet(rmatrix)
    return corr_inv, logdet


def cov_gauss(sigma, size, corr_inv, logcorr_det):                    # Targets.py:162-173
    return corr_inv / (sigma ** 2), (2 * size) * np.log(sigma) + logcorr_det


class OracleTarget(object):
    def __init__(self, ref, x, y, cov="exp", yerr=None, corr_inv=None, logcorr_det=0.0, **params):
        self.ref, self.x, self.y, self.cov = ref, np.asarray(x, float), np.asarray(y, float), cov
        self.yerr, self.corr_inv, self.logcorr_det = yerr, corr_inv, logcorr_det
        self.params = params

    def forward(self, h, vp, vs, rho, count=None):
        if self.ref in SURFTAGS:
            return surfdisp(h, vp, vs, rho, self.ref, self.x, count=count,
                            mode=self.params.get("mode", 1), flsph=self.params.get("flsph", 0))
        kw = {k: v for k, v in self.params.items() if k in ("gauss", "p", "nsv", "use_reference")}
        return recfunc(h, vp, vs, rho, self.x, wtype="SV" if self.ref == "srf" else "P", **kw)

    def covariance(self, corr, sigma):
        n = self.y.size
        if self.cov == "exp":
            return cov_exp(corr, sigma, n)
        if self.cov == "white":
            return cov_nocorr(sigma, n)
        if self.cov == "white_scaled":
            return cov_nocorr_scalederr(sigma, n, self.yerr)
        return cov_gauss(sigma, n, self.corr_inv, self.logcorr_det)


def evaluate(targets, h, vp, vs, noise, rho=None, count=None):
    """JointTarget.evaluate (src/Targets.py:314-347) -> (logL, misfits[T+1], valid, synth list)."""
    if rho is None:
        rho = vp * 0.32 + 0.77
    T = len(targets)
    logL = 0
    misfits = []
    synth = []
    for n, t in enumerate(targets):
        xm, ym = t.forward(h, vp, vs, rho, count=count)
        valid = isinstance(xm, np.ndarray) and len(xm) == len(t.x) and np.sum(t.x - xm) <= 1e-5
        if not valid:
            return -1e15, np.array([1e15] * (T + 1)), False, synth
        synth.append(ym)
        misfits.append(np.sqrt(np.mean((ym - t.y) ** 2)))
        corr, sigma = noise[2 * n:2 * n + 2]
        c_inv, logc_det = t.covariance(corr, sigma)
        ydiff = ym - t.y
        madist = (ydiff.T).dot(c_inv).dot(ydiff)
        logL_part = -0.5 * (t.y.size * np.log(2 * np.pi) + logc_det)
        logL += (logL_part - madist / 2.)
    return logL, np.concatenate((misfits, [np.sum(misfits)])), True, synth


def evaluate_batch(targets, rows, nlay, noise, count=None):
    """Loop of `evaluate` over packed rows (vs, vp/vs, z_top, h)."""
    B = rows.shape[0]
    T = len(targets)
    logL = np.zeros(B)
    misfits = np.zeros((B, T + 1))
    status = np.zeros(B, dtype=np.int32)
    synth = np.full((B, sum(t.y.size for t in targets)), np.nan)
    for b in range(B):
        n = int(nlay[b])
        vs = rows[b, :n, 0].copy()
        vp = vs * rows[b, :n, 1]
        h = rows[b, :n, 3].copy()
        l, m, ok, sy = evaluate(targets, h, vp, vs, noise[b], count=count)
        logL[b], misfits[b], status[b] = l, m, int(ok)
        o = 0
        for t, s in zip(targets, sy):
            synth[b, o:o + t.y.size] = s
            o += t.y.size
    return logL, misfits, status, synth
