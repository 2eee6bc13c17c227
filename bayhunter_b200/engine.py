"""Batched forward-model + log-likelihood engine (host side of the C ABI).

`Engine` binds one joint target set (observed data, covariance laws, forward
parameters) to device-resident constants and evaluates B layered models per
call.  PyTorch is used only as the owner of device buffers and streams; all
arithmetic happens inside libbayhunter_b200.so.
"""
import ctypes

import numpy as np

from . import _lib

_COV_NAMES = {"exp": _lib.COV_EXP, "white": _lib.COV_WHITE,
              "white_scaled": _lib.COV_WHITE_SCALED, "gauss": _lib.COV_GAUSS}


class TargetSpec(object):
    """Plain description of one target for the engine (no torch, no ctypes).

    ref      BayHunter reference string ('rdispph', 'ldispgr', 'prf', ...)
    x, y     observed abscissa / data (float64, same length)
    cov      'exp' | 'white' | 'white_scaled' | 'gauss'   (src/SingleChain.py:159-205)
    yerr     needed by 'white_scaled'
    corr_inv, logcorr_det   R^-1 (n x n) and slogdet(R) for 'gauss' (host computed,
             src/Targets.py:150-160 -- never re-derived on the device)
    params   plugin parameters: mode, flsph (SWD); gauss, p, nsv, qp, qs (RF)
    """

    def __init__(self, ref, x, y, cov="exp", yerr=None, corr_inv=None, logcorr_det=0.0, generic=False, **params):
        # generic: the forward model is the caller's (a user plugin): likelihood-only engine (Engine.loglik_host)
        if ref not in _lib.REF_CODES and not generic:
            raise ReferenceError("unknown target ref %r" % (ref,))
        self.ref = ref
        self.generic = bool(generic)
        self.x = np.ascontiguousarray(x, dtype=np.float64)
        self.y = np.ascontiguousarray(y, dtype=np.float64)
        if self.x.shape != self.y.shape or self.x.ndim != 1:
            raise ValueError("x and y must be 1-D arrays of equal length")
        self.cov = cov
        self.yerr = None if yerr is None else np.ascontiguousarray(yerr, dtype=np.float64)
        self.corr_inv = None if corr_inv is None else np.ascontiguousarray(corr_inv, dtype=np.float64)
        self.logcorr_det = float(logcorr_det)
        self.params = dict(mode=1, flsph=0, gauss=1.0, p=6.4, nsv=None, qp=500.0, qs=225.0)
        self.params.update(params)

    @property
    def n(self):
        return self.x.size

    def to_struct(self):
        s = _lib.BhTarget()
        s.ref = _lib.REF_GENERIC if self.generic else _lib.REF_CODES[self.ref]
        s.n = self.n
        s.x = self.x.ctypes.data_as(_lib.c_double_p)
        s.y = self.y.ctypes.data_as(_lib.c_double_p)
        s.yerr = self.yerr.ctypes.data_as(_lib.c_double_p) if self.yerr is not None else None
        s.cov = _COV_NAMES[self.cov]
        s.corr_inv = self.corr_inv.ctypes.data_as(_lib.c_double_p) if self.corr_inv is not None else None
        s.logcorr_det = self.logcorr_det
        if self.generic:
            s.mode, s.flsph, s.gauss, s.p, s.nsv, s.qp, s.qs = 1, 0, 1.0, 0.0, -1.0, 0.0, 0.0
            return s
        s.mode = int(self.params["mode"])
        s.flsph = int(self.params["flsph"])
        s.gauss = float(self.params["gauss"])
        s.p = float(self.params["p"])
        nsv = self.params["nsv"]
        s.nsv = -1.0 if nsv is None else float(nsv)
        s.qp = float(self.params["qp"])
        s.qs = float(self.params["qs"])
        return s


def gauss_corr_inverse(corr, size, rcond=None):
    """R^-1 and slogdet(R) of the Gaussian correlation law, computed once on the
    host exactly like Valuation.init_covariance_gauss (src/Targets.py:150-160)."""
    idx = np.abs(np.subtract.outer(np.arange(size), np.arange(size))).astype(float)
    rmatrix = corr ** (idx ** 2)
    corr_inv = np.linalg.pinv(rmatrix, rcond=rcond) if rcond is not None else np.linalg.inv(rmatrix)
    _, logdet = np.linalg.slogdet(rmatrix)
    return corr_inv, logdet


class Engine(object):
    """Device engine for a joint target set."""

    def __init__(self, specs, max_batch, max_layers):
        self._lib = _lib.require_device()
        self.specs = list(specs)
        self.ntargets = len(self.specs)
        self.max_batch = int(max_batch)
        self.max_layers = int(max_layers)
        arr = (_lib.BhTarget * self.ntargets)(*[s.to_struct() for s in self.specs])
        handle = ctypes.c_void_p()
        _lib.check(self._lib.bh_engine_create(arr, self.ntargets, self.max_batch, self.max_layers,
                                              ctypes.byref(handle)))
        self._h = handle
        self.synth_stride = self._lib.bh_engine_synth_stride(self._h)
        self.synth_offsets = np.concatenate(([0], np.cumsum([s.n for s in self.specs])))

    def close(self):
        if getattr(self, "_h", None):
            self._lib.bh_engine_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set(self, **tunables):
        for k, v in tunables.items():
            _lib.check(self._lib.bh_engine_set(self._h, k.encode(), int(v)))

    def last_counts(self):
        """(consumed, evaluated) secular-function values of the last eval."""
        c = self.last_counters()
        return c[0], c[1]

    def last_counters(self):
        """All work counters (include/bayhunter_b200.h: bh_engine_last_counts)."""
        buf = (ctypes.c_longlong * _lib.NUM_COUNTERS)()
        _lib.check(self._lib.bh_engine_last_counts(self._h, buf))
        return [int(v) for v in buf]

    def last_kernel_ms(self):
        """{kernel name: device ms} of the last eval (needs set(profile=1))."""
        buf = (ctypes.c_float * len(_lib.KERNEL_NAMES))()
        _lib.check(self._lib.bh_engine_last_kernel_ms(self._h, buf))
        return {n: float(buf[i]) for i, n in enumerate(_lib.KERNEL_NAMES) if buf[i] >= 0}

    # ---- device tensors (torch) ------------------------------------------
    def eval(self, model, nlay, noise, rho=None, want_synth=False, out=None, stream=None):
        """model [B,L,4] f64 cuda (vs, vp/vs, z_top, h); nlay [B] i32; noise [B,2T] f64.
        Returns (logL [B], misfits [B,T+1], status [B] i32, synth [B,stride] | None).
        Work is enqueued on `stream` (default: torch's current stream)."""
        import torch
        B, L = int(model.shape[0]), int(model.shape[1])
        assert model.is_cuda and model.dtype == torch.float64 and model.is_contiguous()
        assert tuple(model.shape) == (B, L, 4)
        assert nlay.dtype == torch.int32 and nlay.is_contiguous() and nlay.numel() == B
        assert noise.dtype == torch.float64 and noise.is_contiguous() and noise.numel() == B * 2 * self.ntargets
        if rho is not None:
            assert rho.dtype == torch.float64 and rho.is_contiguous() and rho.numel() == B * L
        dev = model.device
        if out is None:
            logL = torch.empty(B, dtype=torch.float64, device=dev)
            misfits = torch.empty((B, self.ntargets + 1), dtype=torch.float64, device=dev)
            status = torch.empty(B, dtype=torch.int32, device=dev)
            synth = torch.empty((B, self.synth_stride), dtype=torch.float64, device=dev) if want_synth else None
        else:
            logL, misfits, status, synth = out
        st = stream if stream is not None else torch.cuda.current_stream(dev)
        _lib.check(self._lib.bh_engine_eval(
            self._h, model.data_ptr(), nlay.data_ptr(), noise.data_ptr(),
            rho.data_ptr() if rho is not None else None, B, L,
            logL.data_ptr(), misfits.data_ptr(), status.data_ptr(),
            synth.data_ptr() if synth is not None else None, st.cuda_stream))
        return logL, misfits, status, synth

    # ---- host arrays (numpy or pinned torch CPU tensors) --------------------
    def eval_host(self, model, nlay, noise, rho=None, want_synth=False, out=None):
        """Same evaluation with HOST buffers; copies both ways happen inside the call."""
        model = np.ascontiguousarray(model, dtype=np.float64)
        nlay = np.ascontiguousarray(nlay, dtype=np.int32)
        noise = np.ascontiguousarray(noise, dtype=np.float64)
        B, L = model.shape[0], model.shape[1]
        assert model.shape == (B, L, 4) and nlay.shape == (B,) and noise.size == B * 2 * self.ntargets
        if rho is not None:
            rho = np.ascontiguousarray(rho, dtype=np.float64)
            assert rho.shape == (B, L)
        if out is None:
            logL = np.empty(B)
            misfits = np.empty((B, self.ntargets + 1))
            status = np.empty(B, dtype=np.int32)
            synth = np.empty((B, self.synth_stride)) if want_synth else None
        else:
            logL, misfits, status, synth = out
        _lib.check(self._lib.bh_engine_eval_host(
            self._h, model.ctypes.data, nlay.ctypes.data, noise.ctypes.data,
            rho.ctypes.data if rho is not None else None, B, L,
            logL.ctypes.data, misfits.ctypes.data, status.ctypes.data,
            synth.ctypes.data if synth is not None else None))
        return logL, misfits, status, synth

    def loglik_host(self, synth, tvalid, noise):
        """Likelihood of caller-supplied modelled data (bh_engine_loglik_host): synth [B, synth_stride],
        tvalid [B, T] (0: that target's synthetic was rejected), noise [B, 2T] ->
        (logL [B], misfits [B, T+1], status [B])."""
        synth = np.ascontiguousarray(synth, dtype=np.float64)
        tvalid = np.ascontiguousarray(tvalid, dtype=np.int32)
        noise = np.ascontiguousarray(noise, dtype=np.float64)
        B = synth.shape[0]
        assert synth.shape == (B, self.synth_stride) and tvalid.shape == (B, self.ntargets)
        assert noise.size == B * 2 * self.ntargets
        logL = np.empty(B)
        misfits = np.empty((B, self.ntargets + 1))
        status = np.empty(B, dtype=np.int32)
        _lib.check(self._lib.bh_engine_loglik_host(self._h, synth.ctypes.data, tvalid.ctypes.data, noise.ctypes.data,
                                                   B, logL.ctypes.data, misfits.ctypes.data, status.ctypes.data))
        return logL, misfits, status

    def submit_host(self, model, nlay, noise, out, rho=None):
        """Asynchronous form of eval_host (bh_engine_eval_host_async): enqueues the copies and kernels
        and returns a ticket at once, so that the caller can prepare and submit the next batch while
        this one runs (at most two in flight).  `out` = (logL, misfits, status, synth | None) numpy
        arrays that receive the results when wait(ticket) returns; inputs and outputs must stay alive
        and untouched until then."""
        B, L = model.shape[0], model.shape[1]
        for a, dt in ((model, np.float64), (nlay, np.int32), (noise, np.float64)):
            assert a.dtype == dt and a.flags["C_CONTIGUOUS"]
        assert model.shape == (B, L, 4) and nlay.shape == (B,) and noise.size == B * 2 * self.ntargets
        logL, misfits, status, synth = out
        ticket = ctypes.c_longlong(0)
        _lib.check(self._lib.bh_engine_eval_host_async(
            self._h, model.ctypes.data, nlay.ctypes.data, noise.ctypes.data,
            rho.ctypes.data if rho is not None else None, B, L,
            logL.ctypes.data, misfits.ctypes.data, status.ctypes.data,
            synth.ctypes.data if synth is not None else None, ctypes.byref(ticket)))
        return ticket.value

    def wait(self, ticket):
        """Block until the outputs of submit_host(ticket) are written."""
        _lib.check(self._lib.bh_engine_wait(self._h, int(ticket)))

    def eval_host_ptr(self, model_ptr, nlay_ptr, noise_ptr, B, L, logL_ptr, misfits_ptr, status_ptr,
                      rho_ptr=None, synth_ptr=None):
        """Raw-pointer form of eval_host (pinned torch tensors: pass .data_ptr())."""
        _lib.check(self._lib.bh_engine_eval_host(self._h, model_ptr, nlay_ptr, noise_ptr, rho_ptr,
                                                 int(B), int(L), logL_ptr, misfits_ptr, status_ptr,
                                                 synth_ptr))

    def split_synth(self, synth):
        """Per-target views of a synth row block."""
        o = self.synth_offsets
        return [synth[..., o[i]:o[i + 1]] for i in range(self.ntargets)]
