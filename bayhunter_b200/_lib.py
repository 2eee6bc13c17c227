"""ctypes binding of libbayhunter_b200.so (C ABI in include/bayhunter_b200.h).

The library is the product: if it cannot be loaded (or built), importing the
compute entry points raises -- there is no Python/NumPy fallback for the hot path.
"""
import ctypes
import os
import threading

from . import build as _build

c_double_p = ctypes.POINTER(ctypes.c_double)
c_float_p = ctypes.POINTER(ctypes.c_float)
c_int_p = ctypes.POINTER(ctypes.c_int)

BH_OK = 0
BH_ERR_ARG, BH_ERR_CUDA, BH_ERR_UNSUPPORTED, BH_ERR_NO_DEVICE = -1, -2, -3, -4

REF_CODES = {"rdispph": 0, "rdispgr": 1, "ldispph": 2, "ldispgr": 3, "prf": 4, "srf": 5}
REF_GENERIC = 6     # modelled data supplied by the caller (bh_engine_loglik_host)
COV_EXP, COV_WHITE, COV_WHITE_SCALED, COV_GAUSS = 0, 1, 2, 3
MAX_TARGETS, MAX_PERIODS, MAX_LAYERS = 8, 60, 100
NUM_COUNTERS = 2 + 2 * MAX_TARGETS
KERNEL_NAMES = ("prepare_swd", "swd", "prepare_rf", "rf_spectrum", "rf_synth", "loglik", "swd_love", "swd_general",
                "swd_pool", "swd_pool_love")


class BhTarget(ctypes.Structure):
    """struct bh_target (include/bayhunter_b200.h)."""
    _fields_ = [
        ("ref", ctypes.c_int), ("n", ctypes.c_int),
        ("x", c_double_p), ("y", c_double_p), ("yerr", c_double_p),
        ("cov", ctypes.c_int), ("corr_inv", c_double_p), ("logcorr_det", ctypes.c_double),
        ("mode", ctypes.c_int), ("flsph", ctypes.c_int),
        ("gauss", ctypes.c_double), ("p", ctypes.c_double), ("nsv", ctypes.c_double),
        ("qp", ctypes.c_double), ("qs", ctypes.c_double),
    ]


class BhSamplerConfig(ctypes.Structure):
    """struct bh_sampler_config (include/bayhunter_b200.h)."""
    _fields_ = [
        ("layers_min", ctypes.c_int), ("layers_max", ctypes.c_int),
        ("vs_min", ctypes.c_double), ("vs_max", ctypes.c_double),
        ("z_min", ctypes.c_double), ("z_max", ctypes.c_double),
        ("vpvs_fixed", ctypes.c_int), ("vpvs_min", ctypes.c_double), ("vpvs_max", ctypes.c_double),
        ("has_mantle", ctypes.c_int), ("mantle_vs", ctypes.c_double), ("mantle_vpvs", ctypes.c_double),
        ("noise_fixed", ctypes.c_int * (2 * MAX_TARGETS)),
        ("noise_min", ctypes.c_double * (2 * MAX_TARGETS)), ("noise_max", ctypes.c_double * (2 * MAX_TARGETS)),
        ("thickmin", ctypes.c_double),
        ("has_lvz", ctypes.c_int), ("has_hvz", ctypes.c_int),
        ("lvz", ctypes.c_double), ("hvz", ctypes.c_double),
        ("propdist", ctypes.c_double * 5), ("acceptance", ctypes.c_double * 2),
        ("iter_burnin", ctypes.c_int), ("iter_main", ctypes.c_int), ("max_accepted", ctypes.c_int),
        ("seed", ctypes.c_ulonglong),
    ]


class BayHunterB200Error(RuntimeError):
    def __init__(self, code, message):
        super().__init__("libbayhunter_b200 error %d: %s" % (code, message))
        self.code = code


# every symbol include/bayhunter_b200.h declares: (restype, argtypes)
SIGNATURES = {
    "bh_abi_version": (ctypes.c_int, []),
    "bh_last_error": (ctypes.c_char_p, []),
    "bh_device_count": (ctypes.c_int, []),
    "bh_set_device": (ctypes.c_int, [ctypes.c_int]),
    "bh_engine_create": (ctypes.c_int, [ctypes.POINTER(BhTarget), ctypes.c_int, ctypes.c_int,
                                        ctypes.c_int, ctypes.POINTER(ctypes.c_void_p)]),
    "bh_engine_destroy": (None, [ctypes.c_void_p]),
    "bh_engine_synth_stride": (ctypes.c_int, [ctypes.c_void_p]),
    "bh_engine_is_tuning": (ctypes.c_int, [ctypes.c_void_p]),
    "bh_engine_capture_begin": (ctypes.c_int, [ctypes.c_void_p]),
    "bh_engine_capture_end": (ctypes.c_int, [ctypes.c_void_p]),
    "bh_engine_set": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int]),
    "bh_engine_eval": (ctypes.c_int, [ctypes.c_void_p] + [ctypes.c_void_p] * 4 +
                       [ctypes.c_int, ctypes.c_int] + [ctypes.c_void_p] * 5),
    "bh_engine_eval_host": (ctypes.c_int, [ctypes.c_void_p] + [ctypes.c_void_p] * 4 +
                            [ctypes.c_int, ctypes.c_int] + [ctypes.c_void_p] * 4),
    "bh_engine_eval_host_async": (ctypes.c_int, [ctypes.c_void_p] + [ctypes.c_void_p] * 4 +
                                  [ctypes.c_int, ctypes.c_int] + [ctypes.c_void_p] * 4 +
                                  [ctypes.POINTER(ctypes.c_longlong)]),
    "bh_engine_wait": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_longlong]),
    "bh_engine_loglik_host": (ctypes.c_int, [ctypes.c_void_p] * 4 + [ctypes.c_int] + [ctypes.c_void_p] * 3),
    "bh_engine_last_kernel_ms": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_float)]),
    "bh_engine_last_counts": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_longlong)]),
    "bh_surfdisp96": (ctypes.c_int, [c_float_p] * 4 + [ctypes.c_int] * 6 +
                      [c_double_p, c_double_p, c_int_p]),
    "bh_debug_math": (ctypes.c_int, [ctypes.c_int, c_double_p, c_double_p]),
    "bh_sampler_create": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(BhSamplerConfig), ctypes.c_int,
                                         ctypes.c_int, ctypes.c_longlong, ctypes.POINTER(ctypes.c_void_p)]),
    "bh_sampler_destroy": (None, [ctypes.c_void_p]),
    "bh_sampler_init": (ctypes.c_int, [ctypes.c_void_p] * 5),
    "bh_sampler_run": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
    "bh_sampler_get_state": (ctypes.c_int, [ctypes.c_void_p] * 13),
    "bh_sampler_get_overflow": (ctypes.c_int, [ctypes.c_void_p] * 3),
    "bh_sampler_get_chains": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int] + [ctypes.c_void_p] * 6),
    "bh_sampler_set_state": (ctypes.c_int, [ctypes.c_void_p] * 11),
    "bh_sampler_get_proposal": (ctypes.c_int, [ctypes.c_void_p] * 10),
    "bh_sampler_set_forced_draws": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    "surfdisp96_": (None, [c_float_p] * 4 + [c_int_p] * 6 + [c_double_p, c_double_p, c_int_p]),
    "synrf_cwrap": (ctypes.c_int, [ctypes.c_int] + [ctypes.c_double] * 6 + [ctypes.c_int] * 2 + [c_double_p] * 9),
    "bh_correlated_noise": (ctypes.c_int, [ctypes.c_int] * 3 + [ctypes.c_double] * 2 + [ctypes.c_ulonglong, c_double_p, c_double_p]),
    "bh_measure_fp64_peak": (ctypes.c_int, [c_double_p, c_double_p]),
    "bh_synrf": (ctypes.c_int, [ctypes.c_int] + [ctypes.c_double] * 6 + [ctypes.c_int] * 2 +
                 [c_double_p] * 9),
}

_lock = threading.Lock()
_lib = None


def library_path():
    """In-tree shared object; BH_B200_LIB selects another build of the SAME ABI
    (developer A/B runs of kernel variants, tools/quick_bench.py)."""
    return os.environ.get("BH_B200_LIB") or _build.LIB


def load(build_if_missing=True):
    """Load (building first if the shared object is missing) and type the ABI."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        path = library_path()
        if not os.path.exists(path):
            if not build_if_missing or path != _build.LIB:
                raise BayHunterB200Error(BH_ERR_NO_DEVICE, "%s is missing; run __graft_entry__.build()" % path)
            _build.build()
        lib = ctypes.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)      # AttributeError if the ABI lost a symbol
            fn.restype = res
            fn.argtypes = args
        if lib.bh_abi_version() != 1:
            raise BayHunterB200Error(BH_ERR_ARG, "ABI version mismatch")
        _lib = lib
        return lib


def check(code):
    if code != BH_OK:
        msg = load().bh_last_error()
        raise BayHunterB200Error(code, msg.decode() if msg else "")
    return code


def set_device(index):
    """Select the GPU of this process (torchrun: LOCAL_RANK) for engines / samplers created later."""
    check(require_device().bh_set_device(int(index)))


def require_device():
    """Fail loudly when no CUDA device is usable -- the hot path has no CPU route."""
    lib = load()
    if lib.bh_device_count() < 1:
        raise BayHunterB200Error(BH_ERR_NO_DEVICE,
                                 "no CUDA device visible; bayhunter_b200 has no CPU path")
    return lib
