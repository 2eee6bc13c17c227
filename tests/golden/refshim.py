"""tests/golden/refshim.py -- TEST INFRASTRUCTURE (fixture generation only).

Imports the UNMODIFIED reference Python package from /root/reference/src as
`BayHunter` inside this container, where it cannot be imported as is
(SURVEY.md D.5): matplotlib / configobj are absent, numpy >= 2 dropped the
`np.float`, `np.int`, `np.product` aliases, and the two compiled extensions
(`surfdisp96_ext` from f2py + gfortran, `rfmini` from Cython) cannot be built
with the reference's own build system.

The shim provides, OUTSIDE the reference tree:
  * stub modules for matplotlib (never called by the code paths used here) and a
    minimal ConfigObj that reads the INI dialect of `defaults/defaults.ini`,
  * the removed numpy aliases,
  * `BayHunter.rfmini.synrf`      -> the reference's own rfmini C++ compiled in place
                                    (oracle/_ref/librfmini_ref.so, `synrf_cwrap`),
    with the marshalling of rfmini.pyx:74-114,
  * `BayHunter.surfdisp96_ext.surfdisp96` -> oracle/surf96_oracle.c (no Fortran
    compiler exists here), with f2py's REAL*4 cast and in-place `cg` write.

Everything above those two native entry points -- Targets.py, Models.py,
SingleChain.py, surf96_modsw.py, rfmini_modrf.py, SynthObs.py -- is the
reference's own code, which is what the generated fixtures pin.

Only used by the generator scripts in this directory; nothing at test / bench /
product run time imports it (the GPU box has no /root/reference).
"""
import ctypes
import importlib
import importlib.util
import os
import sys
import types

import numpy as np

REFERENCE_SRC = "/root/reference/src"
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


class _Anything(types.ModuleType):
    """Module stub: any attribute is a callable that returns another stub."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Callable(self.__name__ + "." + name)


class _Callable(object):
    def __init__(self, name):
        self._name = name

    def __call__(self, *a, **k):
        return _Callable(self._name + "()")

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Callable(self._name + "." + name)

    def __iter__(self):
        return iter(())

    def __mro_entries__(self, bases):     # `class X(stub.Button)` in BayWatch
        return (object,)


class _Section(dict):
    pass


class ConfigObj(dict):
    """The subset of configobj.ConfigObj that utils.load_params uses: sections of
    `key = value` lines, comma separated values become lists of strings, inline
    comments start with '#'."""

    def __init__(self, infile=None):
        dict.__init__(self)
        self.sections = []
        if infile is None:
            return
        cur = None
        with open(infile) as fh:
            for raw in fh:
                line = raw.split("#", 1)[0].strip()
                if not line:
                    continue
                if line.startswith("[") and line.endswith("]"):
                    name = line[1:-1].strip()
                    cur = _Section()
                    self[name] = cur
                    self.sections.append(name)
                    continue
                key, val = [s.strip() for s in line.split("=", 1)]
                if "," in val and not (val.startswith("(") or val.startswith("'")):
                    val = [v.strip() for v in val.split(",") if v.strip()]
                elif len(val) >= 2 and val[0] == val[-1] and val[0] in "'\"":
                    val = val[1:-1]
                cur[key] = val


def _install_stubs():
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.cm", "matplotlib.colors",
                 "matplotlib.widgets", "matplotlib.collections", "mpl_toolkits",
                 "mpl_toolkits.axes_grid", "mpl_toolkits.axes_grid.inset_locator", "PyPDF2"):
        if name not in sys.modules:
            sys.modules[name] = _Anything(name)
    if "configobj" not in sys.modules:
        m = types.ModuleType("configobj")
        m.ConfigObj = ConfigObj
        sys.modules["configobj"] = m
    for alias, target in (("float", float), ("int", int), ("bool", bool), ("product", np.prod)):
        if not hasattr(np, alias):
            setattr(np, alias, target)


_D = ctypes.POINTER(ctypes.c_double)
_F = ctypes.POINTER(ctypes.c_float)


def _native_modules():
    sys.path.insert(0, ROOT)
    from oracle import joint_oracle as jo
    jo.build()
    R = jo.ref_rfmini()
    if R is None:
        raise RuntimeError("oracle/_ref/librfmini_ref.so missing: run make -C oracle")
    L = jo.lib()

    rfmini = types.ModuleType("BayHunter.rfmini")

    def synrf(z_arr, vp_arr, vs_arr, rh_arr, qp_arr, qs_arr, p, a, nsamp, fsamp, tshift, nsv, sigma, wave):
        waveno = ["P", "SV", "SH"].index(wave)
        nsamp = int(nsamp)
        arrs = [np.ascontiguousarray(x, dtype=np.float64) for x in (z_arr, vp_arr, vs_arr, rh_arr, qp_arr, qs_arr)]
        fz, fr, rf = np.zeros(nsamp), np.zeros(nsamp), np.zeros(nsamp)
        R.synrf_cwrap(nsamp, float(fsamp), float(tshift), float(p), float(a), float(nsv), float(sigma),
                      waveno, len(arrs[0]), *[x.ctypes.data_as(_D) for x in arrs + [fz, fr, rf]])
        return fz, fr, rf

    rfmini.synrf = synrf

    ext = types.ModuleType("BayHunter.surfdisp96_ext")

    def surfdisp96(thkm, vpm, vsm, rhom, nlayer, iflsph, iwave, mode, igr, kmax, t, cg):
        f = [np.ascontiguousarray(x, dtype=np.float32) for x in (thkm, vpm, vsm, rhom)]    # f2py cast
        assert t.dtype == np.float64 and cg.dtype == np.float64 and cg.flags.c_contiguous
        err = ctypes.c_int(0)
        ns = (ctypes.c_long * 2)()
        L.surf96_oracle(*[x.ctypes.data_as(_F) for x in f], int(nlayer), int(iflsph), int(iwave), int(mode),
                        int(igr), int(kmax), t.ctypes.data_as(_D), cg.ctypes.data_as(_D), ctypes.byref(err), ns)
        return err.value

    ext.surfdisp96 = surfdisp96
    return rfmini, ext


def import_reference():
    """Return the reference package (module `BayHunter`) imported from /root/reference/src."""
    if "BayHunter" in sys.modules and getattr(sys.modules["BayHunter"], "__refshim__", False):
        return sys.modules["BayHunter"]
    if not os.path.isdir(REFERENCE_SRC):
        raise RuntimeError("the reference tree is not present (fixtures are generated in the build container only)")
    _install_stubs()
    rfmini, ext = _native_modules()
    sys.modules["BayHunter.rfmini"] = rfmini
    sys.modules["BayHunter.surfdisp96_ext"] = ext
    spec = importlib.util.spec_from_file_location(
        "BayHunter", os.path.join(REFERENCE_SRC, "__init__.py"), submodule_search_locations=[REFERENCE_SRC])
    pkg = importlib.util.module_from_spec(spec)
    pkg.__refshim__ = True
    sys.modules["BayHunter"] = pkg
    pkg.rfmini = rfmini
    pkg.surfdisp96_ext = ext
    spec.loader.exec_module(pkg)
    return pkg
