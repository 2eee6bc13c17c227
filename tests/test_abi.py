"""CPU tests of the boundary: the shared library loads, exports every symbol the header
declares, and the product path fails loudly without a CUDA device (no CPU fallback)."""
import os
import re

import numpy as np
import pytest

from tests.conftest import ROOT, has_cuda


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "bayhunter_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    # every bh_* entry point plus the reference's raw native symbols exported for link-level drop-in
    return sorted(set(re.findall(r"\b(bh_[a-z0-9_]+|surfdisp96_|synrf_cwrap)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from bayhunter_b200 import _lib
    lib = _lib.load()
    names = _declared_symbols()
    assert len(names) >= 12 and "surfdisp96_" in names and "synrf_cwrap" in names
    for n in names:
        assert hasattr(lib, n), "missing symbol " + n
        assert n in _lib.SIGNATURES, "untyped symbol " + n
    assert lib.bh_abi_version() == 1


def test_struct_layout_matches_header(tmp_path):
    """ctypes mirror of struct bh_target vs the C compiler's view of the header."""
    import ctypes
    import subprocess
    from bayhunter_b200 import _lib
    fields = [n for n, _ in _lib.BhTarget._fields_]
    src = tmp_path / "probe.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "bayhunter_b200.h"\n'
                   'int main(void){printf("%zu", sizeof(bh_target));\n' +
                   "".join('printf(" %%zu", offsetof(bh_target, %s));\n' % f for f in fields) +
                   "return 0;}\n")
    exe = tmp_path / "probe"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    vals = [int(v) for v in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    assert vals[0] == ctypes.sizeof(_lib.BhTarget)
    assert vals[1:] == [getattr(_lib.BhTarget, f).offset for f in fields]


@pytest.mark.skipif(has_cuda(), reason="checks the no-device behaviour")
def test_no_cpu_fallback_without_device():
    import bayhunter_b200 as bh
    from bayhunter_b200._lib import BayHunterB200Error
    x = np.linspace(1, 40, 20)
    with pytest.raises(BayHunterB200Error):
        bh.Engine([bh.TargetSpec("rdispph", x, np.ones(20) * 3.5)], 4, 6)
    with pytest.raises(BayHunterB200Error):
        bh.SurfDisp(x, "rdispph").run_model(np.array([5., 0.]), np.array([6., 8.]),
                                            np.array([3.5, 4.5]), np.array([2.7, 3.3]))
    with pytest.raises(BayHunterB200Error):
        t = -5 + 0.2 * np.arange(201)
        bh.RFminiModRF(t, "prf").run_model(np.array([5., 0.]), np.array([6., 8.]),
                                           np.array([3.5, 4.5]), np.array([2.7, 3.3]))


def test_product_package_does_not_import_the_oracle():
    """The oracle is test infrastructure: nothing under bayhunter_b200/ may reference it."""
    pkg = os.path.join(ROOT, "bayhunter_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "joint_oracle" not in src and "liboracle" not in src and "import oracle" not in src, f


def test_plugins_pickle_and_host_logic():
    import pickle
    import bayhunter_b200 as bh
    from bayhunter_b200 import Targets
    x = np.linspace(1, 40, 20)
    p = pickle.loads(pickle.dumps(bh.SurfDisp(x, "rdispgr")))
    assert (p.wavetype, p.veltype) == (2, 1) and p.kmax == 20
    with pytest.raises(ReferenceError):
        bh.SurfDisp(x, "nonsense")
    t = -5 + 0.2 * np.arange(201)
    r = pickle.loads(pickle.dumps(bh.RFminiModRF(t, "srf")))
    assert r.nsamp == 512 and abs(r.fsamp - 5.0) < 1e-12 and r.tshft == 5.0
    assert r.modelparams["wtype"] == "SV"
    with pytest.raises(ValueError):
        bh.RFminiModRF(np.array([0., 0.2, 0.5]), "prf")
    big = bh.SurfDisp(np.linspace(1, 50, 80), "rdispph")
    assert big.obsx_int.size == 60
    t1 = Targets.RayleighDispersionPhase(x, np.ones(20) * 3.5)
    assert t1.covariance_law() == "exp"
    t1.get_covariance = t1.valuation.get_covariance_nocorr
    assert t1.covariance_law() == "white"
    jt = pickle.loads(pickle.dumps(Targets.JointTarget([t1])))
    assert jt.ntargets == 1 and jt._engine is None
    assert np.all(np.isnan(Targets.ObservedData(x, x, yerr=-np.ones(20)).yerr))


def test_model_packer_matches_get_vp_vs_h():
    from bayhunter_b200 import Model, pack_models
    rng = np.random.default_rng(0)
    B, kmax = 16, 7
    models = np.full((B, 2 * kmax), np.nan)
    vpvs = rng.uniform(1.4, 2.1, B)
    for b in range(B):
        k = int(rng.integers(1, kmax + 1))
        models[b, :k] = np.sort(rng.uniform(2, 5, k))
        models[b, kmax:kmax + k] = np.sort(rng.uniform(0, 60, k))
    rows, nlay = pack_models(models, vpvs, lmax=kmax)
    for b in range(B):
        vp, vs, h = Model.get_vp_vs_h(models[b], vpvs[b])
        n = nlay[b]
        assert n == vs.size
        assert np.array_equal(rows[b, :n, 0], vs) and np.array_equal(rows[b, :n, 3], h)
        assert np.array_equal(rows[b, :n, 0] * rows[b, :n, 1], vp)       # device forms vp exactly like this
        assert np.array_equal(rows[b, :n, 2], np.concatenate(([0], np.cumsum(h)[:-1])))
    vp, vs, h = Model.get_vp_vs_h(models[0], vpvs[0], mantle=(4.3, 1.8))
    assert np.all(vp[vs >= 4.3] == vs[vs >= 4.3] * 1.8) or not (vs >= 4.3).any()
