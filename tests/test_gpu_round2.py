"""GPU tests of the round-2 additions: the lock-step dispersion kernel, the asynchronous host entry,
the link-level drop-in symbols, the batched Gauss-law contraction, and a full-depth ragged batch."""
import ctypes

import numpy as np
import pytest

from tests.test_gpu_parity import _make_targets, _compare, _logl_check, _rel

pytestmark = pytest.mark.gpu


def _engine(cfg, B, seed, rows_spec=None, laws=None):
    from bayhunter_b200 import Engine, synthetic
    c = synthetic.CONFIGS[cfg]
    rng = np.random.default_rng(seed)
    specs, otargets = _make_targets(c["refs"], c["periods"], c["rf"], rng, laws=laws)
    rows, nlay = synthetic.draw_batch(B, rows_spec or c["nrows"], seed=seed + 1)
    noise = synthetic.draw_noise(B, c["refs"], seed=seed + 2)
    return Engine(specs, B, rows.shape[1]), specs, otargets, rows, nlay, noise


@pytest.mark.parametrize("cfg,B", [("joint5", 700), ("transd3", 300), ("swd2", 257)])
def test_lockstep_kernel_is_bit_identical(cfg, B):
    """swd_lockstep_kernel consumes exactly swd_kernel's candidate sequence: same bits, any layout."""
    eng, specs, _, rows, nlay, noise = _engine(cfg, B, 40)
    base = eng.eval_host(rows, nlay, noise, want_synth=True)
    cons0, _ = eng.last_counts()
    for grp, ph in ((0, 0), (8, 16), (16, 20), (5, 9)):
        eng.set(swd_lockstep=1, swd_ls_spw_group=grp, swd_ls_spw_phase=ph)
        out = eng.eval_host(rows, nlay, noise, want_synth=True)
        cons, ev = eng.last_counts()
        assert cons == cons0 and ev >= cons
        for a, b in zip(out, base):
            assert np.array_equal(a, b, equal_nan=True), (cfg, grp, ph)


@pytest.mark.parametrize("cfg,B", [("joint5", 700), ("transd3", 300), ("swd2", 257), ("joint5", 5), ("joint5", 4200), ("joint5", 8192)])
def test_pool_kernel_is_bit_identical(cfg, B):
    """swd_pool_kernel (a CTA's 128 lanes dealt over the chains of M models) consumes swd_kernel's candidate sequence;
    at BASELINE's full batch (joint5, 8192 chains) the engine's rule picks it."""
    eng, specs, _, rows, nlay, noise = _engine(cfg, B, 41)
    eng.set(swd_pool=0, profile=1)
    base = eng.eval_host(rows, nlay, noise, want_synth=True)
    cons0, _ = eng.last_counts()
    assert "swd" in eng.last_kernel_ms() and "swd_pool" not in eng.last_kernel_ms()
    for mode, m in ((-1, 0),) if B > 4096 else ((1, 0), (1, 1), (1, 13), (1, 42), (1, 128)):
        eng.set(swd_pool=mode, swd_pool_models=m)
        out = eng.eval_host(rows, nlay, noise, want_synth=True)
        cons, ev = eng.last_counts()
        assert "swd_pool" in eng.last_kernel_ms(), "the pool kernel did not run"
        assert cons == cons0 and ev >= cons
        for a, b in zip(out, base):
            assert np.array_equal(a, b, equal_nan=True), (cfg, m)


@pytest.mark.parametrize("refs,nrows,B", [(("ldispph", "ldispgr"), (3, 31), 4300),          # Love only, deep: pool by rule
                                          (("rdispph", "rdispgr", "ldispph", "ldispgr"), (3, 31), 4000),   # both waves, deep
                                          (("rdispgr", "ldispgr"), 6, 4100),                 # group curves only
                                          (("rdispph", "ldispph"), 6, 4100),                 # phase curves only
                                          (("rdispph", "rdispph", "ldispgr"), 6, 4100)])     # two phase curves of a wave: no pool
def test_pool_rule_on_other_curve_sets(refs, nrows, B):
    """The engine's kernel rule on curve sets outside the BASELINE configurations: whatever it picks equals swd_kernel."""
    from bayhunter_b200 import Engine, synthetic
    rng = np.random.default_rng(43)
    specs, _ = _make_targets(refs, np.linspace(1, 40, 18), None, rng)
    rows, nlay = synthetic.draw_batch(B, nrows, seed=44)
    noise = synthetic.draw_noise(B, refs, seed=45)
    eng = Engine(specs, B, rows.shape[1])
    eng.set(swd_pool=0, profile=1)
    base = eng.eval_host(rows, nlay, noise, want_synth=True)
    cons0, _ = eng.last_counts()
    eng.set(swd_pool=-1)
    out = eng.eval_host(rows, nlay, noise, want_synth=True)
    cons, ev = eng.last_counts()
    picked_pool = "swd_pool" in eng.last_kernel_ms() or "swd_pool_love" in eng.last_kernel_ms()
    if nrows != 6:
        assert picked_pool, "deep models: the rule takes the pool kernel"
    if len(set(refs)) != len(refs):
        assert not picked_pool, "two curves of one kind in a wave type: swd_kernel"
    assert cons == cons0 and ev >= cons
    for a, b in zip(out, base):
        assert np.array_equal(a, b, equal_nan=True), refs


def test_async_host_entry_overlaps_and_matches():
    """bh_engine_eval_host_async / bh_engine_wait with pageable numpy buffers, two calls in flight."""
    eng, specs, _, rows, nlay, noise = _engine("joint5", 512, 50)
    from bayhunter_b200 import synthetic
    batches = [synthetic.perturb_batch(rows, nlay, np.random.default_rng(s)) for s in range(5)]
    want = [eng.eval_host(b, nlay, noise) for b in batches]
    T = eng.ntargets
    outs = [(np.empty(512), np.empty((512, T + 1)), np.empty(512, dtype=np.int32), None) for _ in batches]
    tickets = []
    for i, b in enumerate(batches):
        tickets.append(eng.submit_host(b, nlay, noise, outs[i]))
        if i >= 1:
            eng.wait(tickets[i - 1])
            for a, w in zip(outs[i - 1][:3], want[i - 1][:3]):
                assert np.array_equal(a, w)
    eng.wait(tickets[-1])
    eng.wait(tickets[0])                       # waiting twice is harmless
    for a, w in zip(outs[-1][:3], want[-1][:3]):
        assert np.array_equal(a, w)


def test_reference_link_symbols(oracle):
    """surfdisp96_ (gfortran calling convention, all by reference, arrays of 100 / 60) and synrf_cwrap
    (the rfmini prototype, returns 1) called through bare ctypes like the reference's own glue would."""
    import bayhunter_b200 as bh
    from bayhunter_b200 import _lib
    lib = ctypes.CDLL(_lib.library_path())
    h = np.array([5., 23., 8., 0.]); vs = np.array([2.7, 3.6, 3.8, 4.4]); vp = vs * 1.73
    rho = vp * 0.32 + 0.77
    periods = np.linspace(1, 40, 25)
    for ref, (iwave, igr) in {"rdispph": (2, 0), "rdispgr": (2, 1), "ldispph": (1, 0), "ldispgr": (1, 1)}.items():
        arr = [np.zeros(100, dtype=np.float32) for _ in range(4)]
        for a, v in zip(arr, (h, vp, vs, rho)):
            a[:4] = v
        t = np.zeros(60); t[:25] = periods
        cg = np.zeros(60)
        ints = [ctypes.c_int(v) for v in (4, 0, iwave, 1, igr, 25)]
        err = ctypes.c_int(-7)
        fp = ctypes.POINTER(ctypes.c_float); dp = ctypes.POINTER(ctypes.c_double)
        lib.surfdisp96_.restype = None
        lib.surfdisp96_(*[a.ctypes.data_as(fp) for a in arr], *[ctypes.byref(i) for i in ints],
                        t.ctypes.data_as(dp), cg.ctypes.data_as(dp), ctypes.byref(err))
        assert err.value == 0
        _, want = bh.SurfDisp(periods, ref).run_model(h, vp, vs, rho)
        assert np.array_equal(cg[:25], want) and np.all(cg[25:] == 0)
        _, yo = oracle.surfdisp(h, vp, vs, rho, ref, periods)
        assert _rel(cg[:25], yo).max() <= (1e-6 if igr == 0 else 5e-5)
    # a model without a Love root: err = 1
    arr = [np.zeros(100, dtype=np.float32) for _ in range(4)]
    for a, v in zip(arr, ([0.], [6.], [3.5], [2.7])):
        a[:1] = v
    ints = [ctypes.c_int(v) for v in (1, 0, 1, 1, 0, 5)]
    t = np.zeros(60); t[:5] = [1, 2, 3, 4, 5]
    cg = np.ones(60)
    lib.surfdisp96_(*[a.ctypes.data_as(fp) for a in arr], *[ctypes.byref(i) for i in ints],
                    t.ctypes.data_as(dp), cg.ctypes.data_as(dp), ctypes.byref(err))
    assert err.value == 1
    # synrf_cwrap
    nsamp = 512
    z = np.array([0., 5., 28., 36.])
    qp = np.full(4, 500.); qs = np.full(4, 225.)
    fz, fr, rf = np.ones(nsamp), np.ones(nsamp), np.zeros(nsamp)
    lib.synrf_cwrap.restype = ctypes.c_int
    lib.synrf_cwrap.argtypes = [ctypes.c_int] + [ctypes.c_double] * 6 + [ctypes.c_int] * 2 + [dp] * 9
    kappa = vp[0] / vs[0]
    sigma = (2 - kappa ** 2) / (2 - 2 * kappa ** 2)
    ret = lib.synrf_cwrap(nsamp, 5.0, 5.0, 6.4, 1.0, vs[0], sigma, 0, 4, *[a.ctypes.data_as(dp) for a in
                                                                          (z, vp, vs, rho, qp, qs, fz, fr, rf)])
    assert ret == 1
    x = -5.0 + 0.2 * np.arange(201)
    _, want = bh.RFminiModRF(x, "prf").run_model(h, vp, vs, rho)
    assert np.array_equal(rf[:201], want)
    _, yo = oracle.recfunc(h, vp, vs, rho, x)
    assert np.abs(rf[:201] - yo).max() <= 1e-9 * np.abs(yo).max()


def test_gauss_law_batched_contraction(oracle):
    """joint5 with the Gauss law on the 512-sample receiver function (r = 0.9, rcond = 1e-5: SURVEY 8d's
    second run): the tensor-core contraction against numpy's dense d^T R^-1 d."""
    eng, specs, otargets, rows, nlay, noise = _engine("joint5", 200, 60, laws={"prf": "gauss"})
    noise[:, 8] = 0.9
    out = eng.eval_host(rows, nlay, noise, want_synth=True)
    ref = oracle.evaluate_batch(otargets, rows, nlay, noise)
    _compare(out, ref, [(s.ref, s.n) for s in specs])
    ok = ref[2] == 1
    assert ok.sum() > 100
    # the likelihood term of the Gauss target alone, from the engine's own synthetic traces
    o = sum(s.n for s in specs[:4])
    d = out[3][ok, o:o + 512] - specs[4].y
    phi = np.einsum("bi,ij,bj->b", d, specs[4].corr_inv, d)
    sig = noise[ok, 9]
    want = -0.5 * (512 * np.log(2 * np.pi) + 2 * 512 * np.log(sig) + specs[4].logcorr_det) - 0.5 * phi / sig ** 2
    eng_exp, _, _, _, _, _ = _engine("joint5", 200, 60)
    # same targets with the Gauss term removed: subtract the first four targets' logL computed by the oracle
    o4 = [oracle.OracleTarget(s.ref, s.x, s.y, cov="exp") for s in specs[:4]]
    l4 = oracle.evaluate_batch(o4, rows, nlay, np.ascontiguousarray(noise[:, :8]))[0]
    got = out[0][ok] - l4[ok]
    assert np.abs(got - want).max() <= 1e-6 * np.abs(want).max()
    e = _logl_check(out[0], ref[0], ok)
    assert e.max() <= 1e-5, e.max()


def test_full_depth_ragged_batch_matches_oracle(oracle):
    """BASELINE config 4 shape: 3..31 rows, Rayleigh phase + group + P-RF, B = 1024, against the oracle."""
    eng, specs, otargets, rows, nlay, noise = _engine("transd3", 1024, 70)
    assert nlay.max() == 31 and nlay.min() == 3
    out = eng.eval_host(rows, nlay, noise, want_synth=True)
    ref = oracle.evaluate_batch(otargets, rows, nlay, noise)
    _compare(out, ref, [(s.ref, s.n) for s in specs])
    ok = ref[2] == 1
    e = _logl_check(out[0], ref[0], ok)
    assert np.median(e) <= 1e-6 and (e <= 1e-6).mean() >= 0.95, (np.median(e), e.max())


def test_fp64_peak_probe():
    from bayhunter_b200 import _lib
    lib = _lib.require_device()
    tf, mhz = ctypes.c_double(0), ctypes.c_double(0)
    _lib.check(lib.bh_measure_fp64_peak(ctypes.byref(tf), ctypes.byref(mhz)))
    assert 15.0 < tf.value < 60.0 and mhz.value > 500


# ---- API edges and the advisor's findings -------------------------------------------------------
ST3 = dict(h=np.array([5., 23., 8., 0.]), vs=np.array([2.7, 3.6, 3.8, 4.4]))


def _st3_joint(oracle):
    from bayhunter_b200 import Targets
    h, vs = ST3["h"], ST3["vs"]
    vp = vs * 1.73
    rho = vp * 0.32 + 0.77
    periods = np.linspace(1, 41, 21)
    xrf = -5.0 + 0.2 * np.arange(201)
    _, ysw = oracle.surfdisp(h, vp, vs, rho, "rdispph", periods)
    _, yrf = oracle.recfunc(h, vp, vs, rho, xrf)
    rng = np.random.default_rng(3)
    jt = Targets.JointTarget([Targets.RayleighDispersionPhase(periods, ysw + rng.normal(0, 0.01, 21)),
                              Targets.PReceiverFunction(xrf, yrf + rng.normal(0, 0.005, 201))])
    return jt, h, vp, vs


def _numpy_loglik(targets, noise):
    """The reference's accumulation (src/Targets.py:330-344) from the targets' current synthetics."""
    from bayhunter_b200.Targets import Valuation
    logL = 0.0
    for t, target in enumerate(targets):
        v = Valuation()
        corr, sigma = noise[2 * t], noise[2 * t + 1]
        c_inv, logdet = v.get_covariance_exp(corr, sigma, target.obsdata.y.size)
        logL += v.get_likelihood(target.obsdata.y, target.moddata.y, c_inv, logdet)
    return logL


def test_evaluate_forwards_kwargs_and_user_plugins(oracle):
    """JointTarget.evaluate forwards keyword arguments to the plugins like the reference (src/Targets.py:322-323)
    and accepts a user plugin (templates/myfwd.py contract) on a target; the likelihood stays on the device."""
    from bayhunter_b200 import Targets
    jt, h, vp, vs = _st3_joint(oracle)
    noise = np.array([0.0, 0.012, 0.6, 0.007])
    jt.evaluate(h, vp, vs, noise)
    base = jt.proposallikelihood
    assert abs(base - _numpy_loglik(jt.targets, noise)) <= 1e-9 * abs(base)
    # the same Q as the defaults, passed explicitly as per-layer arrays -> plugin path, same value
    jt.evaluate(h, vp, vs, noise, qp=np.full(4, 500.), qs=np.full(4, 225.))
    assert abs(jt.proposallikelihood - base) <= 1e-9 * abs(base)
    assert len(jt.proposalmisfits) == 3
    # other Q values change the receiver function, hence the likelihood
    jt.evaluate(h, vp, vs, noise, qp=np.full(4, 80.), qs=np.full(4, 30.))
    assert abs(jt.proposallikelihood - base) > 1e-3

    class MyForward(object):            # src/templates/myfwd.py
        def __init__(self, obsx, ref):
            self.obsx, self.ref = obsx, ref
        def set_modelparams(self, **kw):
            pass
        def run_model(self, h, vp, vs, rho, **kw):
            return self.obsx, np.cumsum(h)[0] * 0.01 + 0.1 * np.sin(self.obsx) * vs[0]

    x = np.linspace(0, 6, 40)
    mine = Targets.SingleTarget(x, 0.05 + 0.27 * np.sin(x) + 0.01, ref="myfwd")
    mine.noiseref = "swd"
    mine.update_plugin(MyForward(x, "myfwd"))
    jt2 = Targets.JointTarget(jt.targets + [mine])
    noise3 = np.concatenate((noise, [0.3, 0.02]))
    jt2.evaluate(h, vp, vs, noise3)
    want = _numpy_loglik(jt2.targets, noise3)
    assert abs(jt2.proposallikelihood - want) <= 1e-9 * abs(want)
    assert abs(jt2.proposalmisfits[-1] - sum(jt2.proposalmisfits[:-1])) < 1e-12
    # a plugin that fails (nan, nan) gives the sentinels
    mine.update_plugin(type("Bad", (MyForward,), {"run_model": lambda self, *a, **k: (np.nan, np.nan)})(x, "myfwd"))
    jt2.evaluate(h, vp, vs, noise3)
    assert jt2.proposallikelihood == -1e15 and list(jt2.proposalmisfits) == [1e15] * 4


def test_engine_cache_sees_rebound_laws(oracle):
    """A re-bound Gauss law (new fixed correlation -> new R^-1) must not reuse the cached engine."""
    jt, h, vp, vs = _st3_joint(oracle)
    rf = jt.targets[1]
    noise = np.array([0.0, 0.012, 0.9, 0.007])
    vals = []
    for corr in (0.9, 0.5):
        rf.valuation.init_covariance_gauss(corr, 201, rcond=1e-5)
        rf.get_covariance = rf.valuation.get_covariance_gauss
        jt.evaluate(h, vp, vs, noise)
        d = rf.moddata.y - rf.obsdata.y
        phi = d @ rf.valuation.corr_inv @ d / noise[3] ** 2
        want_rf = -0.5 * (201 * np.log(2 * np.pi) + 2 * 201 * np.log(noise[3]) + rf.valuation.logcorr_det) - 0.5 * phi
        sw = jt.targets[0]
        ds = sw.moddata.y - sw.obsdata.y
        want_sw = -0.5 * (21 * np.log(2 * np.pi) + 2 * 21 * np.log(noise[1])) - 0.5 * ds @ ds / noise[1] ** 2
        assert abs(jt.proposallikelihood - (want_rf + want_sw)) <= 1e-7 * abs(want_rf + want_sw), corr
        vals.append(jt.proposallikelihood)
    assert abs(vals[0] - vals[1]) > 1.0


def test_ensembles_own_their_engines_and_overflow_is_per_chain(oracle):
    """Two ensembles on one JointTarget (the second larger) and a single evaluate in between: the first keeps
    running (no engine is closed under a live sampler).  A chain whose arrays are full records WHEN that
    happened, and its dwell-time weights end there."""
    from bayhunter_b200 import SingleChain as sc
    jt, h, vp, vs = _st3_joint(oracle)
    priors = dict(vs=(2, 5), z=(0, 60), layers=(1, 8), vpvs=(1.4, 2.1), swdnoise_corr=0., swdnoise_sigma=(1e-5, 0.05),
                  rfnoise_corr=(0.3, 0.8), rfnoise_sigma=(1e-5, 0.05))
    ip = dict(iter_burnin=40, iter_main=60, thickmin=0.1, acceptance=(40, 45))
    a = sc.ChainEnsemble(jt, priors, ip, nchains=24, seed=5, max_accepted=6)
    a.init(); a.run(20)
    b = sc.ChainEnsemble(jt, priors, ip, nchains=64, seed=6, max_accepted=200)
    b.init(); b.run(5)
    jt.evaluate(h, vp, vs, np.array([0.0, 0.012, 0.6, 0.007]))
    a.run(80); b.run(95)
    sa, sb = a.state(), b.state()
    assert np.all(sa["iiter"] == 60) and np.all(sb["iiter"] == 60)
    assert sb["overflow_count"].sum() == 0
    over = sa["overflow_count"] > 0
    assert over.any() and sa["overflow"][0] == sa["overflow_count"].sum()
    arr = a.chain_arrays(0, 24)
    for c in np.where(over)[0]:
        n = int(sa["nstored"][c])
        assert n == 6
        last = int(arr["iters"][c, n - 1])
        assert last <= sa["overflow_iter"][c] <= 60
        one = {k: v[c] for k, v in arr.items()}
        fin = int(sa["overflow_iter"][c])
        w1 = sc.weighted_phase(one, n, 1, min(0, fin))
        w2 = sc.weighted_phase(one, n, 2, fin)
        total = (0 if w1 is None else w1.size) + (0 if w2 is None else w2.size)
        it = arr["iters"][c, :n].astype(int)
        if fin >= 0:
            assert total == fin - it[0]          # every iteration up to the first unstored model, no more
    a.close(); b.close()


def test_sampler_graph_replay_changes_nothing():
    """bh_sampler_run replays its iterations from CUDA graphs (captured between bh_engine_capture_begin / _end):
    the final state of a transdimensional run is bit-identical to the plainly enqueued one."""
    import os, subprocess, sys
    worker = os.path.join(os.path.dirname(os.path.abspath(__file__)), "sampler_graph_worker.py")
    out = {}
    for g in ("0", "1"):
        env = dict(os.environ, BH_SAMPLER_GRAPH=g)
        r = subprocess.run([sys.executable, worker], env=env, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        line = [l for l in r.stdout.splitlines() if l.startswith("DIGEST")][-1].split()
        out[g] = line[1:]
    assert out["0"] == out["1"], out
    assert int(out["1"][1]) == 300 and int(out["1"][2]) > 1000


@pytest.mark.parametrize("law,corr", [("exp", 0.85), ("gauss", 0.9), ("exp", 0.0)])
def test_device_noise_has_the_reference_distribution(law, corr):
    """bh_correlated_noise against the numpy recipe of SynthObs (src/SynthObs.py:136-155): same covariance
    (sample covariance of 40 000 draws within sampling error of sigma^2 R), zero mean, reproducible, and a
    realisation independent of the batch it is drawn in."""
    from bayhunter_b200.SynthObs import SynthObs, _lag_matrix
    n, sigma, B = 48, 0.0125, 40000
    x = SynthObs.device_noise(law, n, corr=corr, sigma=sigma, nreal=B, seed=7)
    lag = _lag_matrix(n)
    R = corr ** lag if law == "exp" else corr ** (lag ** 2)
    if corr == 0.0:
        R = np.eye(n)
    C = sigma ** 2 * R
    S = x.T @ x / B
    # sampling error of a covariance entry: sqrt((C_ii C_jj + C_ij^2) / B) <= sigma^2 sqrt(2 / B)
    assert np.abs(S - C).max() <= 6 * sigma ** 2 * np.sqrt(2.0 / B)
    assert np.abs(x.mean(axis=0)).max() <= 6 * sigma / np.sqrt(B)
    # the numpy recipe, same size of sample: its deviations from C are of the same size
    rs = np.random.RandomState(1)
    y = rs.multivariate_normal(np.zeros(n), C, size=B)
    assert np.abs(y.T @ y / B - C).max() <= 6 * sigma ** 2 * np.sqrt(2.0 / B)
    # reproducible and batch independent
    a = SynthObs.device_noise(law, n, corr=corr, sigma=sigma, nreal=5, seed=7)
    assert np.array_equal(a, x[:5])
    assert not np.array_equal(SynthObs.device_noise(law, n, corr=corr, sigma=sigma, nreal=5, seed=8), a)


def test_config4_full_size_properties_and_oracle_sample(oracle):
    """BASELINE config 4 at one GPU's full share (transd3, B = 4096, 3..31 rows): permutation equivariance and
    batch-split independence over the whole batch (ragged layer counts are dealt to warps sorted by depth and
    may run in two launches), and the oracle on a 384-model sample spread over the whole depth range."""
    eng, specs, otargets, rows, nlay, noise = _engine("transd3", 4096, 80)
    assert nlay.min() == 3 and nlay.max() == 31
    base = eng.eval_host(rows, nlay, noise, want_synth=True)
    perm = np.random.default_rng(5).permutation(4096)
    out = eng.eval_host(rows[perm], nlay[perm], noise[perm], want_synth=True)
    for a, b in zip(out, base):
        assert np.array_equal(a, b[perm], equal_nan=True)
    part = eng.eval_host(rows[1000:2500], nlay[1000:2500], noise[1000:2500], want_synth=True)
    for a, b in zip(part, base):
        assert np.array_equal(a, b[1000:2500], equal_nan=True)
    sel = np.argsort(nlay, kind="stable")[::11][:384]          # every 11th model by depth: 3 .. 31 rows
    ref = oracle.evaluate_batch(otargets, np.ascontiguousarray(rows[sel]), np.ascontiguousarray(nlay[sel]),
                                np.ascontiguousarray(noise[sel]))
    _compare(tuple(a[sel] for a in base), ref, [(s.ref, s.n) for s in specs])
    ok = ref[2] == 1
    e = _logl_check(base[0][sel], ref[0], ok)
    assert np.median(e) <= 1e-6 and (e <= 1e-6).mean() >= 0.95, (np.median(e), e.max())


def test_host_entry_edge_cases(oracle):
    """Empty batches, oversized batches, a likelihood-only call with several models and a Gauss-law target, and the
    link-level Fortran symbol on the non-default branches (higher mode, earth flattening)."""
    import bayhunter_b200 as bh
    from bayhunter_b200 import _lib
    from bayhunter_b200._lib import BayHunterB200Error
    eng, specs, otargets, rows, nlay, noise = _engine("swd2", 32, 90)
    T = eng.ntargets
    out = (np.empty(0), np.empty((0, T + 1)), np.empty(0, dtype=np.int32), None)
    t = eng.submit_host(rows[:0], nlay[:0], noise[:0], out)
    assert t == 0
    eng.wait(t)
    big = np.zeros((33, rows.shape[1], 4))
    with pytest.raises(BayHunterB200Error):
        eng.eval_host(big, np.zeros(33, np.int32), np.zeros((33, 2 * T)))
    # likelihood-only entry: three models, exponential + Gauss law, one target rejected in the last model
    x1, x2 = np.linspace(1, 30, 17), -2.0 + 0.25 * np.arange(70)
    rng = np.random.default_rng(1)
    y1, y2 = 3.0 + 0.02 * x1, rng.normal(0, 0.1, 70)
    ci, ld = bh.gauss_corr_inverse(0.8, 70, rcond=1e-5)
    gen = bh.Engine([bh.TargetSpec("anything", x1, y1, cov="exp", generic=True),
                     bh.TargetSpec("prf", x2, y2, cov="gauss", corr_inv=ci, logcorr_det=ld, generic=True)], 4, 2)
    synth = np.concatenate((y1 + rng.normal(0, 0.02, (3, 17)), y2 + rng.normal(0, 0.05, (3, 70))), axis=1)
    nz = np.tile([0.3, 0.02, 0.8, 0.05], (3, 1))
    tv = np.ones((3, 2), dtype=np.int32); tv[2, 1] = 0
    logL, mis, stat = gen.loglik_host(synth, tv, nz)
    assert stat.tolist() == [1, 1, 0] and logL[2] == -1e15 and np.all(mis[2] == 1e15)
    from bayhunter_b200.Targets import Valuation
    v = Valuation()
    for b in range(2):
        c1, d1 = v.get_covariance_exp(0.3, 0.02, 17)
        want = v.get_likelihood(y1, synth[b, :17], c1, d1) + \
            v.get_likelihood(y2, synth[b, 17:], ci / 0.05 ** 2, 2 * 70 * np.log(0.05) + ld)
        assert abs(logL[b] - want) <= 1e-9 * abs(want)
        assert abs(mis[b, 0] - np.sqrt(np.mean((synth[b, :17] - y1) ** 2))) < 1e-14
    with pytest.raises(BayHunterB200Error):          # a generic target has no forward model in the engine
        gen.eval_host(np.zeros((1, 2, 4)), np.array([2], np.int32), nz[:1])
    # surfdisp96_ on the general branches
    lib = ctypes.CDLL(_lib.library_path())
    h = np.array([5., 23., 8., 0.]); vs = np.array([2.7, 3.6, 3.8, 4.4]); vp = vs * 1.73; rho = vp * 0.32 + 0.77
    periods = np.linspace(2, 20, 12)
    fp = ctypes.POINTER(ctypes.c_float); dp = ctypes.POINTER(ctypes.c_double)
    for mode, flsph in ((2, 0), (1, 1)):
        arr = [np.zeros(100, dtype=np.float32) for _ in range(4)]
        for a, val in zip(arr, (h, vp, vs, rho)):
            a[:4] = val
        tt = np.zeros(60); tt[:12] = periods
        cg = np.zeros(60)
        ints = [ctypes.c_int(x) for x in (4, flsph, 2, mode, 0, 12)]
        err = ctypes.c_int(-1)
        lib.surfdisp96_.restype = None
        lib.surfdisp96_(*[a.ctypes.data_as(fp) for a in arr], *[ctypes.byref(i) for i in ints],
                        tt.ctypes.data_as(dp), cg.ctypes.data_as(dp), ctypes.byref(err))
        xo, yo = oracle.surfdisp(h, vp, vs, rho, "rdispph", periods, mode=mode, flsph=flsph)
        assert err.value == 0
        assert _rel(cg[:12][yo > 0], yo[yo > 0]).max() <= 1e-6
