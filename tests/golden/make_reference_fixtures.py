"""tests/golden/make_reference_fixtures.py -- generate golden vectors by RUNNING THE REFERENCE.

Run once in the build container (the reference tree cannot travel to the GPU box):

    python tests/golden/make_reference_fixtures.py

Writes small .npz fixtures next to this script.  What produced every number:
the reference's own Python (Models.py, Targets.py, SingleChain.py,
surf96_modsw.py, rfmini_modrf.py, imported unmodified through refshim.py) on top of
  * the reference's own rfmini C++ (oracle/_ref/librfmini_ref.so) for receiver functions,
  * oracle/surf96_oracle.c for dispersion (no Fortran compiler in this image) -- so
    dispersion VALUES in these fixtures are only as pinned as that restatement
    (tests/golden/st3_*disp*.dat, 4 decimals); everything downstream of the forward
    values (validity, misfit, covariance laws, log-likelihood, sentinels, the
    Voronoi -> layer adapter) is the reference's code.

Fixtures:
  ref_models.npz      Model.get_vp_vs_h (src/Models.py:40-52) on random Voronoi models,
                      with and without the `mantle` prior
  ref_joint_eval.npz  JointTarget.evaluate (src/Targets.py:314-347) with the covariance law
                      bound by SingleChain.set_target_covariance (src/SingleChain.py:159-205),
                      for four noise set-ups: exp / white / white-scaled / gauss
"""
import os
import sys
import types
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
warnings.simplefilter("ignore")
import refshim  # noqa: E402

BH = refshim.import_reference()
from BayHunter import Model  # noqa: E402
from BayHunter import Targets  # noqa: E402
from BayHunter.SingleChain import SingleChain  # noqa: E402

ST3 = dict(h=np.array([5., 23., 8., 0.]), vs=np.array([2.7, 3.6, 3.8, 4.4]), vpvs=1.73)


def draw_voronoi(rng, k, zmax=60.0, thickmin=0.1):
    while True:
        vs = np.sort(rng.uniform(2.0, 5.0, k))
        z = np.sort(rng.uniform(0.0, zmax, k))
        model = np.concatenate((vs, z))
        vp, vs_, h = Model.get_vp_vs_h(model, 1.73, None)
        if k == 1 or np.all(h[:-1] > thickmin):
            return model


def make_models():
    rng = np.random.default_rng(20260101)
    kmax = 21
    N = 96
    models = np.full((N, 2 * kmax), np.nan)
    vpvs = rng.uniform(1.4, 2.1, N)
    nrow = np.zeros(N, dtype=np.int32)
    use_mantle = np.zeros(N, dtype=np.int32)
    mantle = (4.2, 1.8)
    out_h = np.full((N, kmax), np.nan)
    out_vp = np.full((N, kmax), np.nan)
    out_vs = np.full((N, kmax), np.nan)
    for i in range(N):
        k = int(rng.integers(1, kmax + 1))
        m = draw_voronoi(rng, k)
        if i % 7 == 3:           # unsorted velocities (low-velocity zones) exercise the mantle rule
            m[:k] = rng.permutation(m[:k])
        models[i, :2 * k] = m
        nrow[i] = k
        use_mantle[i] = i % 3 == 0
        vp, vs, h = Model.get_vp_vs_h(m, vpvs[i], mantle if use_mantle[i] else None)
        out_h[i, :k], out_vp[i, :k], out_vs[i, :k] = h, vp, vs
    np.savez_compressed(os.path.join(HERE, "ref_models.npz"), models=models, vpvs=vpvs, nrow=nrow,
                        use_mantle=use_mantle, mantle=np.array(mantle), h=out_h, vp=out_vp, vs=out_vs)
    print("ref_models.npz: %d models" % N)


def build_targets(rng, with_yerr):
    """Five targets observed on the st3 truth model + noise (like tutorial/create_testdata.py)."""
    h, vs = ST3["h"], ST3["vs"]
    vp = vs * ST3["vpvs"]
    rho = vp * 0.32 + 0.77
    periods = np.linspace(1, 40, 20)
    x_rf = -5.0 + 0.2 * np.arange(201)
    classes = [("rdispph", Targets.RayleighDispersionPhase), ("rdispgr", Targets.RayleighDispersionGroup),
               ("ldispph", Targets.LoveDispersionPhase), ("ldispgr", Targets.LoveDispersionGroup),
               ("prf", Targets.PReceiverFunction)]
    targets, obs = [], []
    for ref, cls in classes:
        x = x_rf if ref == "prf" else periods
        probe = cls(x, np.zeros(x.size))
        _, y = probe.moddata.plugin.run_model(h, vp, vs, rho)
        y = y + rng.normal(0, 0.01 if ref == "prf" else 0.02, y.size)
        yerr = None
        if with_yerr and ref != "prf":
            yerr = rng.uniform(0.01, 0.05, y.size)
        t = cls(x, y, yerr=yerr)
        targets.append(t)
        obs.append((ref, x, y, yerr))
    return Targets.JointTarget(targets=targets), obs


CASES = {
    # name: (with_yerr, swd corr prior, rf corr prior)  -- a float prior is "fixed"
    "exp": (False, (0.0, 0.5), (0.35, 0.75)),          # every corr sampled -> exponential law
    "white": (False, 0.0, 0.0),                        # fixed 0, no yerr   -> get_covariance_nocorr
    "white_scaled": (True, 0.0, (0.35, 0.75)),         # fixed 0, yerr      -> nocorr_scalederr (swd); exp (rf)
    "gauss": (False, 0.3, 0.92),                       # fixed non-zero     -> exp (swd), gauss (rf)
}
RCOND = 1e-5


def make_joint():
    out = {}
    for ci, (name, (with_yerr, swdc, rfc)) in enumerate(CASES.items()):
        rng = np.random.default_rng(777 + ci)
        joint, obs = build_targets(rng, with_yerr)
        T = len(joint.targets)
        corrfix = np.zeros(T, dtype=bool)
        corr0 = np.zeros(T)
        for i, t in enumerate(joint.targets):
            prior = rfc if t.noiseref == "rf" else swdc
            corrfix[i] = not isinstance(prior, tuple)
            corr0[i] = prior if corrfix[i] else 0.5
        # the chain's own binding code (SingleChain.py:159-205), run on a bare namespace
        SingleChain.set_target_covariance(types.SimpleNamespace(targets=joint), corrfix, corr0, RCOND)
        laws = [t.get_covariance.__name__ for t in joint.targets]
        B, lmax = 40, 9
        H = np.zeros((B, lmax)); VP = np.zeros((B, lmax)); VS = np.zeros((B, lmax))
        nlay = np.zeros(B, dtype=np.int32)
        noise = np.zeros((B, 2 * T))
        logL = np.zeros(B); misfits = np.zeros((B, T + 1))
        nsyn = sum(o[1].size for o in obs)
        synth = np.full((B, nsyn), np.nan)
        for b in range(B):
            k = int(rng.integers(2, lmax + 1))
            m = draw_voronoi(rng, k)
            if b == 5:
                m[:k] = 0.5 * m[:k]                      # very slow model
            vpvs = rng.uniform(1.5, 2.0)
            vp, vs, h = Model.get_vp_vs_h(m, vpvs, None)
            if b == 7:                                   # a model SURF96 cannot solve -> sentinels
                vs = vs.copy(); vs[0] = 0.4; vs[1:] = np.linspace(6.5, 9.0, k - 1); vp = vs * 1.2
            for i, t in enumerate(joint.targets):
                prior = rfc if t.noiseref == "rf" else swdc
                noise[b, 2 * i] = prior if corrfix[i] else rng.uniform(*prior)
                noise[b, 2 * i + 1] = rng.uniform(0.005, 0.05)
            joint.evaluate(h=h, vp=vp, vs=vs, noise=noise[b])
            H[b, :k], VP[b, :k], VS[b, :k], nlay[b] = h, vp, vs, k
            logL[b] = joint.proposallikelihood
            misfits[b] = joint.proposalmisfits
            o = 0
            for t in joint.targets:
                n = t.obsdata.y.size
                if isinstance(t.moddata.y, np.ndarray) and t.moddata.y.size == n:
                    synth[b, o:o + n] = t.moddata.y
                o += n
        out[name + "_h"] = H; out[name + "_vp"] = VP; out[name + "_vs"] = VS; out[name + "_nlay"] = nlay
        out[name + "_noise"] = noise; out[name + "_logL"] = logL; out[name + "_misfits"] = misfits
        out[name + "_synth"] = synth.astype(np.float64)
        out[name + "_laws"] = np.array(laws)
        out[name + "_refs"] = np.array([o[0] for o in obs])
        for i, (ref, x, y, yerr) in enumerate(obs):
            out["%s_obs%d_x" % (name, i)] = x
            out["%s_obs%d_y" % (name, i)] = y
            if yerr is not None:
                out["%s_obs%d_yerr" % (name, i)] = yerr
        nbad = int(np.sum(logL <= -1e14))
        print("%-13s laws=%s  invalid models=%d  logL range [%.1f, %.1f]"
              % (name, laws, nbad, logL[logL > -1e14].min(), logL.max()))
    out["rcond"] = np.array(RCOND)
    np.savez_compressed(os.path.join(HERE, "ref_joint_eval.npz"), **out)


if __name__ == "__main__":
    make_models()
    make_joint()
