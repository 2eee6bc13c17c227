"""CPU tests of the device cores compiled for the host (tests/host_sim): the search
state machine must reproduce the oracle's root sequence for ANY speculation width,
and the receiver-function algebra must agree with the oracle to rounding."""
import ctypes

import numpy as np
import pytest

F = ctypes.POINTER(ctypes.c_float)
D = ctypes.POINTER(ctypes.c_double)


def _sim_curve(sim, h, vp, vs, rho, wave, igr, t, mode, seed=1):
    rows = np.ascontiguousarray(np.stack([h, vp, vs, rho], 1), dtype=np.float32)
    cg = np.zeros(len(t))
    cnt = (ctypes.c_longlong * 2)()
    err = sim.swd_sim_curve(rows.ctypes.data_as(F), len(h), wave, igr, len(t), t.ctypes.data_as(D),
                            mode, seed, cg.ctypes.data_as(D), cnt)
    return cg, err, cnt[0], cnt[1]


def _models(rng, n, kmax=9):
    from bayhunter_b200 import synthetic
    for it in range(n):
        k = int(rng.integers(1, kmax))
        h, vs = synthetic.draw_model(rng, k)
        vp = vs * rng.uniform(1.4, 2.1)
        yield it, h, vp, vs, vp * 0.32 + 0.77


def test_search_state_machine_equals_oracle_for_any_speculation(oracle, host_sim_reforder):
    """Reference-order arithmetic + state machine == oracle, bit for bit, whatever the
    number of speculative bracket candidates per round."""
    t = np.linspace(1, 40, 30)
    nfail = 0
    for it, h, vp, vs, rho in _models(np.random.default_rng(1), 120):
        for ref, (wave, igr) in oracle.SURFTAGS.items():
            cnt = [0, 0]
            xo, yo = oracle.surfdisp(h, vp, vs, rho, ref, t, count=cnt)
            for mode in (0, 5, -1):        # reference order / fixed width / random width per round
                y, err, consumed, evaluated = _sim_curve(host_sim_reforder, h, vp, vs, rho, wave, igr, t, mode, seed=it)
                if not isinstance(xo, np.ndarray):
                    assert err == 1
                    nfail += mode == 0
                    continue
                assert err == 0
                assert np.array_equal(y, yo), (it, ref, mode)
                assert consumed == sum(cnt)
                assert evaluated >= consumed
    assert nfail > 0


def _sim_general(sim, h, vp, vs, rho, wave, igr, t, mode, flsph):
    rows = np.ascontiguousarray(np.stack([h, vp, vs, rho], 1), dtype=np.float32)
    cg = np.zeros(len(t))
    n = ctypes.c_longlong(0)
    err = sim.swd_sim_general(rows.ctypes.data_as(F), len(h), wave, igr, len(t), mode, flsph,
                              t.ctypes.data_as(D), cg.ctypes.data_as(D), ctypes.byref(n))
    return cg, err, n.value


def _oracle_raw(oracle, h, vp, vs, rho, wave, igr, t, mode, flsph):
    """surf96_oracle with the raw (cg, err) outputs: higher modes that do not exist leave
    zeros in cg with err = 0 (surfdisp96.f:350-354), which SurfDisp.run_model returns as is."""
    arrs = [np.ascontiguousarray(a, dtype=np.float32) for a in (h, vp, vs, rho)]
    cg = np.zeros(len(t))
    err = ctypes.c_int(0)
    ns = (ctypes.c_long * 2)()
    oracle.lib().surf96_oracle(*[a.ctypes.data_as(F) for a in arrs], len(h), flsph, wave, mode, igr,
                               len(t), t.ctypes.data_as(D), cg.ctypes.data_as(D), ctypes.byref(err), ns)
    return cg, err.value, ns[0] + ns[1]


def test_general_path_modes_sphere_water_equal_oracle(oracle, host_sim_reforder):
    """swd_general_core.cuh (what one thread of swd_general_kernel runs) against the oracle:
    fundamental + higher modes, flat / earth-flattened, solid stacks and a water layer on
    top.  Same candidate sequence (equal evaluation counts) and bit-equal curves, except the
    Rayleigh density power of the sphere transform (powf vs pow rounded once: <= 1e-7)."""
    t = np.linspace(1, 40, 30)
    rng = np.random.default_rng(11)
    n_zero = n_water = n_hi = 0
    for it, h, vp, vs, rho in _models(rng, 60, kmax=12):
        water = it % 4 == 1 and len(h) >= 3
        if water:
            h = h.copy(); vp = vp.copy(); vs = vs.copy(); rho = rho.copy()
            h[0] = rng.uniform(0.5, 4.0); vs[0] = 0.0; vp[0] = 1.5; rho[0] = 1.03
            n_water += 1
        for ref, (wave, igr) in oracle.SURFTAGS.items():
            for mode in (1, 2, 3):
                for flsph in (0, 1):
                    yo, erro, no = _oracle_raw(oracle, h, vp, vs, rho, wave, igr, t, mode, flsph)
                    y, err, n = _sim_general(host_sim_reforder, h, vp, vs, rho, wave, igr, t, mode, flsph)
                    assert err == erro, (it, ref, mode, flsph)
                    if err:
                        continue
                    if flsph == 1 and wave == 2:
                        ok = yo != 0
                        assert np.array_equal(ok, y != 0)
                        assert np.abs(y[ok] - yo[ok]).max(initial=0) <= 1e-7 * np.abs(yo).max()
                    else:
                        assert np.array_equal(y, yo), (it, ref, mode, flsph, water)
                        assert n == no
                    n_zero += int((yo == 0).any())
                    n_hi += int(mode > 1 and (yo != 0).any())
    assert n_water > 5 and n_zero > 0 and n_hi > 20


def test_device_formulation_of_secular_functions(oracle, host_sim):
    """The branch-free / reciprocal formulation the kernels use (libm-backed on the host)
    must give the same curves as the oracle to rounding: phase <= 1e-9, group <= 5e-5,
    identical failure flags, and the same number of consumed secular values almost always."""
    t = np.linspace(1, 40, 30)
    same_count = total = 0
    for it, h, vp, vs, rho in _models(np.random.default_rng(7), 150):
        for ref, (wave, igr) in oracle.SURFTAGS.items():
            cnt = [0, 0]
            xo, yo = oracle.surfdisp(h, vp, vs, rho, ref, t, count=cnt)
            y, err, consumed, evaluated = _sim_curve(host_sim, h, vp, vs, rho, wave, igr, t, 3, seed=it)
            if not isinstance(xo, np.ndarray):
                assert err == 1
                continue
            assert err == 0
            rel = np.abs(y - yo) / yo
            assert rel.max() <= (1e-9 if igr == 0 else 5e-5), (it, ref, rel.max())
            total += 1
            same_count += consumed == sum(cnt)
    assert same_count >= 0.98 * total


def test_lane_dealing_closed_form(host_sim):
    """deal_lanes (popcounts of the two ballots) against the sequential definition: one lane per refining
    chain (3, 7, 15 or 31 when lanes are left over even with four per walking chain), the spare lanes dealt evenly to the walking chains in lane order (the first `extra % nbr` of them
    one more), at most max_spec per chain; runs laid out in lane order without gaps."""
    rng = np.random.default_rng(5)
    I32 = ctypes.c_int * 32
    for trial in range(20000):
        p0, p1 = rng.random(2)
        want = np.where(rng.random(32) < p0, 0, np.where(rng.random(32) < p1, 32, 1))
        if trial < 40:                                    # corner cases: all walking / all refining / one lane
            want = [np.full(32, 32), np.full(32, 1), np.eye(32, dtype=int)[trial % 32] * 32,
                    np.eye(32, dtype=int)[trial % 32]][trial % 4]
        if not want.any():
            continue
        max_spec = int(rng.choice([1, 2, 3, 4, 8, 16, 32]))
        active = int(sum(1 << l for l in range(32) if want[l] > 0))
        bracket = int(sum(1 << l for l in range(32) if want[l] > 1))
        nact, nbr = int((want > 0).sum()), int((want > 1).sum())
        nrf = nact - nbr
        # lanes left over even with four per walking chain: two guess lanes per refining chain
        room = 32 - nact - 3 * nbr
        g = 0 if nrf == 0 else next((t for t in (30, 14, 6, 2) if t * nrf <= room), 0)
        extra = 32 - nact - g * nrf
        cnt_ref = []
        for l in range(32):
            c = 0
            if want[l] > 0:
                c = 1 + g
                if want[l] > 1:
                    rank = int((want[:l] > 1).sum())
                    c = min(1 + extra // nbr + (1 if rank < extra % nbr else 0), max_spec)
            cnt_ref.append(c)
        excl_ref = np.concatenate(([0], np.cumsum(cnt_ref)[:-1]))
        cnt, excl, total = I32(), I32(), I32()
        host_sim.swd_sim_deal(ctypes.c_uint(active), ctypes.c_uint(bracket), max_spec, cnt, excl, total)
        assert list(cnt) == cnt_ref, (want, max_spec)
        own = [l for l in range(32) if cnt_ref[l] > 0]
        assert [excl[l] for l in own] == [int(excl_ref[l]) for l in own]
        assert set(total) == {sum(cnt_ref)} and sum(cnt_ref) <= 32


def test_secular_values_of_the_device_formulation(host_sim, host_sim_reforder):
    """The secular functions as the kernels evaluate them -- records of reciprocals, branch-free half
    terms, Rayleigh layers applied through the rank-one structure of Dunkin's matrix without forming it
    (dunkin_apply_factored), power-of-two rescaling -- against the Fortran-order functions on the same
    models, periods and trial velocities: normalised values equal to rounding, signs equal wherever the
    value is not itself at rounding level."""
    from bayhunter_b200 import synthetic
    rng = np.random.default_rng(21)
    worst = {1: 0.0, 2: 0.0}
    nsign = 0
    for it in range(400):
        k = int(rng.integers(1, 31))
        h, vs = synthetic.draw_model(rng, k)
        if it % 5 == 0:
            vs = rng.permutation(vs)                   # low-velocity zones
        vp = vs * rng.uniform(1.4, 2.1)
        rows = np.ascontiguousarray(np.stack([h, vp, vs, vp * 0.32 + 0.77], 1), dtype=np.float32)
        omega = 2 * np.pi / 10 ** rng.uniform(-0.3, 1.8)
        c = np.ascontiguousarray(rng.uniform(0.8 * vs.min(), 1.001 * vs.max(), 16))
        for wave in (1, 2):
            a = np.zeros(c.size); b = np.zeros(c.size)
            for sim, out in ((host_sim, a), (host_sim_reforder, b)):
                sim.swd_sim_secular(rows.ctypes.data_as(F), len(h), wave, ctypes.c_double(omega), c.size,
                                    c.ctypes.data_as(D), out.ctypes.data_as(D))
            assert np.isfinite(a).all() and np.isfinite(b).all()
            d = np.abs(a - b)
            worst[wave] = max(worst[wave], d.max())
            far = np.abs(b) > 1e-9
            nsign += int((np.sign(a[far]) != np.sign(b[far])).sum())
    # values are normalised to max|e| = 1; 30-layer stacks accumulate a few hundred roundings
    assert worst[1] <= 1e-10 and worst[2] <= 1e-10, worst
    assert nsign == 0


def test_rf_core_equals_oracle(oracle, host_sim):
    from bayhunter_b200 import synthetic
    at = [ctypes.c_int] + [ctypes.c_double] * 6 + [ctypes.c_int] * 2 + [D] * 7
    host_sim.rf_sim.argtypes = at
    rng = np.random.default_rng(2)
    for it in range(40):
        k = int(rng.integers(2, 32))
        h, vs = synthetic.draw_model(rng, k)
        vp = vs * rng.uniform(1.4, 2.1)
        rho = vp * 0.32 + 0.77
        z = np.ascontiguousarray(np.concatenate(([0], np.cumsum(h)[:-1])))
        qp = np.ones(k) * 500.; qs = np.ones(k) * 225.
        kk = vp[0] / vs[0]
        poisson = (2 - kk * kk) / (2 - 2 * kk * kk)
        for wave, (nsamp, fsamp) in ((0, (512, 5.)), (1, (1024, 10.))):
            a = np.zeros(nsamp); b = np.zeros(nsamp)
            args = (nsamp, fsamp, 5.0, 6.4, 1.0, float(vs[0]), poisson, wave, k)
            ptrs = lambda out: [x.ctypes.data_as(D) for x in (z, vp, vs, rho, qp, qs, out)]
            oracle.lib().rf_oracle(*args, *ptrs(a))
            host_sim.rf_sim(*args, *ptrs(b))
            assert np.abs(a - b).max() <= 1e-12 * np.abs(a).max()
