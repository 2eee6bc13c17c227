"""oracle/first_principles.py -- TEST INFRASTRUCTURE.  An independent statement of the physics the
SURF96 path solves, used to pin the dispersion oracle beyond the 4 decimals of the reference's
fixtures (tests/test_oracle.py).  It shares nothing with surfdisp96.f: no Dunkin/Haskell layer
matrices, no root bracketing logic, no REAL*4.

The surface-wave eigenproblem of a stack of homogeneous layers over a half-space is the first-order
system for the motion-stress vector (Aki & Richards, Quantitative Seismology, eqs. 7.24 / 7.28):

  Love      d/dz (l1, l2)   = [[0, 1/mu], [k^2 mu - w^2 rho, 0]] (l1, l2)
  Rayleigh  d/dz (r1..r4)   = [[0, k, 1/mu, 0],
                               [-k lam/(lam+2mu), 0, 0, 1/(lam+2mu)],
                               [k^2 zeta - w^2 rho, 0, 0, k lam/(lam+2mu)],
                               [0, -w^2 rho, -k, 0]] (r1..r4),   zeta = 4 mu (lam+mu)/(lam+2mu)

with decaying solutions in the half-space and vanishing tractions (l2; r3, r4) at the free surface.
Here the decaying half-space solutions (eigenvectors of the half-space matrix with Re(eigenvalue) < 0)
are carried up through every layer with the matrix exponential exp(-A d) in multi-precision
arithmetic (mpmath), and the surface tractions give the secular function whose zeros in c = w/k are
the phase velocities.  Only pure-Python/mpmath: small cases only.
"""
import mpmath as mp
import numpy as np


def _f32(a):
    """The model SURF96 sees: REAL*4 values (f2py cast), carried exactly into mpmath."""
    return [mp.mpf(float(v)) for v in np.asarray(a, np.float32).astype(np.float64)]


def love_secular(h, vs, rho, c, period):
    h, vs, rho = _f32(h), _f32(vs), _f32(rho)
    w = 2 * mp.pi / mp.mpf(period)
    k = w / mp.mpf(c)
    mu = [r * b * b for r, b in zip(rho, vs)]
    nu = mp.sqrt(k * k - w * w * rho[-1] / mu[-1])
    y = mp.matrix([1, -mu[-1] * nu])
    for i in range(len(h) - 2, -1, -1):
        A = mp.matrix([[0, 1 / mu[i]], [k * k * mu[i] - w * w * rho[i], 0]])
        y = mp.expm(-A * h[i]) * y
    return mp.re(y[1])


def rayleigh_secular(h, vp, vs, rho, c, period):
    h, vp, vs, rho = _f32(h), _f32(vp), _f32(vs), _f32(rho)
    w = 2 * mp.pi / mp.mpf(period)
    k = w / mp.mpf(c)

    def A_of(i):
        mu = rho[i] * vs[i] ** 2
        lam = rho[i] * vp[i] ** 2 - 2 * mu
        zeta = 4 * mu * (lam + mu) / (lam + 2 * mu)
        return mp.matrix([[0, k, 1 / mu, 0],
                          [-k * lam / (lam + 2 * mu), 0, 0, 1 / (lam + 2 * mu)],
                          [k * k * zeta - w * w * rho[i], 0, 0, k * lam / (lam + 2 * mu)],
                          [0, -w * w * rho[i], -k, 0]])
    n = len(h)
    E, V = mp.eig(A_of(n - 1))
    idx = sorted([i for i in range(4) if mp.re(E[i]) < 0], key=lambda i: mp.re(E[i]))
    if len(idx) != 2:
        raise ValueError("c is not below the half-space shear velocity")
    Y = mp.matrix(4, 2)
    for j, i in enumerate(idx):
        # eigenvectors come with an arbitrary (complex) scale: fix it (r1 = 1) and the column order
        # (by decay rate), so that the determinant below is a continuous real function of c
        for r in range(4):
            Y[r, j] = mp.re(V[r, i] / V[0, i])
    for i in range(n - 2, -1, -1):
        Y = mp.expm(-A_of(i) * h[i]) * Y
    return mp.re(Y[2, 0] * Y[3, 1] - Y[2, 1] * Y[3, 0])


def exact_root(fn, c0, rel=2e-5, tol=1e-12):
    """Zero of fn in [c0 (1 - rel), c0 (1 + rel)] by bisection/Illinois; None without a sign change."""
    a, b = mp.mpf(c0) * (1 - rel), mp.mpf(c0) * (1 + rel)
    fa, fb = fn(a), fn(b)
    if fa * fb > 0:
        return None
    for _ in range(80):
        m = (a * fb - b * fa) / (fb - fa)
        if not (a < m < b):
            m = (a + b) / 2
        fm = fn(m)
        if fm == 0:
            return m
        if fa * fm < 0:
            b, fb = m, fm
            fa = fa / 2
        else:
            a, fa = m, fm
            fb = fb / 2
        if (b - a) <= tol * a:
            break
    return (a + b) / 2
