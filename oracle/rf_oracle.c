/*
 * oracle/rf_oracle.c  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (C99, <complex.h>) of the receiver-function forward path of
 * the reference's rfmini: synrf_cwrap (wrap.cpp:57-80) -> synrf (synrf.cpp:16-55)
 * -> FlatLayer::flatten (model.cpp:223-251) -> calcresp/calcresp_core
 * (greens.cpp:400-683, 685-756) -> compute_rf (greens.cpp:343-398) -> iftr
 * (greens.cpp:136-158) -> ccfork (fork.cpp:10-60).
 *
 * It is pinned two ways: (1) against the reference's own C++ compiled from
 * /root/reference into oracle/_ref/librfmini_ref.so (see oracle/Makefile,
 * tests/test_oracle_rf.py) and (2) against the reference fixtures
 * tutorial/observed/st3_prf.dat, st3_srf.dat (4 decimals).  It exists so that
 * the GPU box, where /root/reference is absent, still has a complete checker
 * in source form.  Only tests/, __graft_entry__.smoke() and bench.py's CPU
 * baseline legs may load it; the CUDA product path never does.
 *
 * Not restated (dead in BayHunter's call): partial derivatives (drdp == NULL,
 * synrf.cpp:52), bottom_up, the SH block (greens.cpp:553-560, result unused),
 * the Z/R time traces (iftr2; BayHunter discards fz, fr).
 */
#include <complex.h>
#include <math.h>
#include <stdlib.h>

typedef double complex cplx;
typedef struct { cplx c11, c12, c21, c22; } cmat2;

#define RF_EARTH_RADIUS 6371.0          /* model.cpp:220 */
#define RF_DEGREES_PER_KM 0.00899       /* wrap.cpp:55 */

static cmat2 m_mul(cmat2 x, cmat2 y)    /* cmat2.h:172-179 */
{
    cmat2 r;
    r.c11 = x.c11 * y.c11 + x.c12 * y.c21;
    r.c12 = x.c11 * y.c12 + x.c12 * y.c22;
    r.c21 = x.c21 * y.c11 + x.c22 * y.c21;
    r.c22 = x.c21 * y.c12 + x.c22 * y.c22;
    return r;
}
static cmat2 m_add(cmat2 x, cmat2 y)
{
    cmat2 r = { x.c11 + y.c11, x.c12 + y.c12, x.c21 + y.c21, x.c22 + y.c22 };
    return r;
}
static cmat2 m_sub(cmat2 x, cmat2 y)
{
    cmat2 r = { x.c11 - y.c11, x.c12 - y.c12, x.c21 - y.c21, x.c22 - y.c22 };
    return r;
}
static cmat2 m_inv(cmat2 x)             /* cmat2.h:143-152 */
{
    cplx q = 1.0 / (x.c11 * x.c22 - x.c12 * x.c21);
    cmat2 r = { q * x.c22, -q * x.c12, -q * x.c21, q * x.c11 };
    return r;
}
static cmat2 m_exe(cmat2 e, cmat2 x)    /* greens.cpp:829-845: e*x*e, e diagonal */
{
    cplx e12 = e.c11 * e.c22;
    cplx e11 = e.c11 * e.c11;
    cplx e22 = e.c22 * e.c22;
    cmat2 r = { x.c11 * e11, x.c12 * e12, x.c21 * e12, x.c22 * e22 };
    return r;
}

typedef struct { double z, h, vp, vs, rh, qp, qs; } layer_t;

/* model.cpp:223-251 (isLowerHalfspace: model.cpp:209-219) */
static void flatten_layer(layer_t *l)
{
    double zb = l->z + l->h;
    double r = RF_EARTH_RADIUS - l->z;
    double q = RF_EARTH_RADIUS / r;
    l->z = RF_EARTH_RADIUS * log(q);
    l->vp *= q;
    l->vs *= q;
    l->rh /= q;
    int lower_halfspace = !(l->h > 0.) && !(l->vp < 1. && l->rh < 0.1);
    if (!lower_halfspace) {
        r = RF_EARTH_RADIUS - zb;
        q = RF_EARTH_RADIUS / r;
        zb = RF_EARTH_RADIUS * log(q);
        l->h = zb - l->z;
    }
}

/* greens.cpp:19-85 (P-SV part only) */
static void coeffm(double u, double vp1, double vs1, double rho1,
                   double vp2, double vs2, double rho2,
                   cmat2 *rd, cmat2 *td, cmat2 *ru, cmat2 *tu)
{
    double mue1 = rho1 * vs1 * vs1, mue2 = rho2 * vs2 * vs2,
           c = 2. * (mue1 - mue2), u2 = u * u, cu2 = c * u2, t1, t2, t3;
    cplx rpp, rps, rsp, rss, tpp, tps, tsp, tss, d1, d2, t4, t5, t7;
    cplx a1 = conj(csqrt(1. / (vp1 * vp1) - u2 + 0.0 * I));
    cplx a2 = conj(csqrt(1. / (vp2 * vp2) - u2 + 0.0 * I));
    cplx b1 = conj(csqrt(1. / (vs1 * vs1) - u2 + 0.0 * I));
    cplx b2 = conj(csqrt(1. / (vs2 * vs2) - u2 + 0.0 * I));

    t1 = cu2 - rho1 + rho2;
    t2 = cu2 - rho1;
    t3 = cu2 + rho2;
    t4 = t3 * a1 - t2 * a2;

    d1 = t1 * t1 * u2 + t2 * t2 * a2 * b2 + rho1 * rho2 * a2 * b1;
    d2 = c * c * u2 * a1 * a2 * b1 * b2 + t3 * t3 * a1 * b1 + rho1 * rho2 * a1 * b2;
    t5 = 1. / (d1 + d2);
    t7 = 2. * rho1 * t5;

    rpp = (d2 - d1) * t5;
    rps = -2. * u * a1 * t5 * (t1 * t3 + c * t2 * a2 * b2);
    tpp = a1 * t7 * (t3 * b1 - t2 * b2);
    tps = -a1 * t7 * u * (t1 + c * a2 * b1);
    rss = (d2 - d1 - 2. * rho1 * rho2 * (a1 * b2 - a2 * b1)) * t5;
    rsp = 2. * u * b1 * t5 * (t1 * t3 + c * t2 * a2 * b2);
    tss = b1 * t7 * t4;
    tsp = b1 * t7 * u * (t1 + c * a1 * b2);
    rd->c11 = rpp; rd->c12 = rsp; rd->c21 = rps; rd->c22 = rss;
    td->c11 = tpp; td->c12 = tsp; td->c21 = tps; td->c22 = tss;

    d1 = t1 * t1 * u2 + t3 * t3 * a1 * b1 + rho1 * rho2 * a1 * b2;
    d2 = c * c * u2 * a1 * a2 * b1 * b2 + t2 * t2 * a2 * b2 + rho1 * rho2 * a2 * b1;
    t5 = 1. / (d1 + d2);
    t7 = 2. * rho2 * t5;

    rpp = (d2 - d1) * t5;
    rps = 2. * u * a2 * t5 * (t1 * t2 + c * t3 * a1 * b1);
    tpp = a2 * t7 * (t3 * b1 - t2 * b2);
    tps = -a2 * t7 * u * (t1 + c * a1 * b2);
    rss = (d2 - d1 - 2. * rho1 * rho2 * (a2 * b1 - a1 * b2)) * t5;
    rsp = -2. * u * b2 * t5 * (t1 * t2 + c * t3 * a1 * b1);
    tss = b2 * t7 * t4;
    tsp = b2 * t7 * u * (t1 + c * a2 * b1);
    ru->c11 = rpp; ru->c12 = rsp; ru->c21 = rps; ru->c22 = rss;
    tu->c11 = tpp; tu->c12 = tsp; tu->c21 = tps; tu->c22 = tss;
}

/* greens.cpp:87-112 free-surface reflection (P-SV part) */
static void coeffs(double u, double vp, double vs, cmat2 *ru)
{
    double u2 = u * u;
    cplx a = csqrt(1. / (vp * vp) - u2 + 0.0 * I);
    cplx b = csqrt(1. / (vs * vs) - u2 + 0.0 * I);
    cplx t1 = 2. * vs * vs;
    cplx t2 = t1 * u2 - 1.;
    cplx d1 = t2 * t2;
    cplx d2 = t1 * t1 * u2 * a * b;
    cplx d = d1 + d2;
    cplx t3 = 2. * t1 * u * t2 / d;
    cplx rpp = (d2 - d1) / d;
    cplx rsp = -b * t3;
    cplx rps = a * t3;
    ru->c11 = rpp; ru->c12 = rsp; ru->c21 = rps; ru->c22 = rpp;
}

/* greens.cpp:307-322 */
static void displacement_matrix(double p, double vp, double vs, cmat2 *m)
{
    double vp2 = vp * vp, vs2 = vs * vs, p2 = p * p, x = 1. - 2. * vs2 * p2;
    cplx a1 = conj(csqrt(1. / vp2 - p2 + 0.0 * I));
    cplx b1 = conj(csqrt(1. / vs2 - p2 + 0.0 * I));
    cplx q = 1. / (x * x + 4. * vs2 * vs2 * p2 * a1 * b1);
    m->c11 = q * a1 * b1 * 2. * vs2 * p;
    m->c12 = q * b1 * (1. - 2. * vs2 * p2);
    m->c21 = q * a1 * (1. - 2. * vs2 * p2);
    m->c22 = -q * a1 * b1 * 2. * vs2 * p;
}

/* fork.cpp:10-60 */
static void ccfork(int n, cplx *x, int signi)
{
    cplx w, tmp;
    double sc = sqrt(1. / (double)n);
    int i, istep, j = 0, l, m;
    for (i = 0; i < n; i++) {
        if (i <= j) {
            tmp = x[j] * sc;
            x[j] = x[i] * sc;
            x[i] = tmp;
        }
        m = n >> 1;
        do {
            if (j < m) break;
            j -= m;
            m >>= 1;
        } while (m >= 1);
        j += m;
    }
    l = 1;
    do {
        istep = 2 * l;
        for (m = 0; m < l; m++) {
            w = cexp(0.0 + I * (M_PI * (double)(signi * m) / (double)l));
            for (i = m; i < n; i += istep) {
                tmp = w * x[i + l];
                x[i + l] = x[i] - tmp;
                x[i] += tmp;
            }
        }
        l = istep;
    } while (l < n);
}

/*
 * Same argument meaning as the reference's extern "C" synrf_cwrap
 * (wrap.cpp:57-80) minus the discarded fz/fr outputs.  waveno 0 = P, 1 = SV.
 * rf must hold nsamp doubles.  Returns 1 like the reference; 0 if nlay < 2
 * (the reference reads uninitialised memory for a half-space-only model;
 * SURVEY App. D.2-7 defines that case as invalid).
 */
int rf_oracle(int nsamp, double fsamp, double tshift, double p, double a,
              double nsv, double sigma, int waveno, int nlay,
              const double *z, const double *vp, const double *vs,
              const double *rh, const double *qp, const double *qs, double *rf)
{
    if (nlay < 2) return 0;
    int nfreq = nsamp / 2 + 1;
    double vptop = nsv * sqrt((1. - sigma) / (.5 - sigma));   /* wrap.cpp:13,73 */
    double vstop = nsv;
    double u = p * RF_DEGREES_PER_KM;                          /* wrap.cpp:76 */
    double fref = 1.;                                          /* synrf.cpp:25 */

    layer_t *lay = (layer_t *)malloc(sizeof(layer_t) * (nlay + 1));
    for (int i = 0; i < nlay - 1; i++) {                       /* synrf.cpp:28-30 */
        layer_t l = { z[i], z[i + 1] - z[i], vp[i], vs[i], rh[i], qp[i], qs[i] };
        lay[i + 1] = l;
    }
    {
        layer_t l = { z[nlay - 1], -1, vp[nlay - 1], vs[nlay - 1], rh[nlay - 1],
                      qp[nlay - 1], qs[nlay - 1] };
        lay[nlay] = l;
    }
    for (int i = 1; i <= nlay; i++) flatten_layer(&lay[i]);

    cmat2 *ru = malloc(sizeof(cmat2) * (nlay + 2)), *rd = malloc(sizeof(cmat2) * (nlay + 2)),
          *tu = malloc(sizeof(cmat2) * (nlay + 2)), *td = malloc(sizeof(cmat2) * (nlay + 2)),
          *nb = malloc(sizeof(cmat2) * (nlay + 2)), *nt = malloc(sizeof(cmat2) * (nlay + 2)),
          *g = malloc(sizeof(cmat2) * (nlay + 2)), *e = malloc(sizeof(cmat2) * (nlay + 2));
    cplx *cz = malloc(sizeof(cplx) * nfreq), *cr = malloc(sizeof(cplx) * nfreq),
         *crf = malloc(sizeof(cplx) * nfreq), *cx = malloc(sizeof(cplx) * nsamp);
    const cmat2 zero = { 0, 0, 0, 0 };

    /* greens.cpp:462-468, 114-132 */
    for (int i = 1; i <= nlay; i++) {
        if (i == 1) {
            coeffs(u, lay[1].vp, lay[1].vs, &ru[1]);
            rd[1] = td[1] = tu[1] = zero;
        } else {
            coeffm(u, lay[i - 1].vp, lay[i - 1].vs, lay[i - 1].rh,
                   lay[i].vp, lay[i].vs, lay[i].rh, &rd[i], &td[i], &ru[i], &tu[i]);
        }
    }
    cmat2 h;
    displacement_matrix(u, lay[1].vp, lay[1].vs, &h);          /* greens.cpp:495 */

    double p2 = u * u;
    double wref = 2. * M_PI * fref;
    double dw = 2.0 * M_PI * fsamp / nsamp;
    double t0 = 0.;                                            /* greens.cpp:510-526 */
    for (int i = 1; i <= nlay; i++) {
        double v = (waveno == 1) ? lay[i].vs : lay[i].vp;
        t0 += lay[i].h * sqrt(1. / (v * v) - p2);
    }

    const cmat2 ident = { 1., 0., 0., 1. };
    for (int j = 0; j < nfreq; j++) {                          /* greens.cpp:528-590 */
        double w = dw * j;
        double lgw = j ? log(w / wref) : 0;
        for (int i = 1; i <= nlay; i++) {
            double d = lay[i].h;
            cplx miwd = 0.0 + I * (-w * d);
            cplx vpc = lay[i].vp * (1. + lgw / (M_PI * lay[i].qp) + I / (2. * lay[i].qp));
            cplx vsc = lay[i].vs * (1. + lgw / (M_PI * lay[i].qs) + I / (2. * lay[i].qs));
            cplx plc = csqrt(1. / (vpc * vpc) - p2);
            cplx slc = csqrt(1. / (vsc * vsc) - p2);
            e[i].c11 = cexp(miwd * plc);
            e[i].c12 = 0; e[i].c21 = 0;
            e[i].c22 = cexp(miwd * slc);
        }
        /* top_down, greens.cpp:196-224 (normal case, options = 0) */
        cmat2 q = zero;
        for (int i = 1; i < nlay; i++) {
            if (i == 1) nt[i] = ru[1];
            else        nt[i] = m_add(ru[i], m_mul(m_mul(td[i], nb[i - 1]), q));
            nb[i] = m_exe(e[i], nt[i]);
            q = m_mul(m_inv(m_sub(ident, m_mul(rd[i + 1], nb[i]))), tu[i + 1]);
            if (i == 1) g[i] = m_mul(e[1], q);
            else        g[i] = m_mul(m_mul(g[i - 1], e[i]), q);
        }
        /* t = 2*h*g[nlay-1]  (greens.cpp:572; double*Cmat2 then Cmat2*Cmat2) */
        cmat2 h2 = { 2. * h.c11, 2. * h.c12, 2. * h.c21, 2. * h.c22 };
        cmat2 t = m_mul(h2, g[nlay - 1]);
        if (waveno == 1) { cr[j] = t.c12; cz[j] = t.c22; }     /* :579-581 */
        else             { cr[j] = t.c11; cz[j] = t.c21; }     /* :576-578 */
        cplx qq = cexp(0.0 + I * (w * t0));                    /* :583-585 */
        cr[j] *= qq;
        cz[j] *= qq;
    }

    /* compute_rf, greens.cpp:343-398 */
    {
        double qn = sqrt(M_PI) * fsamp / a;
        cplx *pz = cz, *pr = cr;
        if (vstop > 0.01 && fabs(u) > 0.0001) {                /* decomp :324-341 */
            double da = sqrt(1. / (vptop * vptop) - u * u),
                   db = sqrt(1. / (vstop * vstop) - u * u),
                   m11 = -(2 * vstop * vstop * u * u - 1.) / (vptop * da),
                   m12 = 2. * u * vstop * vstop / vptop,
                   m21 = -2. * u * vstop,
                   m22 = (1. - 2. * vstop * vstop * u * u) / (vstop * db);
            for (int i = 0; i < nfreq; i++) {
                cplx x = cz[i] * m11 + cr[i] * m12;
                cplx y = cz[i] * m21 + cr[i] * m22;
                cz[i] = x;
                cr[i] = y;
            }
        }
        if (waveno == 1) { cplx *tmp = pz; pz = pr; pr = tmp; }
        for (int j = 0; j < nfreq; j++) {
            double w = dw * j;
            double denom = creal(pz[j] * conj(pz[j]));
            crf[j] = pr[j] * conj(pz[j]) / denom;
            double wa = w / a;
            wa = (wa > 50.0) ? 50.0 : wa;
            cplx cq = qn * cexp(-0.25 * (wa * wa) + I * (-w * tshift));
            crf[j] = crf[j] * cq;
        }
    }

    /* iftr, greens.cpp:136-158 */
    {
        double q = 1. / sqrt((double)nsamp);
        for (int i = 0; i < nsamp / 2 + 1; i++) cx[i] = crf[i];
        for (int i = nsamp / 2 + 1; i < nsamp; i++) cx[i] = conj(cx[nsamp - i]);
        ccfork(nsamp, cx, 1);
        for (int i = 0; i < nsamp; i++) rf[i] = q * creal(cx[i]);
    }

    free(lay); free(ru); free(rd); free(tu); free(td); free(nb); free(nt); free(g); free(e);
    free(cz); free(cr); free(crf); free(cx);
    return 1;
}

/*
 * Batched wrapper over the packed product layout [B][lmax][4] = (h, vp, vs, rho)
 * doing exactly what RFminiModRF.compute_rf does around the native call
 * (rfmini_modrf.py:99-142): z from cumsum(h), Poisson ratio from the top-layer
 * vp/vs, nsv = vs[0] unless given (> 0), Q defaults 500 / 225 unless arrays given.
 * rf_out: [B][ndata] (first ndata samples).
 */
void rf_oracle_batch(const double *model, const int *nlay, int B, int lmax,
                     int nsamp, double fsamp, double tshift, double p, double a,
                     double nsv_in, int waveno, int ndata,
                     const double *qp_in, const double *qs_in, double *rf_out)
{
    double *rf = malloc(sizeof(double) * nsamp);
    double *z = malloc(sizeof(double) * lmax * 6);
    double *vp = z + lmax, *vs = vp + lmax, *rh = vs + lmax, *qp = rh + lmax, *qs = qp + lmax;
    for (int ib = 0; ib < B; ++ib) {
        int n = nlay[ib];
        double zc = 0.0;
        for (int i = 0; i < n; ++i) {
            const double *r = model + ((size_t)ib * lmax + i) * 4;
            z[i] = zc;           /* z = [0, cumsum(h)[:-1]] */
            zc += r[0];
            vp[i] = r[1]; vs[i] = r[2]; rh[i] = r[3];
            qp[i] = qp_in ? qp_in[(size_t)ib * lmax + i] : 500.;
            qs[i] = qs_in ? qs_in[(size_t)ib * lmax + i] : 225.;
        }
        double vpvs = vp[0] / vs[0];
        double poisson = (2 - vpvs * vpvs) / (2 - 2 * vpvs * vpvs);
        double nsv = nsv_in > 0 ? nsv_in : vs[0];
        int ok = rf_oracle(nsamp, fsamp, tshift, p, a, nsv, poisson, waveno, n,
                           z, vp, vs, rh, qp, qs, rf);
        for (int i = 0; i < ndata; ++i)
            rf_out[(size_t)ib * ndata + i] = ok ? rf[i] : NAN;
    }
    free(rf); free(z);
}
