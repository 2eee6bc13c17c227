/*
 * oracle/surf96_oracle.c  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C, scalar) of the SURF96 dispersion routine that
 * BayHunter calls through f2py: reference src/extensions/surfdisp96.f
 * (whole file, 1068 lines).  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load this file's
 * shared object.  The CUDA product path never calls it.
 *
 * Why a restatement: no Fortran compiler exists in the build image
 * (gfortran/flang/nvfortran/f2c all absent), so the reference Fortran
 * cannot be compiled into oracle/_ref.  Parity is pinned to the four
 * golden dispersion fixtures of the reference (tutorial/observed/
 * st3_{r,l}disp{ph,gr}.dat, 4 decimals => +-5e-5 km/s); beyond that
 * precision no reference ARTEFACT pins it (stated in DESIGN.md), but
 * oracle/first_principles.py does: the phase velocities returned here
 * (fundamental and first higher mode) are zeros of the exact layered-
 * medium eigenproblem within 1.2e-6 c, the reference's own search
 * tolerance (tests/test_oracle.py).
 *
 * Typing follows the Fortran exactly (the file has no IMPLICIT NONE in
 * the main routine): model arrays, start value, group-velocity formula
 * and a handful of literals are REAL*4; the root search is REAL*8.
 * Compile with -O2 -ffp-contract=off (x86-64 gfortran emits no FMA).
 *
 * Every function cites the Fortran lines it follows.
 */
#include <math.h>
#include <string.h>

#define S96_NL 100
#define S96_NP 60

typedef struct {
    int mmax, llw;
    float d[S96_NL], a[S96_NL], b[S96_NL], rho[S96_NL];
    float rtp[S96_NL], dtp[S96_NL], btp[S96_NL];
    float dhalf;          /* SAVE dhalf   (surfdisp96.f:515) */
    double del1st;        /* SAVE del1st  (surfdisp96.f:415) */
    long nsec_bracket;    /* secular evaluations issued by getsol */
    long nsec_refine;     /* secular evaluations issued by nevill/half */
} s96_ctx;

static double dsign1(double x) { return copysign(1.0, x); }

/* surfdisp96.f:367-388  gtsolh -- all REAL*4 */
static float s96_gtsolh(float a, float b)
{
    float c = 0.95f * b;
    for (int i = 0; i < 5; ++i) {
        float gamma = b / a;
        float kappa = c / b;
        float k2 = kappa * kappa;
        float gk = gamma * kappa;
        float gk2 = gk * gk;
        float fac1 = sqrtf(1.0f - gk2);
        float fac2 = sqrtf(1.0f - k2);
        float tk = 2.0f - k2;
        float fr = tk * tk - 4.0f * fac1 * fac2;
        float frp = -4.0f * tk * kappa
                  + 4.0f * fac2 * gamma * gamma * kappa / fac1
                  + 4.0f * fac1 * kappa / fac2;
        frp = frp / b;
        c = c - fr / frp;
    }
    return c;
}

/* surfdisp96.f:486-553  sphere (only when iflsph == 1) */
static void s96_sphere(s96_ctx *m, int ifunc, int iflag)
{
    double ar = 6370.0, dr = 0.0, r0 = ar, r1, z0, z1, tmp;
    int n = m->mmax;
    m->d[n - 1] = 1.0f;
    if (iflag == 0) {
        for (int i = 0; i < n; ++i) { m->dtp[i] = m->d[i]; m->rtp[i] = m->rho[i]; }
        for (int i = 0; i < n; ++i) {
            dr = dr + (double)m->d[i];
            r1 = ar - dr;
            z0 = ar * log(ar / r0);
            z1 = ar * log(ar / r1);
            m->d[i] = (float)(z1 - z0);
            tmp = (ar + ar) / (r0 + r1);
            m->a[i] = (float)((double)m->a[i] * tmp);
            m->b[i] = (float)((double)m->b[i] * tmp);
            m->btp[i] = (float)tmp;
            r0 = r1;
        }
        m->dhalf = m->d[n - 1];
    } else {
        m->d[n - 1] = m->dhalf;
        for (int i = 0; i < n; ++i) {
            if (ifunc == 1) {     /* btp**(-5): REAL*4 integer power; gcc's powi
                                     chain x^5 = (x^2*x)*(x^2), then the reciprocal */
                float x1 = m->btp[i], x2 = x1 * x1, x3 = x2 * x1, x5 = x3 * x2;
                m->rho[i] = m->rtp[i] * (1.0f / x5);
            }
            else if (ifunc == 2)  /* btp**(-2.275): REAL*4 pow */
                m->rho[i] = m->rtp[i] * powf(m->btp[i], -2.275f);
        }
    }
    m->d[n - 1] = 0.0f;
}

/* surfdisp96.f:710-769  dltar1 -- Love, Haskell 2-vector from the half-space up */
static double s96_dltar1(const s96_ctx *m, double wvno, double omega)
{
    int n = m->mmax;
    double beta1 = (double)m->b[n - 1];
    double rho1 = (double)m->rho[n - 1];
    double xkb = omega / beta1;
    double wvnop = wvno + xkb;
    double wvnom = fabs(wvno - xkb);
    double rb = sqrt(wvnop * wvnom);
    double e1 = rho1 * rb;
    double e2 = 1.0 / (beta1 * beta1);
    for (int l = n - 2; l >= m->llw - 1; --l) {
        double cosq, y, z, sinq;
        beta1 = (double)m->b[l];
        rho1 = (double)m->rho[l];
        double xmu = rho1 * beta1 * beta1;
        xkb = omega / beta1;
        wvnop = wvno + xkb;
        wvnom = fabs(wvno - xkb);
        rb = sqrt(wvnop * wvnom);
        double q = (double)m->d[l] * rb;
        if (wvno < xkb) {
            sinq = sin(q);
            y = sinq / rb;
            z = -rb * sinq;
            cosq = cos(q);
        } else if (wvno == xkb) {
            cosq = 1.0;
            y = (double)m->d[l];
            z = 0.0;
        } else {
            double fac = 0.0;
            if (q < 16.0) fac = exp(-2.0 * q);
            cosq = (1.0 + fac) * 0.5;
            sinq = (1.0 - fac) * 0.5;
            y = sinq / rb;
            z = rb * sinq;
        }
        double e10 = e1 * cosq + e2 * xmu * z;
        double e20 = e1 * y / xmu + e2 * cosq;
        double xnor = fabs(e10);
        double ynor = fabs(e20);
        if (ynor > xnor) xnor = ynor;
        if (xnor < 1.0e-40) xnor = 1.0;
        e1 = e10 / xnor;
        e2 = e20 / xnor;
    }
    return e1;
}

typedef struct { double a0, cpcq, cpy, cpz, cqw, cqx, xy, xz, wy, wz; } s96_prod;

/* surfdisp96.f:874-991  var */
static void s96_var(double p, double q, double ra, double rb, double wvno,
                    double xka, double xkb, double dpth,
                    double *w_out, double *cosp_out, double *exa_out, s96_prod *o)
{
    double pex = 0.0, sex = 0.0, fac;
    double sinp, w = 0.0, x = 0.0, cosp = 0.0, sinq, y = 0.0, z = 0.0, cosq = 0.0;
    if (wvno < xka) {
        sinp = sin(p);
        w = sinp / ra;
        x = -ra * sinp;
        cosp = cos(p);
    } else if (wvno == xka) {
        cosp = 1.0;
        w = dpth;
        x = 0.0;
    } else if (wvno > xka) {
        pex = p;
        fac = 0.0;
        if (p < 16.0) fac = exp(-2.0 * p);
        cosp = (1.0 + fac) * 0.5;
        sinp = (1.0 - fac) * 0.5;
        w = sinp / ra;
        x = ra * sinp;
    }
    if (wvno < xkb) {
        sinq = sin(q);
        y = sinq / rb;
        z = -rb * sinq;
        cosq = cos(q);
    } else if (wvno == xkb) {
        cosq = 1.0;
        y = dpth;
        z = 0.0;
    } else if (wvno > xkb) {
        sex = q;
        fac = 0.0;
        if (q < 16.0) fac = exp(-2.0 * q);
        cosq = (1.0 + fac) * 0.5;
        sinq = (1.0 - fac) * 0.5;
        y = sinq / rb;
        z = rb * sinq;
    }
    double exa = pex + sex;
    o->a0 = 0.0;
    if (exa < 60.0) o->a0 = exp(-exa);
    o->cpcq = cosp * cosq;
    o->cpy = cosp * y;
    o->cpz = cosp * z;
    o->cqw = cosq * w;
    o->cqx = cosq * x;
    o->xy = x * y;
    o->xz = x * z;
    o->wy = w * y;
    o->wz = w * z;
    /* :984-989 rescale cosq,y,z by exp(sex-pex): results are local and never
       read by dnka -- dead in the Fortran as well, not restated. */
    *w_out = w;
    *cosp_out = cosp;
    *exa_out = exa;
}

/* surfdisp96.f:1024-1068  dnka -- Dunkin 5x5 compound matrix, ca[row][col] 0-based */
static void s96_dnka(double ca[5][5], double wvno2, double gam, double gammk,
                     double rho, const s96_prod *v)
{
    const double one = 1.0, two = 2.0;
    double gamm1 = gam - one;
    double twgm1 = gam + gamm1;
    double gmgmk = gam * gammk;
    double gmgm1 = gam * gamm1;
    double gm1sq = gamm1 * gamm1;
    double rho2 = rho * rho;
    double a0pq = v->a0 - v->cpcq;
    ca[0][0] = v->cpcq - two * gmgm1 * a0pq - gmgmk * v->xz - wvno2 * gm1sq * v->wy;
    ca[0][1] = (wvno2 * v->cpy - v->cqx) / rho;
    ca[0][2] = -(twgm1 * a0pq + gammk * v->xz + wvno2 * gamm1 * v->wy) / rho;
    ca[0][3] = (v->cpz - wvno2 * v->cqw) / rho;
    ca[0][4] = -(two * wvno2 * a0pq + v->xz + wvno2 * wvno2 * v->wy) / rho2;
    ca[1][0] = (gmgmk * v->cpz - gm1sq * v->cqw) * rho;
    ca[1][1] = v->cpcq;
    ca[1][2] = gammk * v->cpz - gamm1 * v->cqw;
    ca[1][3] = -v->wz;
    ca[1][4] = ca[0][3];
    ca[3][0] = (gm1sq * v->cpy - gmgmk * v->cqx) * rho;
    ca[3][1] = -v->xy;
    ca[3][2] = gamm1 * v->cpy - gammk * v->cqx;
    ca[3][3] = ca[1][1];
    ca[3][4] = ca[0][1];
    ca[4][0] = -(two * gmgmk * gm1sq * a0pq + gmgmk * gmgmk * v->xz + gm1sq * gm1sq * v->wy) * rho2;
    ca[4][1] = ca[3][0];
    ca[4][2] = -(gammk * gamm1 * twgm1 * a0pq + gam * gammk * gammk * v->xz + gamm1 * gm1sq * v->wy) * rho;
    ca[4][3] = ca[1][0];
    ca[4][4] = ca[0][0];
    double t = -two * wvno2;
    ca[2][0] = t * ca[4][2];
    ca[2][1] = t * ca[3][2];
    ca[2][2] = v->a0 + two * (v->cpcq - ca[0][0]);
    ca[2][3] = t * ca[1][2];
    ca[2][4] = t * ca[0][2];
}

/* surfdisp96.f:995-1020  normc */
static void s96_normc(double ee[5], double *ex)
{
    double t1 = 0.0;
    for (int i = 0; i < 5; ++i)
        if (fabs(ee[i]) > t1) t1 = fabs(ee[i]);
    if (t1 < 1.0e-40) t1 = 1.0;
    for (int i = 0; i < 5; ++i) ee[i] = ee[i] / t1;
    *ex = log(t1);
}

/* surfdisp96.f:773-871  dltar4 -- Rayleigh, Dunkin compound vector */
static double s96_dltar4(const s96_ctx *m, double wvno, double omga)
{
    double e[5], ee[5], ca[5][5];
    int n = m->mmax;
    double omega = omga;
    if (omega < 1.0e-4) omega = 1.0e-4;
    double wvno2 = wvno * wvno;
    double xka = omega / (double)m->a[n - 1];
    double xkb = omega / (double)m->b[n - 1];
    double wvnop = wvno + xka;
    double wvnom = fabs(wvno - xka);
    double ra = sqrt(wvnop * wvnom);
    wvnop = wvno + xkb;
    wvnom = fabs(wvno - xkb);
    double rb = sqrt(wvnop * wvnom);
    double t = (double)m->b[n - 1] / omega;
    double gammk = 2.0 * t * t;
    double gam = gammk * wvno2;
    double gamm1 = gam - 1.0;
    double rho1 = (double)m->rho[n - 1];
    e[0] = rho1 * rho1 * (gamm1 * gamm1 - gam * gammk * ra * rb);
    e[1] = -rho1 * ra;
    e[2] = rho1 * (gamm1 - gammk * ra * rb);
    e[3] = rho1 * rb;
    e[4] = wvno2 - ra * rb;
    s96_prod v;
    double w, cosp, exa;
    for (int l = n - 2; l >= m->llw - 1; --l) {
        xka = omega / (double)m->a[l];
        xkb = omega / (double)m->b[l];
        t = (double)m->b[l] / omega;
        gammk = 2.0 * t * t;
        gam = gammk * wvno2;
        wvnop = wvno + xka;
        wvnom = fabs(wvno - xka);
        ra = sqrt(wvnop * wvnom);
        wvnop = wvno + xkb;
        wvnom = fabs(wvno - xkb);
        rb = sqrt(wvnop * wvnom);
        double dpth = (double)m->d[l];
        rho1 = (double)m->rho[l];
        double p = ra * dpth;
        double q = rb * dpth;
        s96_var(p, q, ra, rb, wvno, xka, xkb, dpth, &w, &cosp, &exa, &v);
        s96_dnka(ca, wvno2, gam, gammk, rho1, &v);
        for (int i = 0; i < 5; ++i) {
            double cr = 0.0;
            for (int j = 0; j < 5; ++j) cr = cr + e[j] * ca[j][i];
            ee[i] = cr;
        }
        s96_normc(ee, &exa);
        for (int i = 0; i < 5; ++i) e[i] = ee[i];
    }
    if (m->llw != 1) {
        /* :850-867 water layer on top */
        xka = omega / (double)m->a[0];
        wvnop = wvno + xka;
        wvnom = fabs(wvno - xka);
        ra = sqrt(wvnop * wvnom);
        double dpth = (double)m->d[0];
        rho1 = (double)m->rho[0];
        double p = ra * dpth;
        double znul = 1.0e-05;
        s96_var(p, znul, ra, znul, wvno, xka, znul, dpth, &w, &cosp, &exa, &v);
        double w0 = -rho1 * w;
        return cosp * e[0] + w0 * e[1];
    }
    return e[0];
}

/* surfdisp96.f:690-706  dltar */
static double s96_dltar(const s96_ctx *m, double wvno, double omega, int kk)
{
    return kk == 1 ? s96_dltar1(m, wvno, omega) : s96_dltar4(m, wvno, omega);
}

/* surfdisp96.f:676-686  half */
static void s96_half(s96_ctx *m, double c1, double c2, double *c3, double *del3,
                     double omega, int ifunc)
{
    *c3 = 0.5 * (c1 + c2);
    double wvno = omega / *c3;
    *del3 = s96_dltar(m, wvno, omega, ifunc);
    m->nsec_refine++;
}

/* surfdisp96.f:557-674  nevill */
static double s96_nevill(s96_ctx *mc, double t, double c1, double c2,
                         double del1, double del2, int ifunc, double twopi)
{
    double x[20], y[20];
    double c3, del3;
    double omega = twopi / t;
    int m = 0;
    s96_half(mc, c1, c2, &c3, &del3, omega, ifunc);
    int nev = 1;
    int nctrl = 1;
    for (;;) {
        nctrl = nctrl + 1;
        if (nctrl >= 100) break;
        if (c3 < fmin(c1, c2) || c3 > fmax(c1, c2)) {
            nev = 0;
            s96_half(mc, c1, c2, &c3, &del3, omega, ifunc);
        }
        double s13 = del1 - del3;
        double s32 = del3 - del2;
        if (dsign1(del3) * dsign1(del1) < 0.0) {
            c2 = c3;
            del2 = del3;
        } else {
            c1 = c3;
            del1 = del3;
        }
        if (fabs(c1 - c2) <= 1.0e-6 * c1) break;
        if (dsign1(s13) != dsign1(s32)) nev = 0;
        double ss1 = fabs(del1);
        double s1 = (double)0.01f * ss1;   /* 0.01 is a REAL*4 literal (:625) */
        double ss2 = fabs(del2);
        double s2 = (double)0.01f * ss2;   /* (:627) */
        if (s1 > ss2 || s2 > ss1 || nev == 0) {
            s96_half(mc, c1, c2, &c3, &del3, omega, ifunc);
            nev = 1;
            m = 1;
        } else {
            if (nev == 2) {
                x[m] = c3;       /* x(m+1), 1-based */
                y[m] = del3;
            } else {
                x[0] = c1; y[0] = del1;
                x[1] = c2; y[1] = del2;
                m = 1;
            }
            int bad = 0;
            for (int kk = 1; kk <= m; ++kk) {
                int j = m - kk;  /* 0-based index of x(j), j = m-kk+1 */
                double denom = y[m] - y[j];
                if (fabs(denom) < 1.0e-10 * fabs(y[m])) { bad = 1; break; }
                x[j] = (-y[j] * x[j + 1] + y[m] * x[j]) / denom;
            }
            if (bad) {
                s96_half(mc, c1, c2, &c3, &del3, omega, ifunc);
                nev = 1;
                m = 1;
            } else {
                c3 = x[0];
                double wvno = omega / c3;
                del3 = s96_dltar(mc, wvno, omega, ifunc);
                mc->nsec_refine++;
                nev = 2;
                m = m + 1;
                if (m > 10) m = 10;
            }
        }
    }
    return c3;
}

/* surfdisp96.f:390-482  getsol; returns iret (+1 / -1), root in *c1 */
static int s96_getsol(s96_ctx *m, double t1, double *c1io, double clow, double dc,
                      double cm, float betmx, int ifunc, int ifirst)
{
    double c1 = *c1io, c2, del1, del2;
    double twopi = 2.0 * 3.141592653589793;
    double omega = twopi / t1;
    double wvno = omega / c1;
    int idir;
    del1 = s96_dltar(m, wvno, omega, ifunc);
    m->nsec_bracket++;
    if (ifirst == 1) m->del1st = del1;
    double plmn = dsign1(m->del1st) * dsign1(del1);
    if (ifirst == 1) idir = +1;
    else if (plmn >= 0.0) idir = +1;
    else idir = -1;
    for (;;) {
        if (idir > 0) c2 = c1 + dc;
        else          c2 = c1 - dc;
        if (c2 <= clow) {
            idir = +1;
            c1 = clow;
            continue;
        }
        omega = twopi / t1;
        wvno = omega / c2;
        del2 = s96_dltar(m, wvno, omega, ifunc);
        m->nsec_bracket++;
        if (dsign1(del1) != dsign1(del2)) {
            double cn = s96_nevill(m, t1, c1, c2, del1, del2, ifunc, twopi);
            c1 = cn;
            *c1io = c1;
            if (c1 > (double)betmx) return -1;
            return 1;
        }
        c1 = c2;
        del1 = del2;
        if (c1 < cm) break;
        if (c1 >= ((double)betmx + dc)) break;
    }
    *c1io = c1;
    return -1;
}

/*
 * surfdisp96.f:55-360.  Same argument meaning as the Fortran entry
 * (REAL*4 model arrays of length >= nlayer, REAL*8 t/cg of length >= kmax).
 * Extra outputs: nsec[0] bracket evaluations, nsec[1] refine evaluations
 * (may be NULL).  cg must hold kmax entries.
 */
void surf96_oracle(const float *thkm, const float *vpm, const float *vsm,
                   const float *rhom, int nlayer, int iflsph, int iwave,
                   int mode, int igr, int kmax, const double *t, double *cg,
                   int *err, long *nsec)
{
    s96_ctx ctx;
    s96_ctx *m = &ctx;
    double c[S96_NP], cb[S96_NP];
    memset(m, 0, sizeof(*m));
    m->mmax = nlayer;
    *err = 0;
    for (int i = 0; i < nlayer; ++i) {
        m->b[i] = vsm[i];
        m->a[i] = vpm[i];
        m->d[i] = thkm[i];
        m->rho[i] = rhom[i];
    }
    int idispl = 0, idispr = 0;
    if (iwave == 1) { idispl = kmax; idispr = 0; }
    else if (iwave == 2) { idispl = 0; idispr = kmax; }
    int iverb[2] = {0, 0};
    const float sone0 = 1.500f, ddc0 = 0.005f, h0 = 0.005f;
    m->llw = 1;
    if (m->b[0] <= 0.0f) m->llw = 2;
    const double twopi = 2.0 * 3.141592653589793;
    const double one = 1.0e-2;
    (void)twopi;
    if (iflsph == 1) s96_sphere(m, 0, 0);
    int jmn = 0, jsol = 1;
    float betmx = -1.e20f, betmn = 1.e20f;
    for (int i = 0; i < nlayer; ++i) {
        if (m->b[i] > 0.01f && m->b[i] < betmn) {
            betmn = m->b[i]; jmn = i; jsol = 1;
        } else if (m->b[i] <= 0.01f && m->a[i] < betmn) {
            betmn = m->a[i]; jmn = i; jsol = 0;
        }
        if (m->b[i] > betmx) betmx = m->b[i];
    }
    for (int ifunc = 1; ifunc <= 2; ++ifunc) {
        if (ifunc == 1 && idispl <= 0) continue;
        if (ifunc == 2 && idispr <= 0) continue;
        if (iflsph == 1) s96_sphere(m, ifunc, 1);
        float ddc = ddc0, sone = sone0, h = h0;
        if (sone < 0.01f) sone = 2.0f;
        double onea = (double)sone;
        float cc1;
        if (jsol == 0) cc1 = betmn;
        else cc1 = s96_gtsolh(m->a[jmn], m->b[jmn]);
        cc1 = .95f * cc1;
        cc1 = .90f * cc1;
        double cc = (double)cc1;
        double dc = fabs((double)ddc);
        double c1 = cc;
        double cm = cc;
        double clow = cc;
        for (int i = 0; i < kmax; ++i) { cb[i] = 0.0; c[i] = 0.0; }
        int ift = 999;
        for (int iq = 1; iq <= mode; ++iq) {
            int is = 1, ie = kmax, k;
            int failed = 0;
            for (k = is; k <= ie; ++k) {
                if (k >= ift) { failed = 1; break; }
                double t1 = t[k - 1];
                float t1a, t1b = 0.0f;
                if (igr > 0) {
                    t1a = (float)(t1 / (double)(1.0f + h));
                    t1b = (float)(t1 / (double)(1.0f - h));
                    t1 = (double)t1a;
                } else {
                    t1a = (float)t1;
                }
                int ifirst;
                if (k == is && iq == 1) {
                    c1 = cc; clow = cc; ifirst = 1;
                } else if (k == is && iq > 1) {
                    c1 = c[is - 1] + one * dc; clow = c1; ifirst = 1;
                } else if (k > is && iq > 1) {
                    ifirst = 0;
                    clow = c[k - 1] + one * dc;
                    c1 = c[k - 2];
                    if (c1 < clow) c1 = clow;
                } else {
                    ifirst = 0;
                    c1 = c[k - 2] - onea * dc;
                    clow = cm;
                }
                int iret = s96_getsol(m, t1, &c1, clow, dc, cm, betmx, ifunc, ifirst);
                if (iret == -1) { failed = 1; break; }
                c[k - 1] = c1;
                if (igr > 0) {
                    t1 = (double)t1b;
                    ifirst = 0;
                    clow = cb[k - 1] + one * dc;
                    c1 = c1 - onea * dc;
                    iret = s96_getsol(m, t1, &c1, clow, dc, cm, betmx, ifunc, ifirst);
                    if (iret == -1) c1 = c[k - 1];
                    cb[k - 1] = c1;
                } else {
                    c1 = 0.0;
                }
                float cc0 = (float)c[k - 1];
                float cc1s = (float)c1;
                if (igr == 0) {
                    cg[k - 1] = (double)cc0;
                } else {
                    /* :306 -- evaluated entirely in REAL*4 */
                    float num = 1.0f / t1a - 1.0f / t1b;
                    float den = 1.0f / (t1a * cc0) - 1.0f / (t1b * cc1s);
                    float gvel = num / den;
                    cg[k - 1] = (double)gvel;
                }
            }
            if (!failed) continue;
            /* label 1700 / 1750 */
            if (iq <= 1) {
                if (iverb[ifunc - 1] == 0) { iverb[ifunc - 1] = 1; *err = 1; }
            }
            ift = k;
            for (int i = k; i <= ie; ++i) cg[i - 1] = 0.0;
        }
    }
    if (nsec) { nsec[0] = m->nsec_bracket; nsec[1] = m->nsec_refine; }
}

/*
 * Batched convenience wrapper used by tests and by bench.py's CPU baseline:
 * model rows are the packed fp64 layout of the product API
 * [B][lmax][4] = (h, vp, vs, rho); rounding to REAL*4 happens here exactly as
 * f2py does at the reference FFI (surf96_modsw.py:116, surfdisp96.f:82).
 */
void surf96_oracle_batch(const double *model, const int *nlay, int B, int lmax,
                         int iflsph, int iwave, int mode, int igr, int kmax,
                         const double *t, double *cg /*[B][kmax]*/, int *err /*[B]*/,
                         long *nsec /*[B][2] or NULL*/)
{
    for (int ib = 0; ib < B; ++ib) {
        float h[S96_NL], vp[S96_NL], vs[S96_NL], rho[S96_NL];
        int n = nlay[ib];
        for (int i = 0; i < n; ++i) {
            const double *r = model + ((size_t)ib * lmax + i) * 4;
            h[i] = (float)r[0]; vp[i] = (float)r[1]; vs[i] = (float)r[2]; rho[i] = (float)r[3];
        }
        surf96_oracle(h, vp, vs, rho, n, iflsph, iwave, mode, igr, kmax, t,
                      cg + (size_t)ib * kmax, err + ib, nsec ? nsec + 2 * ib : 0);
    }
}
