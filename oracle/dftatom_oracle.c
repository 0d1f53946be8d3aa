/*
 * TEST INFRASTRUCTURE — NOT PART OF THE PRODUCT.  See dftatom_oracle.h for the rules of use and the
 * parity-pinning statement.  Plain C11 restatement of the reference's radial Kohn-Sham SCF; every
 * function cites the reference file:line (relative to /root/reference/DFTAtom/) it follows.
 * Arithmetic order mirrors the reference's expressions so that, compiled with the same flags
 * (gcc -O2, no FMA contraction on baseline x86-64), results agree with it to the last few ulps.
 */
#define _GNU_SOURCE
#include "dftatom_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

static const double FOUR_PI = 4. * M_PI;

/* ------------------------------------------------------------------------------------------------
 * L0: occupations
 * ---------------------------------------------------------------------------------------------- */

/* AufbauPrinciple.h:101-117 (lanthanide / actinide / Lr adjustments).  The transition-metal
 * adjustments (:78-99) are never called by the reference and are deliberately absent here. */
static int f_block_adjust(int occ, int Z, int n0, int l)
{
    if (l == 3) {
        if ((Z == 57 || Z == 58 || Z == 64) && n0 == 3) return occ - 1;          /* :105, :129-132 */
        if (n0 == 4) {
            if (Z == 89 || Z == 90) return 0;                                     /* :109-110 */
            if (Z == 91 || Z == 92 || Z == 93 || Z == 96) return occ - 1;         /* :111-112, :134-137 */
        }
        return occ;
    }
    if (Z == 103 && n0 == 5 && l == 2) return 0;                                  /* :115-116, :139-142 */
    return occ;
}

static int level_order(const void* pa, const void* pb)
{   /* Subshell::operator< AufbauPrinciple.h:10-13 */
    const orc_level* a = (const orc_level*)pa;
    const orc_level* b = (const orc_level*)pb;
    if (a->n0 != b->n0) return a->n0 < b->n0 ? -1 : 1;
    return (a->l > b->l) - (a->l < b->l);
}

int orc_aufbau(int Z, orc_level* out)
{   /* AufbauPrinciple.h:36-75: Madelung order by n0+l, then n0, shells with l <= n0 */
    int count = 0, placed = 0;
    for (int s = 0; s < 10 && placed != Z; ++s) {
        for (int n0 = 0; n0 <= s; ++n0) {
            const int l = s - n0;
            if (l > n0) continue;
            int occ = 2 * (2 * l + 1);
            occ = f_block_adjust(occ, Z, n0, l);         /* first call :53 */
            if (Z - placed < occ) occ = Z - placed;      /* :55-56 */
            occ = f_block_adjust(occ, Z, n0, l);         /* second call :59 */
            if (occ > 0) {
                placed += occ;
                out[count].n0 = n0; out[count].l = l; out[count].occ = occ; out[count].E = 0;
                ++count;
            }
            if (placed == Z) break;
        }
    }
    qsort(out, (size_t)count, sizeof(orc_level), level_order);   /* DFTAtom.cpp:367 (keys are unique) */
    return count;
}

void orc_split_spin(int Z, const orc_level* all, int n_all, orc_level* a, int* na, orc_level* b, int* nb,
                    int* n_alpha_el, int* n_beta_el)
{   /* DFTAtom.cpp:611-638 */
    int ea = 0, kb = 0;
    for (int i = 0; i < n_all; ++i) {
        const int cap = 2 * all[i].l + 1;
        a[i] = all[i];
        orc_level bl = all[i];
        if (all[i].occ >= cap) { a[i].occ = cap; bl.occ = all[i].occ - cap; ea += cap; }
        else { bl.occ = 0; ea += all[i].occ; }
        if (bl.occ != 0) b[kb++] = bl;
    }
    *na = n_all; *nb = kb; *n_alpha_el = ea; *n_beta_el = Z - ea;
}

/* ------------------------------------------------------------------------------------------------
 * grid
 * ---------------------------------------------------------------------------------------------- */

int orc_n_nodes(int levels)
{   /* PoissonSolver.h:127-135 with Ncoarse = 3 */
    int n = 3;
    for (int i = 0; i < levels - 1; ++i) n = 2 * n - 1;
    return n;
}

double orc_rp(int n_nodes, double delta, double max_r)
{   /* DFTAtom.cpp:356, Numerov.h:79 */
    return max_r / (exp(((double)n_nodes - 1.) * delta) - 1.);
}

/* ------------------------------------------------------------------------------------------------
 * L1: Numerov on the logarithmic grid (Numerov.h:73-196 function object, :199-518 sweeps)
 * ---------------------------------------------------------------------------------------------- */

typedef struct {
    const double* V;
    int n_nodes;
    double delta, rp, two_delta, rp2d2, d2q;
} nfun;

static nfun nfun_make(const double* V, int n_nodes, double delta, double max_r)
{   /* Numerov.h:76-87 */
    nfun g;
    g.V = V; g.n_nodes = n_nodes; g.delta = delta;
    g.rp = max_r / (exp(((double)n_nodes - 1.) * delta) - 1.);
    g.two_delta = 2. * delta;
    const double d2 = delta * delta;
    g.rp2d2 = g.rp * g.rp * d2;
    g.d2q = d2 * 0.25;
    return g;
}

static double nf_r(const nfun* g, long i) { return g->rp * (exp((double)i * g->delta) - 1.); }   /* :181-184 */

static double nf_veff(const nfun* g, int l, long i)
{   /* :89-94 */
    const double r = nf_r(g, i);
    return g->V[i] + l * (l + 1.) / (r * r) * 0.5;
}

static double nf_f(const nfun* g, int l, double E, long i)
{   /* :96-101 */
    return 2. * (nf_veff(g, l, i) - E) * g->rp2d2 * exp((double)i * g->two_delta) + g->d2q;
}

static double nf_far(const nfun* g, double idx, double E)
{   /* :103-108 */
    const double r = nf_r(g, (long)(int)idx);
    return exp(-r * sqrt(2. * fabs(E)) - idx * g->delta * 0.5);
}

static double nf_near(const nfun* g, double idx, int l)
{   /* :110-116 */
    const double r = nf_r(g, (long)(int)idx);
    return pow(r, (double)l + 1) * exp(-idx * g->delta * 0.5);
}

static long nf_start(const nfun* g, double E, long n_steps)
{   /* :119-136 — bisection on the index for far(idx) < 1e-200; callers then take min(n_steps, .) */
    long hi = n_steps, lo = 1;
    while (hi - lo > 1) {
        const long mid = (hi + lo) / 2;
        if (nf_far(g, (double)mid, E) < 1E-200) hi = mid; else lo = mid;
    }
    return hi;
}

int orc_numerov_start_index(double E, int n_nodes, double delta, double max_r)
{
    nfun g = nfun_make(NULL, n_nodes, delta, max_r);
    return (int)nf_start(&g, E, n_nodes - 1);
}

#define GETU(w, f) ((w) / (1. - (1. / 12.) * (f)))    /* Numerov.h:510-513 with h2p12 = 1/12 */

int orc_numerov_count_nodes(const double* V, int n_nodes, double delta, double max_r, int l, double E, int limit)
{   /* Numerov.h:272-349, non-uniform branch */
    const nfun g = nfun_make(V, n_nodes, delta, max_r);
    const long start = nf_start(&g, E, n_nodes - 1);
    const double twelfth = 1. / 12.;

    double y = nf_far(&g, (double)start, E);
    double yprev = y;
    double f = nf_f(&g, l, E, start);
    double wprev = (1 - twelfth * f) * y;

    y = nf_far(&g, (double)start - 1., E);
    f = nf_f(&g, l, E, start - 1);
    double w = (1 - twelfth * f) * y;

    int positive = y > 0;
    int count = 0;
    int seen_allowed = 0;
    for (long i = start - 2; i > 0; --i) {
        const double wnext = 2. * w - wprev + y * f;
        wprev = w; w = wnext;
        f = nf_f(&g, l, E, i);
        yprev = y;
        y = GETU(w, f);
        if (fabs(y) == INFINITY) return count;                     /* :323-324 */
        if ((y > 0) != positive) {                                 /* :326-333 */
            if (++count > limit) return count;
            positive = !positive;
        }
        const double veff = nf_veff(&g, l, i);                     /* :336-340 */
        if (veff <= E) seen_allowed = 1;
        else if (seen_allowed) return count;
    }
    if (count <= limit) {                                          /* :343-348 */
        y = y * (2 + f) - yprev;
        if ((y > 0) != positive) ++count;
    }
    return count;
}

int orc_numerov_count_from_nucleus(const double* V, int n_nodes, double delta, double max_r, int l, double E, int limit)
{   /* SolveSchrodingerCountNodesFromNucleus, Numerov.h:204-270, non-uniform branch (no caller in the reference):
     * outward from y_0 = 0, y_1 = near value, counting sign changes up to the cut-off index; returns on overflow, on count > limit,
     * and at the OUTER classical turning point (first forbidden node after the allowed region was seen). */
    const nfun g = nfun_make(V, n_nodes, delta, max_r);
    const long steps = nf_start(&g, E, n_nodes - 1);
    const double twelfth = 1. / 12.;
    double y = nf_near(&g, 1., l);
    double wprev = 0;
    double f = nf_f(&g, l, E, 1);
    double w = (1 - twelfth * f) * y;
    int positive = y > 0, count = 0;
    int seen_allowed = nf_veff(&g, l, 1) <= E;
    for (long i = 2; i <= steps; ++i) {
        const double wnext = 2. * w - wprev + y * f;
        wprev = w; w = wnext;
        f = nf_f(&g, l, E, i);
        y = GETU(w, f);
        if (fabs(y) == INFINITY) return count;
        if ((y > 0) != positive) {
            if (++count > limit) return count;
            positive = !positive;
        }
        const double veff = nf_veff(&g, l, i);
        if (veff <= E) seen_allowed = 1;
        else if (seen_allowed) return count;
    }
    return count;
}

int orc_numerov_count_all(const double* V, int n_nodes, double delta, double max_r, int l, double E)
{   /* NOT a reference function: the sweep of Numerov.h:272-349 with its three early exits disabled, i.e. the number
     * of sign changes in y_start-1 .. y_1, y_0 (the Sturm count the CUDA search uses as its predicate). */
    const nfun g = nfun_make(V, n_nodes, delta, max_r);
    const long start = nf_start(&g, E, n_nodes - 1);
    const double twelfth = 1. / 12.;
    double y = nf_far(&g, (double)start, E);
    double yprev = y;
    double f = nf_f(&g, l, E, start);
    double wprev = (1 - twelfth * f) * y;
    y = nf_far(&g, (double)start - 1., E);
    f = nf_f(&g, l, E, start - 1);
    double w = (1 - twelfth * f) * y;
    int positive = y > 0, count = 0;
    for (long i = start - 2; i > 0; --i) {
        const double wnext = 2. * w - wprev + y * f;
        wprev = w; w = wnext;
        f = nf_f(&g, l, E, i);
        yprev = y;
        y = GETU(w, f);
        if ((y > 0) != positive) { ++count; positive = !positive; }
    }
    y = y * (2 + f) - yprev;
    if ((y > 0) != positive) ++count;
    return count;
}

double orc_numerov_y0(const double* V, int n_nodes, double delta, double max_r, int l, double E)
{   /* Numerov.h:351-401 */
    const nfun g = nfun_make(V, n_nodes, delta, max_r);
    const long start = nf_start(&g, E, n_nodes - 1);
    const double twelfth = 1. / 12.;

    double y = nf_far(&g, (double)start, E);
    double yprev = y;
    double f = nf_f(&g, l, E, start);
    double wprev = (1 - twelfth * f) * y;
    y = nf_far(&g, (double)start - 1., E);
    f = nf_f(&g, l, E, start - 1);
    double w = (1 - twelfth * f) * y;
    for (long i = start - 2; i > 0; --i) {
        const double wnext = 2. * w - wprev + y * f;
        wprev = w; w = wnext;
        f = nf_f(&g, l, E, i);
        yprev = y;
        y = GETU(w, f);
    }
    return y * (2 + f) - yprev;                                    /* :398 */
}

long orc_numerov_match(const double* V, int n_nodes, double delta, double max_r, int l, double E, double* psi)
{   /* Numerov.h:403-504; on the log grid h stays exactly 1 after the :430 re-computation */
    const nfun g = nfun_make(V, n_nodes, delta, max_r);
    const long n_steps = n_nodes - 1;
    const long start = nf_start(&g, E, n_steps);
    const double twelfth = 1. / 12.;

    for (long i = start + 1; i <= n_steps; ++i) psi[i] = 0;        /* :427-428 */

    double y = nf_far(&g, (double)start, E);
    psi[start] = y;
    double f = nf_f(&g, l, E, start);
    double wprev = (1 - twelfth * f) * y;
    y = nf_far(&g, (double)start - 1., E);
    psi[start - 1] = y;
    f = nf_f(&g, l, E, start - 1);
    double w = (1 - twelfth * f) * y;

    long match = 2;                                                /* :449 */
    for (long i = start - 2; i > 0; --i) {
        const double wnext = 2. * w - wprev + y * f;
        wprev = w; w = wnext;
        f = nf_f(&g, l, E, i);
        psi[i] = y = GETU(w, f);
        if (y < psi[i + 1] || fabs(y) > 1E15) { match = i; break; }   /* :463-467 */
    }

    /* outward from the nucleus, :470-490 */
    psi[0] = 0; y = 0; wprev = 0;
    psi[1] = y = nf_near(&g, 1., l);
    f = nf_f(&g, l, E, 1);
    w = (1 - twelfth * f) * y;
    for (long i = 2; i < match; ++i) {
        const double wnext = 2. * w - wprev + y * f;
        wprev = w; w = wnext;
        f = nf_f(&g, l, E, i);
        psi[i] = y = GETU(w, f);
    }
    w = 2. * w - wprev + y * f;                                    /* :492-495 */
    f = nf_f(&g, l, E, match);
    y = GETU(w, f);
    const double factor = y / psi[match];                          /* :497-501 */
    psi[match] = y;
    for (long i = match + 1; i <= start; ++i) psi[i] *= factor;
    return match;
}

double orc_simpson38(double step, const double* v, int n)
{   /* Integral.h:50-73 — note the closing weights are NOT a valid 3/8 panel when (n-1)%3 != 0 */
    double ends = v[0] + v[n - 1];
    double s3 = 0, s2 = 0;
    for (int i = 1; i < n - 1; ++i) {
        if (i % 3 == 0) s2 += v[i]; else s3 += v[i];
    }
    ends += 3. * s3 + 2. * s2;
    return ends * step * (3. / 8.);
}

/* The other quadratures of Integral.h (no call site in the reference; north_star names Romberg).  rule: 0 Trapezoid :11-23,
 * 1 SimpsonOneThird :25-48, 2 Simpson38 :50-73, 3 Boole :75-104, 4 Romberg :106-155 (err 1e-18, minSteps 3). */
double orc_integrate(int rule, double step, const double* v, int n)
{
    const int szm1 = n - 1;
    if (rule == 0) {                                               /* :15-22 */
        double sum = 0.5 * (v[0] + v[n - 1]);
        for (int i = 1; i < szm1; ++i) sum += v[i];
        return sum * step;
    }
    if (rule == 1) {                                               /* :30-47 */
        double sum = v[0] + v[n - 1], sum4 = 0, sum2 = 0;
        for (int i = 1; i < szm1; ++i) {
            sum4 += v[i++];
            if (i < szm1) sum2 += v[i];
        }
        sum += 4. * sum4 + 2. * sum2;
        return sum * step * (1. / 3.);
    }
    if (rule == 2) return orc_simpson38(step, v, n);
    if (rule == 3) {                                               /* :80-103 */
        double sum = 7. * (v[0] + v[n - 1]), sum32 = 0, sum12 = 0, sum14 = 0;
        for (int i = 1; i < szm1; ++i) {
            sum32 += v[i++];
            if (i < szm1) {
                if (i % 4 == 0) sum14 += v[i]; else sum12 += v[i];
            }
        }
        sum += 32. * sum32 + 12. * sum12 + 14. * sum14;
        return sum * step * (2. / 45.);
    }
    {                                                              /* Romberg :108-154 */
        const double err = 1E-18;
        const int minSteps = 3;
        const int numPoints = n - 1;
        int m = numPoints, cnt = 0;
        while (m) { ++cnt; m >>= 1; }
        double Rprev[64], Rcur[64];
        for (int i = 0; i < 64; ++i) Rprev[i] = Rcur[i] = 0;
        double h = step * numPoints;
        Rprev[0] = 0.5 * h * (v[0] + v[numPoints]);
        m = numPoints;
        for (int i = 1; i < cnt; ++i) {
            const int oldStep = m;
            m >>= 1;
            double sum = 0;
            for (int j = m; j < numPoints; j += oldStep) sum += v[j];
            h *= 0.5;
            Rcur[0] = 0.5 * Rprev[0] + h * sum;
            double nk = 1;
            for (int q = 1; q <= i; ++q) {
                nk *= 4;
                Rcur[q] = Rcur[q - 1] + (Rcur[q - 1] - Rprev[q - 1]) / (nk - 1);
            }
            if (i >= minSteps && fabs(Rcur[i] - Rprev[i - 1]) < err) return Rcur[i];
            for (int q = 0; q < 64; ++q) { const double t = Rcur[q]; Rcur[q] = Rprev[q]; Rprev[q] = t; }     /* Rcur.swap(Rprev) */
        }
        return Rprev[cnt - 1];                                     /* Rprev.back() */
    }
}


void orc_normalize(double* psi, int n_nodes, double rp, double delta)
{   /* DFTAtom.cpp:36-56: y -> u = y e^{i delta/2}; integral of u^2 dr with dr = rp delta e^{delta i} di */
    double* sq = (double*)malloc(sizeof(double) * (size_t)n_nodes);
    for (int i = 0; i < n_nodes; ++i) {
        psi[i] *= exp(i * delta * 0.5);
        sq[i] = psi[i] * psi[i];
        const double jac = rp * delta * exp(delta * i);
        sq[i] *= jac;
    }
    const double norm = 1. / sqrt(orc_simpson38(1, sq, n_nodes));
    for (int i = 0; i < n_nodes; ++i) psi[i] *= norm;
    free(sq);
}

double orc_level_search(const double* V, int n_nodes, double delta, double max_r, int n0, int l, double* bottom, int* converged)
{   /* DFTAtom.cpp:497-541 (LoopOverLevels body up to "BottomEnergy = level.E - 3") + LocateInterval :566-604 */
    const double tol = 1E-12;
    const int want = n0 - l;
    double top = 50;

    /* LocateInterval */
    double hi = top, lo = *bottom;
    while (hi - lo > tol) {
        const double E = (hi + lo) / 2;
        if (orc_numerov_count_nodes(V, n_nodes, delta, max_r, l, E, want) > want) hi = E; else lo = E;
    }
    top = hi;
    lo = *bottom;
    while (hi - lo > tol) {
        const double E = (hi + lo) / 2;
        if (orc_numerov_count_nodes(V, n_nodes, delta, max_r, l, E, want) < want) lo = E; else hi = E;
    }
    double bot = hi;

    /* shooting bisection :513-534 */
    double y0 = orc_numerov_y0(V, n_nodes, delta, max_r, l, bot);
    const int sign_bottom = y0 > 0;
    int ok = 0;
    for (int it = 0; it < 500; ++it) {
        const double E = (top + bot) / 2;
        y0 = orc_numerov_y0(V, n_nodes, delta, max_r, l, E);
        if ((y0 > 0) == sign_bottom) bot = E; else top = E;
        const double a = fabs(y0);
        if (top - bot < tol && !isnan(a) && a < 1E15) { ok = 1; break; }
    }
    *converged = ok;
    *bottom = bot - 3;                                             /* :541 */
    return bot;                                                    /* level.E = BottomEnergy :534 */
}

/* ------------------------------------------------------------------------------------------------
 * L1: VWN exchange-correlation (VWNExcCor.h, ExcCorBase.h)
 * ---------------------------------------------------------------------------------------------- */

typedef struct { double A, y0, b, c, Y0; } vwn_set;
static vwn_set vwn_P(void) { vwn_set s = { 0.0310907, -0.10498, 3.72744, 12.93532, 0 }; s.Y0 = s.y0 * s.y0 + s.b * s.y0 + s.c; return s; }   /* :24-29 */
static vwn_set vwn_F(void) { vwn_set s = { 0.01554535, -0.325, 7.06042, 18.0578, 0 }; s.Y0 = s.y0 * s.y0 + s.b * s.y0 + s.c; return s; }     /* :31-36 */
static vwn_set vwn_A(void) { vwn_set s = { -1. / (6. * M_PI * M_PI), -0.0047584, 1.13107, 13.0045, 0 }; s.Y0 = s.y0 * s.y0 + s.b * s.y0 + s.c; return s; } /* :38-42 */

static double vwn_eps(double y, double dy, const vwn_set* p, double Y)
{   /* VWNExcCor.h:43-50, VWN eq. B.5 */
    const double Q = sqrt(4 * p->c - p->b * p->b);
    const double at = atan(Q / (2. * y + p->b));
    return p->A * (log(y * y / Y) + 2. * p->b / Q * at - p->b * p->y0 / p->Y0 * (log(dy * dy / Y) + 2. * (p->b + 2. * p->y0) / Q * at));
}

static double vwn_deps(double y, double dy, const vwn_set* p, double Y)
{   /* VWNExcCor.h:52-55, eq. B.6 */
    return p->A * (p->c * dy - p->b * p->y0 * y) / (dy * Y);
}

void orc_vwn_lda(const double* rho, int n, double* vexc, double* eexcdif)
{   /* VWNExcCor.h:73-101 (Vexc) and :103-128 (eexcDif) */
    const double third = 1. / 3.;
    const double X1 = pow(3. / (2. * M_PI), 2. * third);
    const double X1q = 0.25 * pow(3. / (2. * M_PI), 2. * third);
    const vwn_set P = vwn_P();
    for (int i = 0; i < n; ++i) {
        const double ro = rho[i];
        if (ro < 1E-18) { vexc[i] = 0.; eexcdif[i] = 0.; continue; }
        const double rs = pow(3. / (FOUR_PI * ro), third);
        const double y = sqrt(rs);
        const double Y = y * y + P.b * y + P.c;
        const double dy = y - P.y0;
        vexc[i] = -X1 / rs + vwn_eps(y, dy, &P, Y) - third * vwn_deps(y, dy, &P, Y);
        eexcdif[i] = X1q / rs + third * vwn_deps(y, dy, &P, Y);
    }
}

/* Chachiyo's correlation with Dirac exchange (ExcCor.h:27-95): reachable from no option of the reference (only from comments,
 * DFTAtom.cpp:383,412,421); restated for completeness of the XC family.  improved: 0 = parameters of ExcCor.h:12-17
 * (b = 20.4562557), 1 = :20-25 (b = 21.7392245). */
void orc_xc_chachiyo(const double* rho, int n, int improved, double* vexc, double* eexcdif)
{
    const double a = (M_LN2 - 1.) / (2. * M_PI * M_PI);            /* :30 */
    const double b = improved ? 21.7392245 : 20.4562557;
    const double X1v = pow(3. / (2. * M_PI), 2. / 3.);             /* :43 */
    const double X1e = 0.25 * pow(3. / (2. * M_PI), 2. / 3.);      /* :73 */
    for (int i = 0; i < n; ++i) {
        const double ro = rho[i];
        if (ro < 1E-18) { vexc[i] = 0.; eexcdif[i] = 0.; continue; }
        const double rs = pow(3. / (FOUR_PI * ro), 1. / 3.);
        const double bprs = b / rs;
        const double bprs2 = bprs / rs;
        vexc[i] = -X1v / rs + a * log(1. + bprs + bprs / rs) - a / (1. + bprs + bprs2) * (bprs + 2. * bprs2) * rs / 3.;   /* :59-62 */
        eexcdif[i] = X1e / rs + a / (1. + bprs + bprs2) * (bprs + 2. * bprs2) * rs / 3.;                                  /* :89-91 */
    }
}

static double spin_f(double z)
{   /* ExcCorBase.h:14-19 */
    const double third = 1. / 3.;
    const double mul = 1. / (2. * (pow(2., third) - 1.));
    return mul * (pow(1. + z, 4. * third) + pow(1. - z, 4. * third) - 2.);
}
static double spin_df(double z)
{   /* ExcCorBase.h:21-26 */
    const double third = 1. / 3.;
    const double mul = 2. / (3. * (pow(2., third) - 1.));
    return mul * (pow(1. + z, third) - pow(1. - z, third));
}

void orc_vwn_lsda(const double* na, const double* nb, int n, double* va, double* vb, double* vexc, double* eexcdif)
{   /* VWNExcCor.h:134-240 (potentials) and :242-312 (eexcDif) */
    const double third = 1. / 3.;
    const double X1 = pow(3. / (2. * M_PI), 2. * third);
    const double X2 = pow(2., third);
    const double X12 = X1 * X2;
    const double X1d = 0.25 * pow(3. / (2. * M_PI), 2. * third);
    const double fdd = 4. / (9. * (pow(2., third) - 1.));
    const vwn_set P = vwn_P(), F = vwn_F(), A = vwn_A();
    for (int i = 0; i < n; ++i) {
        const double roa = na[i], rob = nb[i];
        const double tot = roa + rob;
        if (tot < 1E-18) { va[i] = vb[i] = vexc[i] = eexcdif[i] = 0.; continue; }
        const double rs = pow(3. / (FOUR_PI * tot), third);
        const double rsa = pow(3. / (FOUR_PI * roa), third);
        const double rsb = pow(3. / (FOUR_PI * rob), third);
        const double exp_ = -X1 / rs;
        const double exf = X2 * exp_;
        const double exdif = exf - exp_;
        const double exfa = -X12 / rsa;
        const double exfb = -X12 / rsb;
        const double zeta = (roa - rob) / tot;
        const double z3 = zeta * zeta * zeta;
        const double z4 = z3 * zeta;
        const double fv = spin_f(zeta);
        const double dfv = spin_df(zeta);
        const double y = sqrt(rs);
        const double YP = y * (y + P.b) + P.c, dP = y - P.y0;
        const double YF = y * (y + F.b) + F.c, dF = y - F.y0;
        const double YA = y * (y + A.b) + A.c, dA = y - A.y0;
        const double ecp = vwn_eps(y, dP, &P, YP);
        const double ecf = vwn_eps(y, dF, &F, YF);
        const double eca = vwn_eps(y, dA, &A, YA);
        const double ecpd = vwn_deps(y, dP, &P, YP);
        const double ecfd = vwn_deps(y, dF, &F, YF);
        const double ecad = vwn_deps(y, dA, &A, YA);
        const double dfp = ecf - ecp;
        const double beta = fdd * dfp / eca - 1.;
        const double opbz4 = 1. + beta * z4;
        const double interp = fv / fdd * opbz4;
        const double deltaec = eca * interp;
        const double betad = fdd / eca * (ecfd - ecpd - ecad * dfp / eca);
        const double interpd = fv / fdd * z4 * betad;
        const double deriv = third * (ecpd + ecad * interp + eca * interpd);
        const double dterm = eca / fdd * (4. * beta * z3 * fv + opbz4 * dfv);
        const double core = ecp + deltaec - deriv;
        va[i] = exfa + core + (1. - zeta) * dterm;
        vb[i] = exfb + core - (1. + zeta) * dterm;
        vexc[i] = core + (exp_ + exdif * fv);
        /* eexcDif :254-308 */
        const double expd = X1d / rs;
        const double exfd = X2 * expd;
        eexcdif[i] = expd + (exfd - expd) * fv + deriv;
    }
}

/* ------------------------------------------------------------------------------------------------
 * L1: radial Poisson multigrid (PoissonSolver.h / PoissonSolver.cpp)
 * ---------------------------------------------------------------------------------------------- */

typedef struct {
    int L;
    int* size;          /* size[l] = 2^(L-l) + 1 */
    double** phi;
    double** src;
    double* dl;         /* first-derivative coefficient per level: delta * 2^l (.cpp:21-26) */
    double lo_bc, hi_bc;
} mgrid;

static mgrid* mg_new(int L, double delta)
{   /* PoissonSolver.cpp:8-27 */
    mgrid* m = (mgrid*)calloc(1, sizeof(mgrid));
    m->L = L;
    m->size = (int*)malloc(sizeof(int) * (size_t)L);
    m->phi = (double**)malloc(sizeof(double*) * (size_t)L);
    m->src = (double**)malloc(sizeof(double*) * (size_t)L);
    m->dl = (double*)malloc(sizeof(double) * (size_t)L);
    int sz = 3;
    for (int l = L - 1; l >= 0; --l) {
        m->size[l] = sz;
        m->phi[l] = (double*)calloc((size_t)sz, sizeof(double));
        m->src[l] = (double*)calloc((size_t)sz, sizeof(double));
        sz = 2 * sz - 1;
    }
    double d = delta;
    for (int l = 0; l < L; ++l) { m->dl[l] = d; d *= 2; }
    return m;
}

static void mg_free(mgrid* m)
{
    for (int l = 0; l < m->L; ++l) { free(m->phi[l]); free(m->src[l]); }
    free(m->phi); free(m->src); free(m->size); free(m->dl); free(m);
}

static double mg_sweep(mgrid* m, int l)
{   /* GaussSeidel, PoissonSolver.cpp:40-64: lexicographic, in place */
    double* p = m->phi[l];
    const double* s = m->src[l];
    const double d = m->dl[l];
    double acc = 0;
    for (int i = 1; i < m->size[l] - 1; ++i) {
        const double old = p[i];
        p[i] = 0.5 * (s[i] + p[i - 1] + p[i + 1] - d * (p[i + 1] - p[i - 1]) * 0.5);
        const double dif = old - p[i];
        acc += dif * dif;
    }
    return sqrt(acc);
}

static double mg_smooth(mgrid* m, int l, double tol, int sweeps)
{   /* IterateGaussSeidel, .cpp:66-77 */
    double err = 1E10;
    for (int k = 0; k < sweeps; ++k) { err = mg_sweep(m, l); if (err < tol) break; }
    return err;
}

static void mg_restrict_to(mgrid* m, int l)
{   /* Restrict(lvl), .cpp:126-157: injected, rescaled residual of level l-1; coarse phi zeroed */
    const double* pf = m->phi[l - 1];
    const double* sf = m->src[l - 1];
    double* pc = m->phi[l];
    double* sc = m->src[l];
    const int nc = m->size[l];
    for (int i = 0; i < nc; ++i) pc[i] = 0;
    for (int i = 1; i < nc - 1; ++i) {
        const int k = 2 * i;
        sc[i] = 4. * (sf[k] + pf[k - 1] - 2. * pf[k] + pf[k + 1]) - m->dl[l] * (pf[k + 1] - pf[k - 1]);
    }
    sc[0] = sc[nc - 1] = 0;
}

static void mg_prolong_from(mgrid* m, int l)
{   /* Prolong(Phi[l] -> Phi[l-1]), .cpp:110-123: linear interpolation, additive correction */
    const double* pc = m->phi[l];
    double* pf = m->phi[l - 1];
    pf[0] += pc[0];
    for (int i = 1; i < m->size[l]; ++i) {
        pf[2 * i] += pc[i];
        pf[2 * i - 1] += 0.5 * (pc[i - 1] + pc[i]);
    }
}

static void mg_to_coarse(mgrid* m, int from, int to, double tol, int sweeps)
{   /* "Ascend", .cpp:162-171 */
    for (int l = from; l < to;) { mg_smooth(m, l, tol, sweeps); mg_restrict_to(m, ++l); }
    mg_smooth(m, to, tol, sweeps);
}

static double mg_to_fine(mgrid* m, int from, int to, double tol, int sweeps)
{   /* "Descend", .cpp:173-186 */
    double err = 1E10;
    for (int l = from; l > to; --l) { mg_prolong_from(m, l); err = mg_smooth(m, l - 1, tol, sweeps); }
    return err;
}

static void mg_initialize(mgrid* m, double tol)
{   /* Initialize, .cpp:80-106 */
    memset(m->phi[0], 0, sizeof(double) * (size_t)m->size[0]);
    for (int l = 1; l < m->L; ++l) {
        const int last = m->size[l] - 1;
        for (int p = 1; p < last; ++p) { m->src[l][p] = 4 * m->src[l - 1][2 * p]; m->phi[l][p] = 0; }
        m->src[l][0] = m->src[l][last] = 0;
        m->phi[l][0] = m->phi[l][last] = 0;
    }
    const int c = m->L - 1;
    m->phi[c][0] = m->lo_bc;
    m->phi[c][m->size[c] - 1] = m->hi_bc;
    mg_smooth(m, c, tol, 15);
}

static int mg_full_cycle(mgrid* m, double tol, double tol_last, int max_vcycles, double* errs)
{   /* FullCycle, PoissonSolver.h:89-124 */
    const int sweeps = 3, c = m->L - 1;
    mg_initialize(m, tol);
    for (int l = m->L - 2; l > 0; --l) {
        mg_to_fine(m, c, l, tol, sweeps);
        mg_to_coarse(m, l, c, tol, sweeps);
    }
    mg_to_fine(m, c, 0, tol_last, sweeps);
    int k = 0;
    for (; k < max_vcycles; ++k) {
        mg_to_coarse(m, 0, c, tol_last, sweeps);                 /* VCycle :155-159 */
        const double err = mg_to_fine(m, c, 0, tol_last, sweeps);
        if (errs) errs[k] = err;
        if (err < tol_last) { ++k; break; }
    }
    return k;
}

void orc_poisson(int levels, double delta, int Z, double max_r, const double* rho, double* U,
                 int max_vcycles, double* vcycle_err, int* n_vcycles)
{   /* SolvePoissonNonUniform, PoissonSolver.h:51-81 */
    mgrid* m = mg_new(levels, delta);
    const int n = m->size[0];
    double* S = m->src[0];
    const double rp0 = max_r / (exp((n - 1) * delta) - 1.);            /* FillRNonuniformR .cpp:212-223 */
    for (int i = 0; i < n; ++i) S[i] = rp0 * (exp(i * delta) - 1.);
    const double rp = max_r / (exp(((double)n - 1.) * delta) - 1.);    /* .h:65 */
    const double c = FOUR_PI * (rp * rp * (delta * delta));
    const double two_delta = 2. * delta;
    for (int i = 1; i < n - 1; ++i) S[i] *= c * exp(i * two_delta) * rho[i];
    m->lo_bc = 0; m->hi_bc = Z;
    const int k = mg_full_cycle(m, 1E-3, 1E-14, max_vcycles, vcycle_err);
    if (n_vcycles) *n_vcycles = k;
    memcpy(U, m->phi[0], sizeof(double) * (size_t)n);
    mg_free(m);
}

double orc_poisson_vcycles(int levels, double delta, double* phi0, const double* src0, int n_cycles)
{
    mgrid* m = mg_new(levels, delta);
    const int n = m->size[0], c = levels - 1;
    memcpy(m->phi[0], phi0, sizeof(double) * (size_t)n);
    memcpy(m->src[0], src0, sizeof(double) * (size_t)n);
    double err = 0;
    for (int k = 0; k < n_cycles; ++k) {
        mg_to_coarse(m, 0, c, 1E-14, 3);
        err = mg_to_fine(m, c, 0, 1E-14, 3);
    }
    memcpy(phi0, m->phi[0], sizeof(double) * (size_t)n);
    mg_free(m);
    return err;
}

/* ------------------------------------------------------------------------------------------------
 * L2: SCF driver (DFTAtom.cpp:346-491 LDA, :847-1022 LSDA)
 * ---------------------------------------------------------------------------------------------- */

static int by_energy(const void* pa, const void* pb)
{
    const double a = ((const orc_level*)pa)->E, b = ((const orc_level*)pb)->E;
    return (a > b) - (a < b);
}

/* stable sort by E like std::sort's result for distinct keys; ties are not expected */
static void sort_levels_by_energy(orc_level* lv, int n) { qsort(lv, (size_t)n, sizeof(orc_level), by_energy); }

/* LoopOverLevels (:493-563) + the tail of CalculateNonUniformDensity (:332-342) for one spin channel */
static void spin_channel_density(const double* V, int n, double delta, double max_r, double rp, double mixing,
                                 orc_level* lv, int n_lv, double Z, double* rho, double* acc, double* psi,
                                 double* e_el, int* all_converged)
{
    double bottom = -Z * Z - 1.;                                   /* :407 */
    memset(acc, 0, sizeof(double) * (size_t)n);
    for (int k = 0; k < n_lv; ++k) {
        int ok = 0;
        lv[k].E = orc_level_search(V, n, delta, max_r, lv[k].n0, lv[k].l, &bottom, &ok);
        if (!ok) *all_converged = 0;
        orc_numerov_match(V, n, delta, max_r, lv[k].l, lv[k].E, psi);      /* :545 */
        orc_normalize(psi, n, rp, delta);                                   /* :546 */
        for (int i = 0; i < n - 1; ++i) acc[i] += lv[k].occ * psi[i] * psi[i];   /* :558-559 */
        *e_el += lv[k].occ * lv[k].E;                                        /* :561 */
    }
    const double keep = mixing, take = 1. - mixing;
    for (int i = 1; i < n; ++i) {                                  /* :332-342 */
        const double r = rp * (exp(i * delta) - 1.);
        acc[i] /= FOUR_PI * r * r;
        rho[i] = keep * rho[i] + take * acc[i];
    }
}

int orc_scf(const orc_options* opt, orc_result* res, orc_step_cb cb, void* user, int max_vcycles)
{
    return orc_scf_ex(opt, res, cb, user, max_vcycles, NULL, NULL);
}

/* same, additionally exporting the final potentials V_out[2][N] and spin densities rho_out[2][N] (either may be NULL) */
int orc_scf_ex(const orc_options* opt, orc_result* res, orc_step_cb cb, void* user, int max_vcycles, double* V_out, double* rho_out)
{
    const int Z = opt->Z, lsda = opt->method != 0;
    const int n = orc_n_nodes(opt->levels);
    const double delta = opt->delta, max_r = opt->max_r;
    const double rp = max_r / (exp((n - 1) * delta) - 1.);

    orc_level all[ORC_MAX_LEVELS];
    const int n_all = orc_aufbau(Z, all);
    orc_level lv[2][ORC_MAX_LEVELS];
    int n_lv[2] = { n_all, 0 };
    int n_el[2] = { Z, 0 };
    if (lsda) orc_split_spin(Z, all, n_all, lv[0], &n_lv[0], lv[1], &n_lv[1], &n_el[0], &n_el[1]);
    else memcpy(lv[0], all, sizeof(orc_level) * (size_t)n_all);
    const int n_spin = lsda ? 2 : 1;

    const size_t bytes = sizeof(double) * (size_t)n;
    double* rho_s[2] = { (double*)calloc(1, bytes), (double*)calloc(1, bytes) };
    double* rho = lsda ? (double*)calloc(1, bytes) : rho_s[0];
    double* V[2] = { (double*)calloc(1, bytes), (double*)calloc(1, bytes) };
    double* vxc_s[2] = { (double*)calloc(1, bytes), (double*)calloc(1, bytes) };
    double* U = (double*)calloc(1, bytes), *vexc = (double*)calloc(1, bytes), *edif = (double*)calloc(1, bytes);
    double* acc = (double*)calloc(1, bytes), *psi = (double*)calloc(1, bytes);
    double* g_nuc = (double*)calloc(1, bytes), *g_xc = (double*)calloc(1, bytes), *g_dif = (double*)calloc(1, bytes);
    double* g_har = (double*)calloc(1, bytes), *g_pot = (double*)calloc(1, bytes);

    /* initial guess: uniform charge in the sphere, :371-376 / :876-884 */
    const double volume = FOUR_PI / 3. * max_r * max_r * max_r;
    if (lsda) {
        const double ca = n_el[0] / volume, cbeta = n_el[1] / volume;
        for (int i = 1; i < n; ++i) { rho_s[0][i] = ca; rho_s[1][i] = cbeta; rho[i] = ca + cbeta; }
    } else {
        const double c = Z / volume;
        for (int i = 1; i < n; ++i) rho[i] = c;
    }
    orc_poisson(opt->levels, delta, Z, max_r, rho, U, max_vcycles, NULL, NULL);
    if (lsda) orc_vwn_lsda(rho_s[0], rho_s[1], n, vxc_s[0], vxc_s[1], vexc, edif);
    else orc_vwn_lda(rho, n, vexc, edif);
    for (int i = 1; i < n; ++i) {                                  /* :387-392 / :895-904 */
        const double r = rp * (exp(i * delta) - 1.);
        if (lsda) { const double uc = (-Z + U[i]) / r; V[0][i] = uc + vxc_s[0][i]; V[1][i] = uc + vxc_s[1][i]; }
        else V[0][i] = (-Z + U[i]) / r + vexc[i];
    }

    const int max_steps = lsda ? 150 : 100;                        /* :396 / :908 */
    double e_old = 0;
    int prev_ok = 0;
    memset(res, 0, sizeof(*res));
    orc_step st;
    for (int sp = 0; sp < max_steps; ++sp) {
        memset(&st, 0, sizeof(st));
        st.step = sp;
        double e_el = 0;
        int ok = 1;
        for (int s = 0; s < n_spin; ++s)
            spin_channel_density(V[s], n, delta, max_r, rp, opt->mixing, lv[s], n_lv[s], (double)Z, rho_s[s], acc, psi, &e_el, &ok);
        if (lsda) for (int i = 1; i < n; ++i) rho[i] = rho_s[0][i] + rho_s[1][i];   /* :933-934 */

        orc_poisson(opt->levels, delta, Z, max_r, rho, U, max_vcycles, NULL, NULL);
        if (lsda) orc_vwn_lsda(rho_s[0], rho_s[1], n, vxc_s[0], vxc_s[1], vexc, edif);
        else orc_vwn_lda(rho, n, vexc, edif);

        g_nuc[0] = g_xc[0] = g_dif[0] = g_har[0] = g_pot[0] = 0;
        V[0][0] = V[1][0] = 0;
        for (int i = 1; i < n; ++i) {
            const double ex = exp(delta * i);
            const double r = rp * (ex - 1.);
            const double jac = rp * delta * ex;
            if (!lsda) {                                           /* :437-457 */
                V[0][i] = (-Z + U[i]) / r + vexc[i];
                const double rd = r * rho[i] * jac;
                g_nuc[i] = Z * rd;
                const double r2d = r * r * rho[i] * jac;
                g_xc[i] = r2d * vexc[i];
                g_dif[i] = r2d * edif[i];
                g_har[i] = rd * U[i];
                g_pot[i] = r2d * V[0][i];
            } else {                                               /* :956-983 */
                const double uc = (-Z + U[i]) / r;
                V[0][i] = uc + vxc_s[0][i];
                V[1][i] = uc + vxc_s[1][i];
                const double rj = r * jac;
                const double rd = rj * rho[i];
                g_nuc[i] = Z * rd;
                const double r2j = r * rj;
                const double r2d = r2j * rho[i];
                g_xc[i] = r2d * vexc[i];
                g_dif[i] = r2d * edif[i];
                g_har[i] = rd * U[i];
                g_pot[i] = r2j * rho_s[0][i] * V[0][i] + r2j * rho_s[1][i] * V[1][i];
            }
        }
        const double e_nuc = -FOUR_PI * orc_simpson38(1, g_nuc, n);    /* :459-470 */
        double e_xc = FOUR_PI * orc_simpson38(1, g_xc, n);
        const double e_dif = FOUR_PI * orc_simpson38(1, g_dif, n);
        e_xc += e_dif;
        const double e_har = -2 * M_PI * orc_simpson38(1, g_har, n);
        const double e_pot = FOUR_PI * orc_simpson38(1, g_pot, n);
        st.Ekin = e_el - e_pot;
        st.Etotal = e_el + e_har + e_dif;
        st.Ecoul = -e_har;
        st.Eenuc = e_nuc;
        st.Exc = e_xc;
        st.level_search_converged = ok;
        for (int s = 0; s < n_spin; ++s) { st.n_levels[s] = n_lv[s]; memcpy(st.lv[s], lv[s], sizeof(orc_level) * (size_t)n_lv[s]); }
        if (cb) cb(&st, user);
        res->n_steps = sp + 1;
        if (fabs((e_old - st.Etotal) / st.Etotal) < 1E-11 && ok && prev_ok) { res->finished = 1; break; }   /* :474-479 */
        e_old = st.Etotal;
        prev_ok = ok;
    }
    res->last = st;
    for (int s = 0; s < n_spin; ++s) {
        res->n_sorted[s] = n_lv[s];
        memcpy(res->sorted[s], lv[s], sizeof(orc_level) * (size_t)n_lv[s]);
        sort_levels_by_energy(res->sorted[s], n_lv[s]);            /* :487 / :1012-1013 */
    }

    if (V_out) { memcpy(V_out, V[0], bytes); memcpy(V_out + n, V[1], bytes); }
    if (rho_out) { memcpy(rho_out, rho_s[0], bytes); memcpy(rho_out + n, rho_s[1], bytes); }
    free(rho_s[0]); free(rho_s[1]); if (lsda) free(rho);
    free(V[0]); free(V[1]); free(vxc_s[0]); free(vxc_s[1]); free(U); free(vexc); free(edif); free(acc); free(psi);
    free(g_nuc); free(g_xc); free(g_dif); free(g_har); free(g_pot);
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * The uniform-grid pair: CalculateUniformLDA DFTAtom.cpp:60-210, CalculateUniformLSDA :646-844,
 * LoopOverLevels :213-284, LocateInterval :287-325, NormalizeUniform :21-32, Numerov<NumerovFunctionRegularGrid>
 * (Numerov.h:16-70 function object, :272-504 sweeps, IsUniform() branches), SolvePoissonUniform PoissonSolver.h:20-49.
 * Public API of the reference without a live caller; restated here ahead of its kernels and pinned digit for digit on
 * tests/golden/uniform.json (generated by the unmodified reference, oracle/_ref/dftatom_ref methods 2 / 3).
 * ---------------------------------------------------------------------------------------------- */

static double uf_veff(const double* V, int l, double pos, long i) { return V[i] + l * (l + 1.) / (pos * pos) * 0.5; }   /* Numerov.h:21-24 */
static double uf_f(const double* V, int l, double E, double pos, long i) { return 2. * (uf_veff(V, l, pos, i) - E); }   /* :26-31 */
static double uf_far(double pos, double E) { return exp(-pos * sqrt(2. * fabs(E))); }                                   /* :33-36 */
static double uf_near(double pos, int l) { return pow(pos, (double)l + 1.); }                                           /* :38-41 */
static double uf_maxr(double E) { return 200. / sqrt(2. * fabs(E)); }                                                   /* :53-56 */

int orc_u_count_nodes(const double* V, double start_point, int l, double E, long steps, int limit)
{   /* Numerov.h:272-349, IsUniform() branch */
    const double h = start_point / steps, h2 = h * h, h2p12 = h2 / 12.;
    start_point = fmin(start_point, uf_maxr(E));
    steps = (long)(start_point / h);
    double position = start_point;
    double solution = uf_far(position, E);
    double prev = solution;
    double f = uf_f(V, l, E, position, steps);
    double wprev = (1 - h2p12 * f) * solution;
    position -= h;
    solution = uf_far(position, E);
    f = uf_f(V, l, E, position, steps - 1);
    double w = (1 - h2p12 * f) * solution;
    int old_sign = solution > 0, count = 0, seen_allowed = 0;
    for (long i = steps - 2; i > 0; --i) {
        const double wnext = 2. * w - wprev + h2 * solution * f;
        position = h * i;
        wprev = w; w = wnext;
        f = uf_f(V, l, E, position, i);
        prev = solution;
        solution = w / (1. - h2p12 * f);
        if (fabs(solution) == INFINITY) return count;
        if ((solution > 0) != old_sign) {
            if (++count > limit) return count;
            old_sign = !old_sign;
        }
        const double veff = uf_veff(V, l, position, i);
        if (veff <= E) seen_allowed = 1;
        else if (seen_allowed) return count;
    }
    if (count <= limit) {
        solution = solution * (2 + h2 * f) - prev;
        if ((solution > 0) != old_sign) ++count;
    }
    return count;
}

double orc_u_y0(const double* V, double start_point, int l, double E, long steps)
{   /* Numerov.h:351-401, IsUniform() branch */
    const double h = start_point / steps, h2 = h * h, h2p12 = h2 / 12.;
    start_point = fmin(start_point, uf_maxr(E));
    steps = (long)(start_point / h);
    double position = start_point;
    double solution = uf_far(position, E);
    double prev = solution;
    double f = uf_f(V, l, E, position, steps);
    double wprev = (1 - h2p12 * f) * solution;
    position -= h;
    solution = uf_far(position, E);
    f = uf_f(V, l, E, position, steps - 1);
    double w = (1 - h2p12 * f) * solution;
    for (long i = steps - 2; i > 0; --i) {
        const double wnext = 2. * w - wprev + h2 * solution * f;
        position = h * i;
        wprev = w; w = wnext;
        f = uf_f(V, l, E, position, i);
        prev = solution;
        solution = w / (1. - h2p12 * f);
    }
    return solution * (2 + h2 * f) - prev;
}

long orc_u_match(const double* V, double start_point, int l, double E, long steps, double* psi)
{   /* Numerov.h:403-504, IsUniform() branch, including the re-computation of h from the truncated range (:430-432) */
    const long high = steps + 1;
    double h = start_point / steps, h2, h2p12;
    start_point = fmin(start_point, uf_maxr(E));
    steps = (long)(start_point / h);
    for (long i = steps + 1; i < high; ++i) psi[i] = 0;
    h = start_point / steps;
    h2 = h * h;
    h2p12 = h2 / 12.;
    const long size = steps + 1;
    double position = start_point;
    double solution = uf_far(position, E);
    psi[steps] = solution;
    double f = uf_f(V, l, E, position, steps);
    double wprev = (1 - h2p12 * f) * solution;
    position -= h;
    psi[steps - 1] = solution = uf_far(position, E);
    f = uf_f(V, l, E, position, steps - 1);
    double w = (1 - h2p12 * f) * solution;
    long match = 2;
    for (long i = steps - 2; i > 0; --i) {
        const double wnext = 2. * w - wprev + h2 * solution * f;
        position = h * i;
        wprev = w; w = wnext;
        f = uf_f(V, l, E, position, i);
        psi[i] = solution = w / (1. - h2p12 * f);
        if (solution < psi[i + 1] || fabs(solution) > 1E15) { match = i; break; }
    }
    position = 0;
    psi[0] = solution = 0;
    wprev = 0;
    position += h;
    psi[1] = solution = uf_near(position, l);
    f = uf_f(V, l, E, position, 1);
    w = (1 - h2p12 * f) * solution;
    for (long i = 2; i < match; ++i) {
        const double wnext = 2. * w - wprev + h2 * solution * f;
        position = h * i;
        wprev = w; w = wnext;
        f = uf_f(V, l, E, position, i);
        psi[i] = solution = w / (1. - h2p12 * f);
    }
    w = 2. * w - wprev + h2 * solution * f;
    position = h * match;
    f = uf_f(V, l, E, position, match);
    solution = w / (1. - h2p12 * f);
    const double factor = solution / psi[match];
    psi[match] = solution;
    for (long i = match + 1; i < size; ++i) psi[i] *= factor;
    return match;
}

void orc_poisson_uniform(int levels, int Z, double max_r, const double* rho, double* U, int max_vcycles)
{   /* SolvePoissonUniform, PoissonSolver.h:20-49 (constructor with dGrid = 0: no first-derivative term) */
    mgrid* m = mg_new(levels, 0.);
    const int n = m->size[0];
    double* S = m->src[0];
    const size_t N = (size_t)n - 1;
    for (size_t i = 0; i < (size_t)n; ++i) S[i] = (0. * (N - i) + max_r * i) / N;      /* FillR, PoissonSolver.cpp:199-209 */
    const double delta = S[1] - S[0];
    const double c = (delta * delta) * FOUR_PI;
    for (int i = 0; i < n; ++i) S[i] *= c * rho[i];
    m->lo_bc = 0; m->hi_bc = Z;
    mg_full_cycle(m, 1E-3, 1E-14, max_vcycles, NULL);
    memcpy(U, m->phi[0], sizeof(double) * (size_t)n);
    mg_free(m);
}

/* LoopOverLevels :213-284 (with LocateInterval :287-325 and NormalizeUniform :21-32) + the mixing loop of the drivers */
static void spin_channel_density_uniform(const double* V, int n, double max_r, double h, double mixing, orc_level* lv, int n_lv, double Z,
                                         double* rho, double* acc, double* psi, double* sq, double* e_el, int* all_converged)
{
    const double tol = 1E-12;
    const long n_steps = n - 1;
    double bottom = -Z * Z - 1.;
    memset(acc, 0, sizeof(double) * (size_t)n);
    for (int k = 0; k < n_lv; ++k) {
        const int l = lv[k].l, want = lv[k].n0 - l;
        double top = 50;
        double hi = top, lo = bottom;                               /* LocateInterval */
        while (hi - lo > tol) {
            const double E = (hi + lo) / 2;
            if (orc_u_count_nodes(V, max_r, l, E, n_steps, want) > want) hi = E; else lo = E;
        }
        top = hi;
        lo = bottom;
        while (hi - lo > tol) {
            const double E = (hi + lo) / 2;
            if (orc_u_count_nodes(V, max_r, l, E, n_steps, want) < want) lo = E; else hi = E;
        }
        bottom = hi;
        double y0 = orc_u_y0(V, max_r, l, bottom, n_steps);         /* :234-254 */
        const int sign_bottom = y0 > 0;
        int ok = 0;
        double E = bottom;
        for (int it = 0; it < 500; ++it) {
            E = (top + bottom) / 2;
            y0 = orc_u_y0(V, max_r, l, E, n_steps);
            if ((y0 > 0) == sign_bottom) bottom = E; else top = E;
            const double a = fabs(y0);
            if (top - bottom < tol && !isnan(a) && a < 1E15) { ok = 1; break; }
        }
        lv[k].E = bottom;                                           /* :255 */
        if (!ok) *all_converged = 0;
        bottom = lv[k].E - 3;                                       /* :262 */
        orc_u_match(V, max_r, l, lv[k].E, n_steps, psi);            /* :266 */
        for (int i = 0; i < n; ++i) sq[i] = psi[i] * psi[i];        /* NormalizeUniform */
        const double unorm = 1. / sqrt(orc_simpson38(h, sq, n));
        for (int i = 0; i < n; ++i) psi[i] *= unorm;
        for (int i = 0; i < n - 1; ++i) acc[i] += lv[k].occ * psi[i] * psi[i];    /* :279-280 */
        *e_el += lv[k].occ * lv[k].E;
    }
    const double keep = mixing, take = 1. - mixing;
    for (int i = 1; i < n; ++i) {                                   /* :127-137 / :713-718 */
        const double position = i * h;
        acc[i] /= FOUR_PI * position * position;
        rho[i] = keep * rho[i] + take * acc[i];
    }
}

/* opt->method: 2 = CalculateUniformLDA, 3 = CalculateUniformLSDA; opt->delta is not used */
int orc_scf_uniform(const orc_options* opt, orc_result* res, orc_step_cb cb, void* user, int max_vcycles)
{
    const int Z = opt->Z, lsda = opt->method == 3;
    const int n = orc_n_nodes(opt->levels);
    const double max_r = opt->max_r;
    const double h = max_r / (n - 1);

    orc_level all[ORC_MAX_LEVELS];
    const int n_all = orc_aufbau(Z, all);
    orc_level lv[2][ORC_MAX_LEVELS];
    int n_lv[2] = { n_all, 0 };
    int n_el[2] = { Z, 0 };
    if (lsda) orc_split_spin(Z, all, n_all, lv[0], &n_lv[0], lv[1], &n_lv[1], &n_el[0], &n_el[1]);
    else memcpy(lv[0], all, sizeof(orc_level) * (size_t)n_all);
    const int n_spin = lsda ? 2 : 1;

    const size_t bytes = sizeof(double) * (size_t)n;
    double* rho_s[2] = { (double*)calloc(1, bytes), (double*)calloc(1, bytes) };
    double* rho = lsda ? (double*)calloc(1, bytes) : rho_s[0];
    double* V[2] = { (double*)calloc(1, bytes), (double*)calloc(1, bytes) };
    double* vxc_s[2] = { (double*)calloc(1, bytes), (double*)calloc(1, bytes) };
    double* U = (double*)calloc(1, bytes), *vexc = (double*)calloc(1, bytes), *edif = (double*)calloc(1, bytes);
    double* acc = (double*)calloc(1, bytes), *psi = (double*)calloc(1, bytes), *sq = (double*)calloc(1, bytes);
    double* g_nuc = (double*)calloc(1, bytes), *g_xc = (double*)calloc(1, bytes), *g_dif = (double*)calloc(1, bytes);
    double* g_har = (double*)calloc(1, bytes), *g_pot = (double*)calloc(1, bytes);

    const double volume = 4. / 3. * M_PI * max_r * max_r * max_r;   /* :83 / :671 */
    if (lsda) {
        const double ca = n_el[0] / volume, cbeta = n_el[1] / volume;
        for (int i = 1; i < n; ++i) { rho_s[0][i] = ca; rho_s[1][i] = cbeta; rho[i] = ca + cbeta; }
    } else {
        const double c = Z / volume;
        for (int i = 1; i < n; ++i) rho[i] = c;
    }
    orc_poisson_uniform(opt->levels, Z, max_r, rho, U, max_vcycles);
    if (lsda) orc_vwn_lsda(rho_s[0], rho_s[1], n, vxc_s[0], vxc_s[1], vexc, edif);
    else orc_vwn_lda(rho, n, vexc, edif);
    for (int i = 1; i < n; ++i) {                                   /* :97-102 / :688-695 */
        const double pos = h * i;
        if (lsda) { const double uc = (-Z + U[i]) / pos; V[0][i] = uc + vxc_s[0][i]; V[1][i] = uc + vxc_s[1][i]; }
        else V[0][i] = (-Z + U[i]) / pos + vexc[i];
    }

    const int max_steps = lsda ? 150 : 100;                         /* :106 / :699 */
    double e_old = 0;
    int prev_ok = 0;
    memset(res, 0, sizeof(*res));
    orc_step st;
    for (int sp = 0; sp < max_steps; ++sp) {
        memset(&st, 0, sizeof(st));
        st.step = sp;
        double e_el = 0;
        int ok = 1;
        for (int s = 0; s < n_spin; ++s)
            spin_channel_density_uniform(V[s], n, max_r, h, opt->mixing, lv[s], n_lv[s], (double)Z, rho_s[s], acc, psi, sq, &e_el, &ok);
        if (lsda) for (int i = 1; i < n; ++i) rho[i] = rho_s[0][i] + rho_s[1][i];

        orc_poisson_uniform(opt->levels, Z, max_r, rho, U, max_vcycles);
        if (lsda) orc_vwn_lsda(rho_s[0], rho_s[1], n, vxc_s[0], vxc_s[1], vexc, edif);
        else orc_vwn_lda(rho, n, vexc, edif);

        g_nuc[0] = g_xc[0] = g_dif[0] = g_har[0] = g_pot[0] = 0;
        V[0][0] = V[1][0] = 0;
        for (int i = 1; i < n; ++i) {
            const double position = i * h;
            if (!lsda) {                                            /* :160-180 */
                V[0][i] = (-Z + U[i]) / position + vexc[i];
                const double pd = position * rho[i];
                g_nuc[i] = Z * pd;
                const double p2d = position * position * rho[i];
                g_xc[i] = p2d * vexc[i];
                g_dif[i] = p2d * edif[i];
                g_har[i] = pd * U[i];
                g_pot[i] = p2d * V[0][i];
            } else {                                                /* :786-806 */
                const double uc = (-Z + U[i]) / position;
                V[0][i] = uc + vxc_s[0][i];
                V[1][i] = uc + vxc_s[1][i];
                const double pd = position * rho[i];
                g_nuc[i] = Z * pd;
                const double p2 = position * position;
                const double p2d = p2 * rho[i];
                g_xc[i] = p2d * vexc[i];
                g_dif[i] = p2d * edif[i];
                g_har[i] = pd * U[i];
                g_pot[i] = p2 * (rho_s[0][i] * V[0][i] + rho_s[1][i] * V[1][i]);
            }
        }
        const double e_nuc = -FOUR_PI * orc_simpson38(h, g_nuc, n);     /* :182-194 */
        double e_xc = 4 * M_PI * orc_simpson38(h, g_xc, n);
        const double e_dif = FOUR_PI * orc_simpson38(h, g_dif, n);
        e_xc += e_dif;
        const double e_har = -2 * M_PI * orc_simpson38(h, g_har, n);
        const double e_pot = FOUR_PI * orc_simpson38(h, g_pot, n);
        st.Ekin = e_el - e_pot;
        st.Etotal = e_el + e_har + e_dif;
        st.Ecoul = -e_har;
        st.Eenuc = e_nuc;
        st.Exc = e_xc;
        st.level_search_converged = ok;
        for (int s = 0; s < n_spin; ++s) { st.n_levels[s] = n_lv[s]; memcpy(st.lv[s], lv[s], sizeof(orc_level) * (size_t)n_lv[s]); }
        if (cb) cb(&st, user);
        res->n_steps = sp + 1;
        if (fabs((e_old - st.Etotal) / st.Etotal) < 1E-11 && ok && prev_ok) { res->finished = 1; break; }   /* :198-203 */
        e_old = st.Etotal;
        prev_ok = ok;
    }
    res->last = st;
    for (int s = 0; s < n_spin; ++s) {
        res->n_sorted[s] = n_lv[s];
        memcpy(res->sorted[s], lv[s], sizeof(orc_level) * (size_t)n_lv[s]);
        sort_levels_by_energy(res->sorted[s], n_lv[s]);
    }
    free(rho_s[0]); free(rho_s[1]); if (lsda) free(rho);
    free(V[0]); free(V[1]); free(vxc_s[0]); free(vxc_s[1]); free(U); free(vexc); free(edif); free(acc); free(psi); free(sq);
    free(g_nuc); free(g_xc); free(g_dif); free(g_har); free(g_pot);
    return 0;
}

/* ---- text report, same line formats as the reference (DFTAtom.cpp:358,398,548-556,472,476,483,487-490) ---- */

typedef struct { int precision; int last_step_printed; int pending_sep; } print_ctx;

static void print_step(const orc_step* st, void* user)
{
    print_ctx* pc = (print_ctx*)user;
    static const char orb[] = "spdf";
    if (pc->pending_sep) printf("********************************************************************************\n");
    printf("Step: %d\n", st->step);
    for (int s = 0; s < 2; ++s)
        for (int k = 0; k < st->n_levels[s]; ++k)
            printf("Energy %d%c: %.*f Num nodes: %d\n", st->lv[s][k].n0 + 1, orb[st->lv[s][k].l], pc->precision, st->lv[s][k].E,
                   st->lv[s][k].n0 - st->lv[s][k].l);
    printf("Etotal = %.*f Ekin = %.*f Ecoul = %.*f Eenuc = %.*f Exc = %.*f\n", pc->precision, st->Etotal, pc->precision, st->Ekin,
           pc->precision, st->Ecoul, pc->precision, st->Eenuc, pc->precision, st->Exc);
    pc->pending_sep = 1;
}

int orc_scf_print(const orc_options* opt, int precision, int max_vcycles)
{
    static const char orb[] = "spdf";
    print_ctx pc = { precision, -1, 0 };
    orc_result res;
    printf("Computing atom with Z=%d using %s with non-uniform grid\n", opt->Z, opt->method ? "LSDA" : "LSD");
    orc_scf(opt, &res, print_step, &pc, max_vcycles);
    if (res.finished) printf("\nFinished!\n\n");
    else printf("********************************************************************************\n");
    if (opt->method) {
        printf("Alpha: ");
        for (int k = 0; k < res.n_sorted[0]; ++k) printf("%d%c%d ", res.sorted[0][k].n0 + 1, orb[res.sorted[0][k].l], res.sorted[0][k].occ);
        printf("\nBeta: ");
        for (int k = 0; k < res.n_sorted[1]; ++k) printf("%d%c%d ", res.sorted[1][k].n0 + 1, orb[res.sorted[1][k].l], res.sorted[1][k].occ);
    } else {
        for (int k = 0; k < res.n_sorted[0]; ++k) printf("%d%c%d ", res.sorted[0][k].n0 + 1, orb[res.sorted[0][k].l], res.sorted[0][k].occ);
    }
    printf("\n");
    return res.finished;
}
