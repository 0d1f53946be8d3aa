// Shared device helpers of the Numerov kernels (included by numerov.cu and numerov_fast.cu).
#pragma once
#include "internal.h"
#include <cmath>

namespace dft {

__device__ __forceinline__ int hi32(double x) { return __double2hiint(x); }

// Numerov.h:119-136: bisection on the index for far(idx) = exp(-r_idx sqrt(2|E|) - idx δ/2) < 1e-200
// Uniform grid (NumerovFunctionRegularGrid, Numerov.h:16-70): the sweeps start at startPoint = min(MaxR, 200 / sqrt(2|E|)) (:53-56,
// :278-282), steps = (long)(startPoint / h).  The two far seeds sit at the POSITIONS startPoint, startPoint - h - which are not grid
// nodes when the start point was clipped - paired with the potential of the nodes steps, steps - 1 (:285-296); every later node i is
// at position h i.
__device__ __forceinline__ double uniform_start_point(const GridDev& g, double kappa) { return fmin(g.max_r, 200. / kappa); }

__device__ __forceinline__ int start_index(const GridDev& g, double kappa)
{
    if (g.uniform) return (int)(uniform_start_point(g, kappa) / g.h);
    // r_i = Rp (e^{δ i} - 1) evaluated in place of the table: 14 dependent table loads would cost ~5000 cycles per call
    // (L2 latency) on kernels whose whole round is ~50000; a last-ulp difference can move the index by one node, where the
    // solution is 1e-200
    int hi = g.N - 1, lo = 1;
    const double hd = 0.5 * g.delta;
    while (hi - lo > 1) {
        const int mid = (hi + lo) >> 1;
        const double arg = -(g.rp * expm1(g.delta * (double)mid)) * kappa - (double)mid * hd;
        if (arg < kFarLog) hi = mid; else lo = mid;
    }
    return hi;
}

// far cut-off index, Numerov.h:119-136: the result of start_index() above - the smallest idx in [2, N-1] whose far value
// is below 1e-200, N-1 if there is none - found from the closed-form estimate and confirmed with the same predicate (2-4 evaluations
// instead of 14 dependent ones); anything unexpected falls back to the bisection
__device__ __forceinline__ bool far_below(const GridDev& g, double kappa, int idx)
{
    return -(g.rp * expm1(g.delta * (double)idx)) * kappa - (double)idx * (0.5 * g.delta) < kFarLog;
}
__device__ __forceinline__ int start_index_fast(const GridDev& g, double kappa)
{
    if (g.uniform) return start_index(g, kappa);
    const int nmax = g.N - 1;
    // r* kappa + idx delta/2 = 460.5 with r = Rp (e^{delta idx} - 1): two fixed-point steps on idx
    double x = (double)nmax;
#pragma unroll
    for (int it = 0; it < 2; ++it) {
        const double rr = (-kFarLog - x * (0.5 * g.delta)) / kappa;
        x = rr > 0. ? log1p(rr / g.rp) / g.delta : 1.;
        x = fmin(fmax(x, 1.), (double)nmax);
    }
    int idx = min(max((int)x + 1, 2), nmax);
    // walk to the boundary: want far_below(idx) && !far_below(idx - 1)   (idx == nmax is never tested by the bisection: it is its initial hi)
    for (int it = 0; it < 6; ++it) {
        const bool b1 = idx >= nmax ? true : far_below(g, kappa, idx);
        if (!b1) { ++idx; continue; }
        const bool b0 = idx <= 2 ? false : far_below(g, kappa, idx - 1);
        if (b0) { --idx; continue; }
        return idx;
    }
    return start_index(g, kappa);
}

// far boundary value of seed node idx (= start or start - 1)
__device__ __forceinline__ double far_value(const GridDev& g, double kappa, int idx, int start)
{
    if (g.uniform) return exp(-(uniform_start_point(g, kappa) - (double)(start - idx) * g.h) * kappa);      // Numerov.h:33-36
    return exp(-__ldg(g.r + idx) * kappa - (double)idx * (0.5 * g.delta));                                   // Numerov.h:103-108
}
// what a seed node's f/12 differs by from the table value: on the uniform grid its centrifugal term is taken at the seed's own
// position, not at the node's (0 on the logarithmic grid, for l = 0, and when the start point is a node)
__device__ __forceinline__ double seed_g_shift(const GridDev& g, double ll1, double kappa, int idx, int start)
{
    if (!g.uniform || ll1 == 0.) return 0.;
    const double pos = uniform_start_point(g, kappa) - (double)(start - idx) * g.h;
    const double ri = g.h * (double)idx;
    return ll1 * (g.h * g.h * (1. / 12.)) * (1. / (pos * pos) - 1. / (ri * ri));
}

// The two-sided matched solution (Numerov.h:403-504) on either grid: far seeds, near-nucleus start value, and - uniform grid only - the
// reference's re-computation of the step from the truncated range, h' = startPoint / steps (:430-432): every position becomes h' i,
// i.e. the potential and energy parts of f_i h'^2 / 12 carry the factor (h'/h)^2 while the centrifugal part l(l+1) / (12 i^2) does not.
struct MatchScale { double rho2, y_s0, y_s1, y1; };
__device__ __forceinline__ MatchScale match_scale(const GridDev& g, double kappa, int start, int l)
{
    MatchScale m;
    if (g.uniform) {
        const double sp = uniform_start_point(g, kappa);
        const double hp = sp / (double)start;
        m.rho2 = (hp / g.h) * (hp / g.h);
        m.y_s0 = exp(-sp * kappa);                                   // GetBoundaryValueFar at startPoint, startPoint - h'  (:434-446)
        m.y_s1 = exp(-(sp - hp) * kappa);
        m.y1 = pow(hp, (double)l + 1.);                              // GetBoundaryValueZero at h'  (Numerov.h:38-41, :470-477)
    } else {
        m.rho2 = 1.;
        m.y_s0 = far_value(g, kappa, start, start);
        m.y_s1 = far_value(g, kappa, start - 1, start);
        m.y1 = pow(__ldg(g.r + 1), (double)l + 1.) * exp(-0.5 * g.delta);                                   // Numerov.h:110-116
    }
    return m;
}
__device__ __forceinline__ double match_g(const GridDev& g, const double* __restrict__ atab, double ll1, double E, double rho2, int i)
{   // f_i / 12 of the matched solve
    const double a = __ldg(atab + i), b = __ldg(g.b12 + i), c = __ldg(g.c6 + i);
    return g.uniform ? fma(rho2, fma(-E, c, a), ll1 * b) : fma(-E, c, fma(ll1, b, a));
}

struct LaneOut { int count; int count_full; int y0_pos; int seen; int steps; double y0_log2; double d_first; };

// One inward sweep for one (l, E) lane; all 32 lanes of the warp walk the same node index so table loads
// are warp-uniform broadcasts.  Semantics of `count`: SolveSchrodingerCountNodes incl. its early exits;
// (y0_pos, y0_log2): sign and log2|.| of SolveSchrodingerSolutionInZero.
static __device__ __noinline__ LaneOut sweep_lane(const GridDev& g, const double* __restrict__ atab, int l, double E, int limit)
{
    const double kappa = sqrt(2. * fabs(E));
    const int start = start_index(g, kappa);
    int imax = start;
#pragma unroll
    for (int o = 16; o; o >>= 1) imax = max(imax, __shfl_xor_sync(0xffffffffu, imax, o));

    const double ll1 = (double)(l * (l + 1));
    const double thr = 1. - g.delta * g.delta * (1. / 48.);   // d_i >= thr  <=>  Veff_i <= E  (Numerov.h:336-337; uniform grid: delta = 0)

    double W1 = 0., W2 = 0.;      // W_{i+1}, W_{i+2}
    double d1 = 1., d2 = 1.;      // d_{i+1}, d_{i+2}
    double P = 1.;                // prod_{j=i+1..start} d_j  (magnitude and sign)
    int count = 0, snap = -1, cfull = 0;
    unsigned old_sign = 0;        // sign bit of y at the previous node (far values are positive)
    unsigned full_sign = 0;
    bool seen = false;

    for (int i = imax; i >= 1; --i) {
        const double a = __ldg(atab + i), b = __ldg(g.b12 + i), c = __ldg(g.c6 + i);
        double gq = fma(-E, c, fma(ll1, b, a));                // f_i / 12
        if (i >= start - 1 && i <= start) gq += seed_g_shift(g, ll1, kappa, i, start);
        const double d = 1. - gq;
        if (i > start) continue;
        P *= d1;                                       // P = P_i = prod_{j>i} d_j  (d1 = 1 at i = start)
        double W;
        if (i <= start - 2) {
            const double n1 = fma(-10., d1, 12.);
            W = fma(n1, W1, -(d2 * d1) * W2);
            // y_i = W_i / (P_i d_i)
            const unsigned sy = ((unsigned)hi32(W) ^ (unsigned)hi32(P) ^ (unsigned)hi32(d)) >> 31;
            if (sy != full_sign) { ++cfull; full_sign = sy; }       // every sign change down to i = 1 (Sturm count)
            if (snap < 0) {
                if (sy != old_sign) { ++count; old_sign = sy; }
                if (d >= thr) seen = true;
                else if (seen) snap = count;          // inner turning point: Numerov.h:339-340
            }
        } else if (i == start) {
            W = d * far_value(g, kappa, i, start);    // w_start (P_start = 1)
        } else {
            W = d * far_value(g, kappa, i, start) * d1;      // w_{start-1} d_start
        }
        W2 = W1; W1 = W; d2 = d1; d1 = d;
    }
    // here W1 = W_1, W2 = W_2, d1 = d_1, d2 = d_2, P = P_1 = prod_{j>=2} d_j
    // y_0 = y_1 (2 + f_1) - y_2  (Numerov.h:398) ;  y_1 = W_1/(P_1 d_1), y_2 = W_2/P_1, f_1 = 12 (1 - d_1)
    const double Y0s = W1 * fma(-12., d1, 14.) / d1 - W2;
    LaneOut o;
    o.y0_pos = (Y0s > 0.) != (P < 0.);
    o.y0_log2 = log2(fabs(Y0s)) - log2(fabs(P));
    if (!(fabs(Y0s) <= 1.7e308)) o.y0_log2 = INFINITY;     // NaN or Inf
    if (snap >= 0) count = snap;
    else if (count <= limit) {                              // Numerov.h:343-348
        const unsigned s0 = o.y0_pos ? 0u : 1u;
        if (s0 != old_sign) ++count;
    }
    o.count = min(count, limit + 1);                        // the reference returns as soon as count > limit
    o.count_full = cfull + (((o.y0_pos ? 0u : 1u) != full_sign) ? 1 : 0);
    o.d_first = d1;
    o.seen = seen;
    o.steps = start - 1;                                    // algorithmic node-steps of this sweep
    return o;
}


// "virtual bisection" over K sampled points: emulates what a bisection restricted to the sampled set would do,
// so that a non-monotone predicate is resolved the way the reference's bisection resolves it.
// hi_mask bit j = 1 means "point j belongs to the HIGH side" (move hi down to it).
__device__ __forceinline__ void virtual_bisect(unsigned hi_mask, int K, int& lo_idx, int& hi_idx, int& last_mid)
{
    lo_idx = -1; hi_idx = K; last_mid = -1;
    while (hi_idx - lo_idx > 1) {
        const int mid = (lo_idx + hi_idx) >> 1;
        last_mid = mid;
        if ((hi_mask >> mid) & 1u) hi_idx = mid; else lo_idx = mid;
    }
}

__device__ __forceinline__ bool bracket_open(double lo, double hi)
{   // the reference loops while (toe - boe > energyErr); also stop when the bracket cannot shrink any more
    return (hi - lo > kEnergyTol) && (0.5 * (lo + hi) != lo) && (0.5 * (lo + hi) != hi);
}

// ---------------------------------------------------------------------------------------------------------
// Energy search of one level: sampling of the trial energies and bracket update, shared by the search kernels.
// One round = 32 trial energies (one per lane, ascending in energy) + the monotone predicate
//     Q(E) = [full sign-change count of the inward solution > wanted node count]
// evaluated on each.  The predicate is monotone, so ANY ascending set of trial energies brackets the same root; what
// the set looks like only decides how fast the bracket shrinks.  Two shapes are used:
//   uniform : 32 points that cut [lo, hi] into 33 equal parts (cold start, and whenever nothing better is known);
//   ladder  : a two-sided geometric ladder  c -+ eps g^m  (m = 0 .. 15, outermost offset = R) around an estimate c of
//             the root.  Round 0 of SCF step >= 1 centres it on the previous step's eigenvalue; later rounds centre
//             it on the zero of y0(E) interpolated (inverse cubic Lagrange) through the samples next to the sign
//             change, with R = 4 |cubic - secant| as the trust radius.  A wrong estimate only leaves a wide bracket
//             (next round: uniform); a good one closes the bracket to 1e-12 in 2-3 rounds instead of ~6 (33-section)
//             or ~140 serial sweeps (the reference's three bisections, DFTAtom.cpp:513-604).
// ---------------------------------------------------------------------------------------------------------
struct Bracket {
    double lo, hi;          // the root is in (lo, hi]
    double c_est, radius;   // ladder centre and outermost offset
    double ylog;            // log2 |y0| of the last virtual-bisection midpoint (1e15 guard, DFTAtom.cpp:528)
    bool ladder;
    // after a round whose 32 samples all fell on one side of the root (a ladder that missed): the next uniform round covers
    // only a window of this width next to the samples (32 x their span), not the whole remaining bracket, which after a
    // warm start still reaches to -Z^2 - 1 or +50 and would cost ~6 more 33-section rounds.  side: -1 below hi, +1 above lo
    double win = 0.;
    int win_side = 0;
};
constexpr double kLadderEps = 2.4e-13;

__device__ __forceinline__ double sample_energy(const Bracket& b, int lane)
{
    if (b.ladder) {
        const double lg = log2(fmax(b.radius, 2. * kLadderEps) / kLadderEps) * (1. / 15.);
        const int mstep = (lane < 16) ? (15 - lane) : (lane - 16);            // 0 = closest to the estimate
        const double off = kLadderEps * exp2((double)mstep * lg);
        return fmin(fmax((lane < 16) ? b.c_est - off : b.c_est + off, b.lo), b.hi);
    }
    double lo = b.lo, hi = b.hi;
    if (b.win_side < 0) lo = fmax(lo, hi - b.win); else if (b.win_side > 0) hi = fmin(hi, lo + b.win);
    return lo + (hi - lo) * ((double)(lane + 1) * (1. / 33.));
}

// warp-collective: every lane passes its own trial energy and results; all lanes end with the same bracket
__device__ __forceinline__ void update_bracket(Bracket& b, double E, bool high, int y0_pos, double y0_log2)
{
    const unsigned full = 0xffffffffu;
    const unsigned m_hi = __ballot_sync(full, high);
    int lo_i, hi_i, lm;
    virtual_bisect(m_hi, 32, lo_i, hi_i, lm);
    const double a_lo = __shfl_sync(full, E, max(lo_i, 0)), a_hi = __shfl_sync(full, E, min(hi_i, 31));
    const double e_lo = lo_i >= 0 ? a_lo : b.lo, e_hi = hi_i < 32 ? a_hi : b.hi;
    b.ylog = __shfl_sync(full, y0_log2, lm);
    b.ladder = false;
    {
        const double span = __shfl_sync(full, E, 31) - __shfl_sync(full, E, 0);
        b.win_side = (lo_i < 0 && hi_i < 32) ? -1 : ((hi_i == 32 && lo_i >= 0) ? 1 : 0);
        b.win = 32. * fmax(span, 16. * kLadderEps);
    }
    // estimate of the root for the next round: zero of y0(E) through the samples around the sign change
    if (lo_i >= 0 && hi_i < 32 && e_lo < e_hi) {
        double Ek[4], yk[4], lgv[4];
        bool ok[4];
        double ref = -INFINITY;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int idx = lo_i - 1 + q;
            const int src_lane = min(max(idx, 0), 31);
            Ek[q] = __shfl_sync(full, E, src_lane);
            lgv[q] = __shfl_sync(full, y0_log2, src_lane);
            yk[q] = __shfl_sync(full, y0_pos, src_lane) ? 1. : -1.;
            ok[q] = idx >= 0 && idx < 32 && lgv[q] > -1e300 && lgv[q] < 1e300;
            if (ok[q]) ref = fmax(ref, lgv[q]);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) yk[q] = ok[q] ? yk[q] * exp2(lgv[q] - ref) : 0.;      // relative to the largest sample
        // the bracket ends must be proper samples with opposite signs and distinct energies
        if (ok[1] && ok[2] && yk[1] * yk[2] < 0. && Ek[1] < Ek[2]) {
            const double E2 = Ek[1] - yk[1] * (Ek[2] - Ek[1]) / (yk[2] - yk[1]);          // secant
            // outer points are usable when they extend the table monotonically in E and in y (inverse interpolation)
            const bool use0 = ok[0] && Ek[0] < Ek[1] && (yk[0] - yk[1]) * (yk[1] - yk[2]) > 0.;
            const bool use3 = ok[3] && Ek[3] > Ek[2] && (yk[2] - yk[3]) * (yk[1] - yk[2]) > 0.;
            double Eh = E2;
            if (use0 || use3) {
                double num = 0.;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const bool uq = (q == 0) ? use0 : (q == 3 ? use3 : true);
                    if (!uq) continue;
                    double wgt = Ek[q];
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        const bool ur = (r == 0) ? use0 : (r == 3 ? use3 : true);
                        if (r == q || !ur) continue;
                        wgt *= (0. - yk[r]) / (yk[q] - yk[r]);
                    }
                    num += wgt;
                }
                Eh = num;
            }
            if (!(Eh > e_lo && Eh < e_hi)) Eh = E2;
            if (Eh > e_lo && Eh < e_hi) {
                b.c_est = Eh;
                const double trust = (use0 || use3) ? 4. * fabs(Eh - E2) : 0.25 * (e_hi - e_lo);
                b.radius = fmin(fmax(trust, 16. * kLadderEps), fmax(e_hi - Eh, Eh - e_lo));
                b.ladder = true;
            }
        }
    }
    b.lo = e_lo; b.hi = e_hi;
}

}  // namespace dft
