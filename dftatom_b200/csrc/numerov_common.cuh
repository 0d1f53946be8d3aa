// Shared device helpers of the Numerov kernels (included by numerov.cu and numerov_fast.cu).
#pragma once
#include "internal.h"
#include <cmath>

namespace dft {

__device__ __forceinline__ int hi32(double x) { return __double2hiint(x); }

// Numerov.h:119-136: bisection on the index for far(idx) = exp(-r_idx sqrt(2|E|) - idx δ/2) < 1e-200
__device__ __forceinline__ int start_index(const GridDev& g, double kappa)
{
    int hi = g.N - 1, lo = 1;
    const double hd = 0.5 * g.delta;
    while (hi - lo > 1) {
        const int mid = (hi + lo) >> 1;
        const double arg = -__ldg(g.r + mid) * kappa - (double)mid * hd;
        if (arg < kFarLog) hi = mid; else lo = mid;
    }
    return hi;
}

__device__ __forceinline__ double far_value(const GridDev& g, double kappa, int idx)
{   // Numerov.h:103-108
    return exp(-__ldg(g.r + idx) * kappa - (double)idx * (0.5 * g.delta));
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
    const double thr = 1. - g.delta * g.delta * (1. / 48.);   // d_i >= thr  <=>  Veff_i <= E  (Numerov.h:336-337)

    double W1 = 0., W2 = 0.;      // W_{i+1}, W_{i+2}
    double d1 = 1., d2 = 1.;      // d_{i+1}, d_{i+2}
    double P = 1.;                // prod_{j=i+1..start} d_j  (magnitude and sign)
    int count = 0, snap = -1, cfull = 0;
    unsigned old_sign = 0;        // sign bit of y at the previous node (far values are positive)
    unsigned full_sign = 0;
    bool seen = false;

    for (int i = imax; i >= 1; --i) {
        const double a = __ldg(atab + i), b = __ldg(g.b12 + i), c = __ldg(g.c6 + i);
        const double gq = fma(-E, c, fma(ll1, b, a));          // f_i / 12
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
            W = d * far_value(g, kappa, i);           // w_start (P_start = 1)
        } else {
            W = d * far_value(g, kappa, i) * d1;      // w_{start-1} d_start
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

}  // namespace dft
