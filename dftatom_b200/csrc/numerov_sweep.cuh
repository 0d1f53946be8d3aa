// The tile-staged, division-free inward Numerov sweep shared by the search kernels (numerov_fast.cu: one warp sweeps the
// whole grid; numerov_seg.cu: one warp sweeps one radial segment).  Replaces the node loops of Numerov.h:272-401.
//
// One warp, its 32 lanes = 32 trial energies walking the same node index; every lane carries EPL independent chains.
// The per-node tables (ab_i = a_i + l(l+1) b_i, c_i) are staged through shared memory in tiles of 32 nodes (coalesced
// global loads one tile ahead, broadcast LDS.128 in the loop).
//
// Difference form.  With g_i = f_i/12 (small: ~1e-9..1e-6 on the fine grids), d_i = 1 - g_i and
//   s_i = 1 - d_i d_{i+1} = g_i + g_{i+1} - g_i g_{i+1},
// the scaled recurrence W_{i-1} = (12 - 10 d_i) W_i - d_i d_{i+1} W_{i+1}  (W_i = w_i prod_{j>i} d_j) is evaluated as
//   D_i = D_{i+1} + 10 g_i W_i + s_i W_{i+1},   W_{i-1} = W_i + D_i        (D_i = W_{i-1} - W_i).
// Forming 12 - 10 d_i or d_i d_{i+1} as numbers near 1..2 would round the physics (g ~ 1e-8) to 1e-16 absolute, i.e.
// perturb the local potential by ~1e-8 relative at every node - measured as 2e-6 Ha on the Rn 1s level at 131073
// nodes; in the difference form every coefficient keeps full relative precision, like the reference's
// w_next = 2w - w_prev + y f (Numerov.h:311).  Sign bits of y_i = W_i / (P_i d_i) are shifted into a register (one SHF
// per node) and popcounted per tile: the count of ALL sign changes is the Sturm count of the recurrence.
#pragma once
#include "numerov_common.cuh"

namespace dft {

template <int EPL> struct FastOut {
    int cfull[EPL]; int y0_pos[EPL]; double y0_log2[EPL]; double d_first[EPL];
    int bad; long long steps;
    // range sweeps: state after the lowest node `bot` of the range, in the form the next range enters with:
    // (W_bot, D = W_bot - W_{bot+1}); count = sign changes inside the range; prev = sign of y_bot; P = product of
    // d_i d_{i+1} over the even nodes of the range; Y0s = y_0 scaled by P (only meaningful when bot == 1)
    double W[EPL], D[EPL], P[EPL], Y0s[EPL];
    int count[EPL]; unsigned prev[EPL];
};

// Chain e of a lane either starts inside the range at its own far seeds (Numerov.h:294-303; start[e] = seed index, or
// < bot: the chain never starts) or is already running when the range begins (running[e]: entry state (W_in, D_in) =
// (W_{top+1}, W_{top+1} - W_{top+2}); its start is then treated as above the range).
template <int EPL> struct SweepIn {
    double E[EPL];
    int start[EPL];
    bool running[EPL];
    double W_in[EPL], D_in[EPL];
};

// warp-collective; sbuf = this warp's double-buffered tile staging area [2][32]; nodes top .. bot (descending), bot >= 1
template <int EPL>
__device__ __forceinline__ void range_sweep(const GridDev& g, const double* __restrict__ atab, double ll1, const SweepIn<EPL>& in,
                                            int top, int bot, double2* sbuf, FastOut<EPL>& out)
{
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int nmax = g.N - 1;
    double E[EPL], kappa[EPL];
    int start[EPL];
#pragma unroll
    for (int e = 0; e < EPL; ++e) {
        E[e] = in.E[e];
        kappa[e] = sqrt(2. * fabs(E[e]));
        start[e] = in.running[e] ? 0x3fffffff : in.start[e];
    }
    // W1 = W_{i+1}, W2 = W_{i+2}, D = W_{i+1} - W_{i+2}, g1 = g_{i+1}, s1 = s_{i+1}, t1 = 10 g_{i+1}, P = prod d
    double W1[EPL], W2[EPL], D[EPL], g1[EPL], s1[EPL], t1[EPL], P[EPL];
    unsigned prev[EPL];
    int count[EPL], bad = 0;
#pragma unroll
    for (int e = 0; e < EPL; ++e) {
        W1[e] = 0.; W2[e] = 0.; D[e] = 0.; g1[e] = 0.; s1[e] = 0.; t1[e] = 0.; P[e] = 1.; prev[e] = 0; count[e] = 0;
        if (in.running[e]) {
            const int i1 = min(top + 1, nmax), i2 = min(top + 2, nmax);
            const double ga = fma(-E[e], __ldg(g.c6 + i1), fma(ll1, __ldg(g.b12 + i1), __ldg(atab + i1)));
            const double gb = fma(-E[e], __ldg(g.c6 + i2), fma(ll1, __ldg(g.b12 + i2), __ldg(atab + i2)));
            W1[e] = in.W_in[e]; D[e] = in.D_in[e]; W2[e] = in.W_in[e] - in.D_in[e];
            g1[e] = ga; s1[e] = fma(-ga, gb, ga + gb); t1[e] = 10. * ga;
            prev[e] = ((unsigned)hi32(W1[e]) ^ (unsigned)hi32(1. - ga)) >> 31;
        } else if (in.start[e] == top + 1 && top + 1 <= nmax) {
            // the far seed w_start lies just above the range (it depends on the tables only): the range begins with the
            // second seed w_{start-1}
            const int i1 = top + 1;
            const double ga = fma(-E[e], __ldg(g.c6 + i1), fma(ll1, __ldg(g.b12 + i1), __ldg(atab + i1))) + seed_g_shift(g, ll1, kappa[e], i1, in.start[e]);
            W1[e] = (1. - ga) * far_value(g, kappa[e], i1, in.start[e]);
            g1[e] = ga;
            if (!(i1 & 1)) P[e] = 1. - ga;                 // d_{start+1} := 1
            bad |= !(1. - ga > 0.);
        }
    }

    int m = top >> 5;
    const int m_last = bot >> 5;
    // prefetch the top tile: lane j holds node 32 m + 31 - j
    double pa, pb, pc;
    {
        const int i = min((m << 5) + 31 - lane, nmax);
        pa = __ldg(atab + i); pb = __ldg(g.b12 + i); pc = __ldg(g.c6 + i);
    }
    int cur = 0;
    for (; m >= m_last; --m) {
        sbuf[cur * 32 + lane] = make_double2(fma(ll1, pb, pa), pc);      // (g_i at E = 0, dg_i/d(-E))
        __syncwarp();
        if (m > m_last) {
            const int i = ((m - 1) << 5) + 31 - lane;
            pa = __ldg(atab + i); pb = __ldg(g.b12 + i); pc = __ldg(g.c6 + i);
        }
        const int hi_i = (m << 5) + 31, lo_i = m << 5;
        bool uniform = true;
#pragma unroll
        for (int e = 0; e < EPL; ++e) uniform = uniform && ((start[e] >= hi_i + 2) || (start[e] < lo_i));
        const double2* tile = sbuf + cur * 32;
        if (lo_i >= bot && hi_i <= top && lo_i >= 1 && __all_sync(full, uniform)) {
            // ---- fast tile: every chain is either fully inside its sweep or has not started yet (W stays 0) ----
            // Four quarters of 8 nodes.  Phase A (no loop-carried dependence): g, s, 10 g of the quarter.
            // Phase B: the (D, W) chain, two dependent FP64 operations per node.
            unsigned sb[EPL];
#pragma unroll
            for (int e = 0; e < EPL; ++e) sb[e] = 0;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                double sq[EPL][8], tq[EPL][8];
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const double2 t = tile[q * 8 + k];
#pragma unroll
                    for (int e = 0; e < EPL; ++e) {
                        const double gk = fma(-E[e], t.y, t.x);
                        sq[e][k] = fma(-gk, g1[e], gk + g1[e]);
                        tq[e][k] = 10. * gk;
                        g1[e] = gk;
                    }
                }
#pragma unroll
                for (int e = 0; e < EPL; ++e)      // even node index <=> odd k: d_i d_{i+1} = 1 - s_i of the pairs (i, i+1)
                    P[e] *= ((1. - sq[e][1]) * (1. - sq[e][3])) * ((1. - sq[e][5]) * (1. - sq[e][7]));
#pragma unroll
                for (int k = 0; k < 8; ++k) {
#pragma unroll
                    for (int e = 0; e < EPL; ++e) {
                        const double tu = k ? tq[e][k - 1] : t1[e], su = k ? sq[e][k - 1] : s1[e];
                        const double Dn = fma(tu, W1[e], fma(su, W2[e], D[e]));
                        const double W = W1[e] + Dn;
                        sb[e] = __funnelshift_l((unsigned)hi32(W), sb[e], 1);
                        W2[e] = W1[e]; W1[e] = W; D[e] = Dn;
                    }
                }
#pragma unroll
                for (int e = 0; e < EPL; ++e) { t1[e] = tq[e][7]; s1[e] = sq[e][7]; }
            }
#pragma unroll
            for (int e = 0; e < EPL; ++e) {
                const unsigned x = sb[e] ^ ((sb[e] >> 1) | (prev[e] << 31));
                count[e] += __popc(x);
                prev[e] = sb[e] & 1u;
            }
        } else {
            // ---- general tile: seeds (far boundary values), partial tiles at the ends of the range, sign of d ----
            for (int k = 0; k < 32; ++k) {
                const int i = hi_i - k;
                if (i < bot) break;
                if (i > top) continue;
                const double2 t = tile[k];
#pragma unroll
                for (int e = 0; e < EPL; ++e) {
                    double gk = fma(-E[e], t.y, t.x);
                    if (g.uniform && i >= start[e] - 1 && i <= start[e]) gk += seed_g_shift(g, ll1, kappa[e], i, start[e]);
                    const double d = 1. - gk;
                    if (i <= start[e]) {
                        double W, s, Dnew;
                        if (i == start[e]) {                      // w_start = d_start far(start)   (Numerov.h:294-298)
                            W = d * far_value(g, kappa[e], i, start[e]);
                            s = gk;                               // d_{start+1} := 1
                            P[e] = 1.; count[e] = 0; prev[e] = 0;
                            Dnew = 0.;                            // overwritten at the next node
                            bad |= !(d > 0.);
                        } else if (i == start[e] - 1) {           // w_{start-1} d_start            (Numerov.h:300-303)
                            W = d * far_value(g, kappa[e], i, start[e]) * (1. - g1[e]);
                            s = fma(-gk, g1[e], gk + g1[e]);
                            Dnew = W - W1[e];                     // D_start = W_{start-1} - W_start
                            bad |= !(d > 0.);
                        } else {
                            Dnew = fma(t1[e], W1[e], fma(s1[e], W2[e], D[e]));
                            W = W1[e] + Dnew;
                            s = fma(-gk, g1[e], gk + g1[e]);
                            const unsigned sy = ((unsigned)hi32(W) ^ (unsigned)hi32(d)) >> 31;    // y_i = W_i / (P_i d_i), P_i > 0
                            count[e] += (sy != prev[e]);
                            prev[e] = sy;
                            if (i == 2) bad |= !(d > 0.);
                        }
                        if (!(i & 1)) P[e] *= (1. - s);
                        D[e] = Dnew;
                        W2[e] = W1[e]; W1[e] = W; g1[e] = gk; s1[e] = s; t1[e] = 10. * gk;
                    }
                }
            }
        }
        cur ^= 1;
    }
    out.bad = bad;
    out.steps = 0;
#pragma unroll
    for (int e = 0; e < EPL; ++e) {
        // bot == 1: W1 = W_1, W2 = W_2, g1 = g_1, P = prod_{j=2..start} d_j;  y_0 = y_1 (2 + f_1) - y_2  (Numerov.h:398)
        const double d1 = 1. - g1[e];
        const double Y0s = W1[e] * fma(12., g1[e], 2.) / d1 - W2[e];
        out.Y0s[e] = Y0s;
        out.y0_pos[e] = Y0s > 0.;
        out.y0_log2[e] = (fabs(Y0s) <= 1.7e308) ? log2(fabs(Y0s)) - log2(fabs(P[e])) : INFINITY;
        out.cfull[e] = count[e] + (((out.y0_pos[e] ? 0u : 1u) != prev[e]) ? 1 : 0);
        out.bad |= !(P[e] > 0.);
        out.steps += in.start[e] - 1;
        out.d_first[e] = d1;
        out.W[e] = W1[e]; out.D[e] = D[e]; out.P[e] = P[e]; out.count[e] = count[e]; out.prev[e] = prev[e];
    }
}

// the whole inward sweep of EPL trial energies per lane: nodes start .. 1
template <int EPL>
__device__ __forceinline__ void fast_sweep(const GridDev& g, const double* __restrict__ atab, double ll1, const double (&E)[EPL],
                                           double2* sbuf, FastOut<EPL>& out)
{
    const unsigned full = 0xffffffffu;
    SweepIn<EPL> in;
    int imax = 0;
#pragma unroll
    for (int e = 0; e < EPL; ++e) {
        in.E[e] = E[e];
        in.start[e] = start_index(g, sqrt(2. * fabs(E[e])));
        in.running[e] = false; in.W_in[e] = 0.; in.D_in[e] = 0.;
        imax = max(imax, in.start[e]);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) imax = max(imax, __shfl_xor_sync(full, imax, o));
    range_sweep<EPL>(g, atab, ll1, in, imax, 1, sbuf, out);
}

}  // namespace dft
