// Fast Numerov shooting: tile-staged, division-free inward sweeps and the fused multisection eigenvalue search.
//
// Replaces the per-level energy search of reference DFTAtom/DFTAtom.cpp:493-541 (LoopOverLevels) + :566-604
// (LocateInterval) and the sweeps it calls (Numerov.h:272-401).
//
// Search predicate.  The reference brackets a level in three bisections (upper edge of the node-count window,
// lower edge, then the sign change of y(0) inside the window).  Its node count stops at the inner classical
// turning point; the count of ALL sign changes of y_start..y_1,y_0 (no early exit) is the Sturm count of the
// three-term recurrence: it is monotone in E and steps exactly where the reference's y(0) changes sign.  A node
// can hide inside the inner forbidden region at most once, so inside the reference's window that full count takes
// the values {want, want+1} only: the eigenvalue the reference returns is the single step  full_count: want ->
// want+1  (plus a constant 1 for l = 3, where 1 - f_1/12 < 0 flips the sign of y_1 at every energy).
// One K-section search on  Q(E) = [full_count(E) > want + off]  therefore lands on the same eigenvalue (checked
// against the reference on every level of every golden atom, tests/test_gpu_scf.py) in ~11 rounds of 32 trial
// energies instead of ~140 serial sweeps.
//
// Sweep.  One warp = one orbital, its 32 lanes = 32 trial energies walking the same node index.  The per-node
// tables (ab_i = a_i - l(l+1) b_i, c_i) are staged through shared memory in tiles of 32 nodes (coalesced global
// loads one tile ahead, broadcast LDS.128 in the loop); d_i = ab_i + E c_i; the recurrence is the division-free
//     W_{i-1} = (12 - 10 d_i) W_i - d_i d_{i+1} W_{i+1}     (5.5 FP64 instructions per lane and node).
// Sign bits are shifted into a register (one SHF per node) and popcounted per tile.
#include "numerov_common.cuh"

namespace dft {

// ld.shared.v2.f64 as volatile asm: keeps the staged-table loads where they are written (ptxas otherwise sinks every
// LDS next to its first use and recycles the same destination registers, which serialises the loads and puts the
// ~30-cycle shared-memory latency on every node).
__device__ __forceinline__ double2 lds_f64x2(unsigned addr)
{
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
    return v;
}

template <int EPL> struct FastOut {
    int cfull[EPL]; int y0_pos[EPL]; double y0_log2[EPL]; double d_first[EPL];
    int bad; long long steps;
};

// warp-collective; sbuf = this warp's double-buffered tile staging area [2][32].
// Every lane carries EPL independent trial energies (EPL chains of the recurrence interleave in the pipeline).
//
// Difference form.  With g_i = f_i/12 (small: ~1e-9..1e-6 on the fine grids), d_i = 1 - g_i and
//   s_i = 1 - d_i d_{i+1} = g_i + g_{i+1} - g_i g_{i+1},
// the scaled recurrence W_{i-1} = (12 - 10 d_i) W_i - d_i d_{i+1} W_{i+1} is evaluated as
//   D_i = D_{i+1} + 10 g_i W_i + s_i W_{i+1},   W_{i-1} = W_i + D_i        (D_i = W_{i-1} - W_i).
// Forming 12 - 10 d_i or d_i d_{i+1} as numbers near 1..2 would round the physics (g ~ 1e-8) to 1e-16 absolute, i.e.
// perturb the local potential by ~1e-8 relative at every node - measured as 2e-6 Ha on the Rn 1s level at 131073
// nodes; in the difference form every coefficient keeps full relative precision, like the reference's
// w_next = 2w - w_prev + y f (Numerov.h:311).
template <int EPL>
__device__ __forceinline__ void fast_sweep(const GridDev& g, const double* __restrict__ atab, double ll1, const double (&E)[EPL],
                                           double2* sbuf, FastOut<EPL>& out)
{
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    double kappa[EPL];
    int start[EPL];
    int imax = 0;
#pragma unroll
    for (int e = 0; e < EPL; ++e) {
        kappa[e] = sqrt(2. * fabs(E[e]));
        start[e] = start_index(g, kappa[e]);
        imax = max(imax, start[e]);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) imax = max(imax, __shfl_xor_sync(full, imax, o));
    const int nmax = g.N - 1;

    // W1 = W_{i+1}, W2 = W_{i+2}, D = W_{i+1} - W_{i+2}, g1 = g_{i+1}, s1 = s_{i+1}, t1 = 10 g_{i+1}, P = prod d
    double W1[EPL], W2[EPL], D[EPL], g1[EPL], s1[EPL], t1[EPL], P[EPL];
    unsigned prev[EPL];
    int count[EPL], bad = 0;
#pragma unroll
    for (int e = 0; e < EPL; ++e) { W1[e] = 0.; W2[e] = 0.; D[e] = 0.; g1[e] = 0.; s1[e] = 0.; t1[e] = 0.; P[e] = 1.; prev[e] = 0; count[e] = 0; }

    int m = imax >> 5;
    // prefetch the top tile: lane j holds node 32 m + 31 - j
    double pa, pb, pc;
    {
        const int i = min((m << 5) + 31 - lane, nmax);
        pa = __ldg(atab + i); pb = __ldg(g.b12 + i); pc = __ldg(g.c6 + i);
    }
    int cur = 0;
    for (; m >= 0; --m) {
        sbuf[cur * 32 + lane] = make_double2(fma(ll1, pb, pa), pc);      // (g_i at E = 0, dg_i/d(-E))
        __syncwarp();
        if (m > 0) {
            const int i = ((m - 1) << 5) + 31 - lane;
            pa = __ldg(atab + i); pb = __ldg(g.b12 + i); pc = __ldg(g.c6 + i);
        }
        const int hi_i = (m << 5) + 31, lo_i = m << 5;
        bool uniform = true;
#pragma unroll
        for (int e = 0; e < EPL; ++e) uniform = uniform && ((start[e] >= hi_i + 2) || (start[e] < lo_i));
        const double2* tile = sbuf + cur * 32;
        if (m > 0 && __all_sync(full, uniform)) {
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
            // ---- general tile: seeds (far boundary values), the last tile down to i = 1, sign of d ----
            for (int k = 0; k < 32; ++k) {
                const int i = hi_i - k;
                if (i < 1) break;
                const double2 t = tile[k];
#pragma unroll
                for (int e = 0; e < EPL; ++e) {
                    const double gk = fma(-E[e], t.y, t.x);
                    const double d = 1. - gk;
                    if (i <= start[e]) {
                        double W, s, Dnew;
                        if (i == start[e]) {                      // w_start = d_start far(start)   (Numerov.h:294-298)
                            W = d * far_value(g, kappa[e], i);
                            s = gk;                               // d_{start+1} := 1
                            P[e] = 1.; count[e] = 0; prev[e] = 0;
                            Dnew = 0.;                            // overwritten at the next node
                            bad |= !(d > 0.);
                        } else if (i == start[e] - 1) {           // w_{start-1} d_start            (Numerov.h:300-303)
                            W = d * far_value(g, kappa[e], i) * (1. - g1[e]);
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
        // W1 = W_1, W2 = W_2, g1 = g_1, P = prod_{j=2..start} d_j;  y_0 = y_1 (2 + f_1) - y_2  (Numerov.h:398)
        const double d1 = 1. - g1[e];
        const double Y0s = W1[e] * fma(12., g1[e], 2.) / d1 - W2[e];
        out.y0_pos[e] = Y0s > 0.;
        out.y0_log2[e] = (fabs(Y0s) <= 1.7e308) ? log2(fabs(Y0s)) - log2(fabs(P[e])) : INFINITY;
        out.cfull[e] = count[e] + (((out.y0_pos[e] ? 0u : 1u) != prev[e]) ? 1 : 0);
        out.bad |= !(P[e] > 0.);
        out.steps += start[e] - 1;
        out.d_first[e] = d1;
    }
}

// ---------------------------------------------------------------------------------------------------------
// lanes kernel (component entry point): every warp's lanes must share (tab, l)
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) numerov_lanes_fast_kernel(GridDev g, NumerovLaneArgs a)
{
    __shared__ double2 sbuf[4 * 64];
    const int warp = threadIdx.x >> 5;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if ((k & ~31) >= a.n_lanes) return;
    const int kk = min(k, a.n_lanes - 1);
    const int l = a.l[kk];
    FastOut<1> o;
    const double E1[1] = { a.E[kk] };
    fast_sweep<1>(g, a.atab + (size_t)a.tab[kk] * g.N, (double)(l * (l + 1)), E1, sbuf + warp * 64, o);
    if (k < a.n_lanes) {
        if (a.y0_sign) a.y0_sign[k] = o.y0_pos[0];
        if (a.y0_log2) a.y0_log2[k] = o.y0_log2[0];
        if (a.count) a.count[k] = o.bad ? -1 : o.cfull[0];
    }
}

void launch_numerov_lanes_fast(const GridDev& g, const NumerovLaneArgs& a, cudaStream_t st)
{
    numerov_lanes_fast_kernel<<<(a.n_lanes + 127) / 128, 128, 0, st>>>(g, a);
}

// ---------------------------------------------------------------------------------------------------------
// fused search: one warp per orbital, all rounds in one launch
// ---------------------------------------------------------------------------------------------------------
template <int EPL>
__global__ void __launch_bounds__(128, 1) search_fused_kernel(GridDev g, const double* __restrict__ atab_all, const AtomDev* atoms,
                                                           const OrbitalDev* orbs, const AtomState* astate, SearchState* ss, int n_orbs,
                                                           unsigned long long* work, int warm_start)
{
    __shared__ double2 sbuf[4 * 64];
    constexpr int K = 32 * EPL;                           // trial energies per round
    const unsigned full = 0xffffffffu;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int k = blockIdx.x * 4 + warp;
    if (k >= n_orbs) return;
    const OrbitalDev ob = orbs[k];
    if (astate[ob.atom].done) return;
    const double* atab = atab_all + (size_t)ob.tab * g.N;
    const double ll1 = (double)(ob.l * (ob.l + 1));
    const double Z = (double)atoms[ob.atom].Z;
    double lo = -Z * Z - 1., hi = kTopEnergy;             // DFTAtom.cpp:407,499
    double ylog = 0.;
    long long steps = 0;
    int rounds = 0;
    // Sampling.  The predicate is monotone, so ANY ascending set of trial energies brackets the same root; what the set
    // looks like only decides how fast the bracket shrinks.  Two shapes are used:
    //   uniform : K points that cut [lo, hi] into K + 1 equal parts (cold start, and whenever nothing better is known);
    //   ladder  : a two-sided geometric ladder  c -+ eps g^m  (m = 0 .. K/2 - 1, outermost offset = R) around an estimate
    //             c of the root.  Round 0 of SCF step >= 1 centres it on the previous step's eigenvalue; later rounds
    //             centre it on the zero of y0(E) interpolated (inverse cubic Lagrange) through the samples next to the
    //             sign change, with R = 4 |cubic - secant| as the trust radius.  A wrong estimate only leaves a wide
    //             bracket (next round: uniform); a good one closes the bracket to 1e-12 in 2-3 rounds instead of ~6.
    const bool warm = warm_start && ss[k].pad == 1;
    bool ladder = warm;
    double c_est = ss[k].E, radius = 8.4;
    constexpr double kEps = 2.4e-13;
    for (int round = 0; round < 64 && bracket_open(lo, hi); ++round) {
        // point j of the round (j = 0..K-1, ascending in energy) lives in lane j % 32, slot j / 32
        double E[EPL];
        const double lg = ladder ? log2(fmax(radius, 2. * kEps) / kEps) / (double)(K / 2 - 1) : 0.;
#pragma unroll
        for (int e = 0; e < EPL; ++e) {
            const int j = e * 32 + lane;
            if (ladder) {
                const int half = K / 2;
                const int mstep = (j < half) ? (half - 1 - j) : (j - half);            // 0 = closest to the estimate
                const double off = kEps * exp2((double)mstep * lg);
                E[e] = fmin(fmax((j < half) ? c_est - off : c_est + off, lo), hi);
            } else {
                E[e] = lo + (hi - lo) * ((double)(j + 1) / (double)(K + 1));
            }
        }
        FastOut<EPL> o;
        fast_sweep<EPL>(g, atab, ll1, E, sbuf + warp * 64, o);
        if (__any_sync(full, o.bad)) {
            // a non-positive 1 - f/12 inside the sweep (grid far too coarse for this energy): generic path
#pragma unroll
            for (int e = 0; e < EPL; ++e) {
                const LaneOut s = sweep_lane(g, atab, ob.l, E[e], ob.want);
                o.cfull[e] = s.count_full; o.d_first[e] = s.d_first; o.y0_log2[e] = s.y0_log2; o.y0_pos[e] = s.y0_pos;
            }
        }
        steps += o.steps;
        ++rounds;
        unsigned m_hi[EPL];
#pragma unroll
        for (int e = 0; e < EPL; ++e) m_hi[e] = __ballot_sync(full, o.cfull[e] > ob.want + (o.d_first[e] < 0. ? 1 : 0));
        // virtual bisection over the K sampled points (what a bisection restricted to them would do)
        int lo_i = -1, hi_i = K, lm = -1;
        while (hi_i - lo_i > 1) {
            const int mid = (lo_i + hi_i) >> 1;
            lm = mid;
            bool high = false;
#pragma unroll
            for (int e = 0; e < EPL; ++e) if ((mid >> 5) == e) high = (m_hi[e] >> (mid & 31)) & 1u;
            if (high) hi_i = mid; else lo_i = mid;
        }
        double e_lo = lo, e_hi = hi, yl = 0.;
#pragma unroll
        for (int e = 0; e < EPL; ++e) {
            const double a_lo = __shfl_sync(full, E[e], max(lo_i, 0) & 31), a_hi = __shfl_sync(full, E[e], min(hi_i, K - 1) & 31);
            const double a_y = __shfl_sync(full, o.y0_log2[e], lm & 31);
            if (lo_i >= 0 && (lo_i >> 5) == e) e_lo = a_lo;
            if (hi_i < K && (hi_i >> 5) == e) e_hi = a_hi;
            if ((lm >> 5) == e) yl = a_y;
        }
        // estimate of the root for the next round: zero of y0(E) through the samples around the sign change
        ladder = false;
        if (EPL == 1 && lo_i >= 0 && hi_i < K && e_lo < e_hi) {
            double Ek[4], yk[4], lgv[4];
            bool ok[4];
            double ref = -INFINITY;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int idx = lo_i - 1 + q;
                const int src_lane = min(max(idx, 0), 31);
                Ek[q] = __shfl_sync(full, E[0], src_lane);
                lgv[q] = __shfl_sync(full, o.y0_log2[0], src_lane);
                yk[q] = __shfl_sync(full, o.y0_pos[0], src_lane) ? 1. : -1.;
                ok[q] = idx >= 0 && idx < K && lgv[q] > -1e300 && lgv[q] < 1e300;
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
                        double w = Ek[q];
#pragma unroll
                        for (int r = 0; r < 4; ++r) {
                            const bool ur = (r == 0) ? use0 : (r == 3 ? use3 : true);
                            if (r == q || !ur) continue;
                            w *= (0. - yk[r]) / (yk[q] - yk[r]);
                        }
                        num += w;
                    }
                    Eh = num;
                }
                if (!(Eh > e_lo && Eh < e_hi)) Eh = E2;
                if (Eh > e_lo && Eh < e_hi) {
                    c_est = Eh;
                    const double trust = (use0 || use3) ? 4. * fabs(Eh - E2) : 0.25 * (e_hi - e_lo);
                    radius = fmin(fmax(trust, 16. * kEps), fmax(e_hi - Eh, Eh - e_lo));
                    ladder = true;
                }
            }
        }
        lo = e_lo; hi = e_hi; ylog = yl;
    }
    if (lane == 0) {
        SearchState s = ss[k];
        s.bot = lo; s.top = hi; s.E = lo;                                    // level.E = BottomEnergy, DFTAtom.cpp:534
        s.y0_log2 = ylog;
        s.converged = (hi - lo < kEnergyTol) && (ylog < 49.828921423310435); // DFTAtom.cpp:528
        s.stage = 3;
        s.pad = 1;                                                           // E is a valid warm start for the next step
        ss[k] = s;
    }
    if (work) {
#pragma unroll
        for (int o = 16; o; o >>= 1) steps += __shfl_xor_sync(full, steps, o);
        if (lane == 0) {
            atomicAdd(work, (unsigned long long)steps);
            atomicAdd(work + DFTATOM_K_MATCH, 1ULL);                          // orbital solves
            atomicAdd(work + DFTATOM_K_DENSITY, (unsigned long long)rounds);  // search rounds
        }
    }
}

void launch_search_fused(const GridDev& g, const double* atab, const AtomDev* atoms, const OrbitalDev* orbs, const AtomState* astate,
                         SearchState* ss, int n_orbs, unsigned long long* work, int epl, int warm_start, cudaStream_t st)
{
    if (epl == 2) search_fused_kernel<2><<<(n_orbs + 3) / 4, 128, 0, st>>>(g, atab, atoms, orbs, astate, ss, n_orbs, work, warm_start);
    else search_fused_kernel<1><<<(n_orbs + 3) / 4, 128, 0, st>>>(g, atab, atoms, orbs, astate, ss, n_orbs, work, warm_start);
}

}  // namespace dft
