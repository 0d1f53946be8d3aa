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

struct FastOut { int cfull; int y0_pos; int bad; int steps; double y0_log2; double d_first; };

// warp-collective; sbuf = this warp's double-buffered tile staging area [2][32]
__device__ __forceinline__ FastOut fast_sweep(const GridDev& g, const double* __restrict__ atab, double nll1, double E, double2* sbuf)
{
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const double kappa = sqrt(2. * fabs(E));
    const int start = start_index(g, kappa);
    int imax = start;
#pragma unroll
    for (int o = 16; o; o >>= 1) imax = max(imax, __shfl_xor_sync(full, imax, o));
    const int nmax = g.N - 1;

    double W1 = 0., W2 = 0., d1 = 1., dd1 = 1., n1 = 2., P = 1.;
    unsigned prev = 0;
    int count = 0, bad = 0;

    int m = imax >> 5;
    // prefetch the top tile: lane j holds node 32 m + 31 - j
    double pa, pb, pc;
    {
        const int i = min((m << 5) + 31 - lane, nmax);
        pa = __ldg(atab + i); pb = __ldg(g.b12 + i); pc = __ldg(g.c6 + i);
    }
    int cur = 0;
    for (; m >= 0; --m) {
        sbuf[cur * 32 + lane] = make_double2(fma(nll1, pb, pa), pc);
        __syncwarp();
        if (m > 0) {
            const int i = ((m - 1) << 5) + 31 - lane;
            pa = __ldg(atab + i); pb = __ldg(g.b12 + i); pc = __ldg(g.c6 + i);
        }
        const int hi_i = (m << 5) + 31, lo_i = m << 5;
        const bool uniform = (start >= hi_i + 2) || (start < lo_i);
        const double2* tile = sbuf + cur * 32;
        if (m > 0 && __all_sync(full, uniform)) {
            // ---- fast tile: every lane is either fully inside its sweep or has not started yet (W stays 0) ----
            unsigned sb = 0;
#pragma unroll
            for (int k = 0; k < 32; ++k) {
                const double2 t = tile[k];
                const double W = fma(n1, W1, -(dd1 * W2));
                sb = __funnelshift_l((unsigned)hi32(W), sb, 1);
                const double d = fma(E, t.y, t.x);
                const double dd = d * d1;
                n1 = fma(-10., d, 12.);
                if (k & 1) P *= dd;                     // even node index: pairs (i, i+1)
                W2 = W1; W1 = W; d1 = d; dd1 = dd;
            }
            const unsigned x = sb ^ ((sb >> 1) | (prev << 31));
            count += __popc(x);
            prev = sb & 1u;
        } else {
            // ---- general tile: seeds (far boundary values), the last tile down to i = 1, sign of d ----
            for (int k = 0; k < 32; ++k) {
                const int i = hi_i - k;
                if (i < 1) break;
                const double2 t = tile[k];
                const double d = fma(E, t.y, t.x);
                if (i <= start) {
                    double W, dd;
                    if (i == start) {                      // w_start = d_start far(start)   (Numerov.h:294-298)
                        W = d * far_value(g, kappa, i);
                        dd = d; P = 1.; count = 0; prev = 0;
                        bad |= !(d > 0.);
                    } else if (i == start - 1) {           // w_{start-1}                    (Numerov.h:300-303)
                        W = d * far_value(g, kappa, i) * d1;
                        dd = d * d1;
                        bad |= !(d > 0.);
                    } else {
                        W = fma(n1, W1, -(dd1 * W2));
                        dd = d * d1;
                        const unsigned sy = ((unsigned)hi32(W) ^ (unsigned)hi32(d)) >> 31;    // y_i = W_i / (P_i d_i), P_i > 0
                        count += (sy != prev);
                        prev = sy;
                        if (i == 2) bad |= !(d > 0.);
                    }
                    if (!(i & 1)) P *= dd;
                    n1 = fma(-10., d, 12.);
                    W2 = W1; W1 = W; d1 = d; dd1 = dd;
                }
            }
        }
        cur ^= 1;
    }
    // W1 = W_1, W2 = W_2, d1 = d_1, P = prod_{j=2..start} d_j;  y_0 = y_1 (2 + f_1) - y_2  (Numerov.h:398)
    const double Y0s = W1 * fma(-12., d1, 14.) / d1 - W2;
    FastOut o;
    o.y0_pos = Y0s > 0.;
    o.y0_log2 = (fabs(Y0s) <= 1.7e308) ? log2(fabs(Y0s)) - log2(fabs(P)) : INFINITY;
    o.cfull = count + (((o.y0_pos ? 0u : 1u) != prev) ? 1 : 0);
    o.bad = bad | !(P > 0.);
    o.steps = start - 1;
    o.d_first = d1;
    return o;
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
    const FastOut o = fast_sweep(g, a.atab + (size_t)a.tab[kk] * g.N, -(double)(l * (l + 1)), a.E[kk], sbuf + warp * 64);
    if (k < a.n_lanes) {
        if (a.y0_sign) a.y0_sign[k] = o.y0_pos;
        if (a.y0_log2) a.y0_log2[k] = o.y0_log2;
        if (a.count) a.count[k] = o.bad ? -1 : o.cfull;
    }
}

void launch_numerov_lanes_fast(const GridDev& g, const NumerovLaneArgs& a, cudaStream_t st)
{
    numerov_lanes_fast_kernel<<<(a.n_lanes + 127) / 128, 128, 0, st>>>(g, a);
}

// ---------------------------------------------------------------------------------------------------------
// fused search: one warp per orbital, all rounds in one launch
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) search_fused_kernel(GridDev g, const double* __restrict__ atab_all, const AtomDev* atoms,
                                                           const OrbitalDev* orbs, const AtomState* astate, SearchState* ss, int n_orbs,
                                                           unsigned long long* work)
{
    __shared__ double2 sbuf[4 * 64];
    const unsigned full = 0xffffffffu;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int k = blockIdx.x * 4 + warp;
    if (k >= n_orbs) return;
    const OrbitalDev ob = orbs[k];
    if (astate[ob.atom].done) return;
    const double* atab = atab_all + (size_t)ob.tab * g.N;
    const double nll1 = -(double)(ob.l * (ob.l + 1));
    const double Z = (double)atoms[ob.atom].Z;
    double lo = -Z * Z - 1., hi = kTopEnergy;             // DFTAtom.cpp:407,499
    double ylog = 0.;
    long long steps = 0;
    for (int round = 0; round < 64 && bracket_open(lo, hi); ++round) {
        const double E = lo + (hi - lo) * ((double)(lane + 1) / 33.);
        FastOut o = fast_sweep(g, atab, nll1, E, sbuf + warp * 64);
        int cfull = o.cfull; int off = o.d_first < 0.;
        if (__any_sync(full, o.bad)) {
            // a non-positive 1 - f/12 inside the sweep (grid far too coarse for this energy): generic path
            const LaneOut s = sweep_lane(g, atab, ob.l, E, ob.want);
            cfull = s.count_full; off = s.d_first < 0.; o.y0_log2 = s.y0_log2;
        }
        steps += o.steps;
        const unsigned m_hi = __ballot_sync(full, cfull > ob.want + off);
        int lo_i, hi_i, lm;
        virtual_bisect(m_hi, 32, lo_i, hi_i, lm);
        const double e_lo = __shfl_sync(full, E, max(lo_i, 0)), e_hi = __shfl_sync(full, E, min(hi_i, 31));
        ylog = __shfl_sync(full, o.y0_log2, lm);
        if (lo_i >= 0) lo = e_lo;
        if (hi_i < 32) hi = e_hi;
    }
    if (lane == 0) {
        SearchState s = ss[k];
        s.bot = lo; s.top = hi; s.E = lo;                                    // level.E = BottomEnergy, DFTAtom.cpp:534
        s.y0_log2 = ylog;
        s.converged = (hi - lo < kEnergyTol) && (ylog < 49.828921423310435); // DFTAtom.cpp:528
        s.stage = 3;
        ss[k] = s;
    }
    if (work) {
#pragma unroll
        for (int o = 16; o; o >>= 1) steps += __shfl_xor_sync(full, steps, o);
        if (lane == 0) atomicAdd(work, (unsigned long long)steps);
    }
}

void launch_search_fused(const GridDev& g, const double* atab, const AtomDev* atoms, const OrbitalDev* orbs, const AtomState* astate,
                         SearchState* ss, int n_orbs, unsigned long long* work, cudaStream_t st)
{
    search_fused_kernel<<<(n_orbs + 3) / 4, 128, 0, st>>>(g, atab, atoms, orbs, astate, ss, n_orbs, work);
}

}  // namespace dft
