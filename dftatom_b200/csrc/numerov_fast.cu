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
#include <cstdio>

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

}  // namespace dft
#include "numerov_sweep.cuh"
namespace dft {

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
// fused search: one warp per orbital, 32 lanes = 32 trial energies, all rounds in one launch (numerov_common.cuh: Bracket)
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1) search_fused_kernel(GridDev g, const double* __restrict__ atab_all, const AtomDev* atoms,
                                                           const OrbitalDev* orbs, const AtomState* astate, SearchState* ss, int n_orbs,
                                                           unsigned long long* work, const int* n_active_orbs, int threshold, int warm_start)
{
    __shared__ double2 sbuf[4 * 64];
    const unsigned full = 0xffffffffu;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int k = blockIdx.x * 4 + warp;
    if (k >= n_orbs) return;
    if (n_active_orbs && *n_active_orbs <= threshold) return;      // the parallel-in-r kernel has taken over
    const OrbitalDev ob = orbs[k];
    if (astate[ob.atom].done) return;
    const double* atab = atab_all + (size_t)ob.tab * g.N;
    const double ll1 = (double)(ob.l * (ob.l + 1));
    const double Z = (double)atoms[ob.atom].Z;
    Bracket b;
    b.lo = -Z * Z - 1.; b.hi = kTopEnergy;                // DFTAtom.cpp:407,499
    b.ylog = 0.;
    b.ladder = warm_start && ss[k].pad == 1;              // the previous step's eigenvalue is a valid centre
    // first ladder: the levels move geometrically from one SCF step to the next (linear mixing), by tens of Hartree in
    // the first steps; centre = previous eigenvalue + last shift x (ratio of the last two shifts), radius = 1.5 x last shift
    const double e_prev = ss[k].E, s1 = ss[k].up_lo, s2 = ss[k].up_hi;
    const double ratio = (s2 != 0. && fabs(s1) < fabs(s2)) ? s1 / s2 : 0.;
    b.c_est = fmin(fmax(e_prev + s1 * ratio, b.lo), b.hi); b.radius = fmin(fmax(1.5 * fabs(s1), 1e-3), Z * Z + 51.);
    long long steps = 0;
    int rounds = 0;
    for (int round = 0; round < 64 && bracket_open(b.lo, b.hi); ++round) {
        const double E1[1] = { sample_energy(b, lane) };
        FastOut<1> o;
        fast_sweep<1>(g, atab, ll1, E1, sbuf + warp * 64, o);
        if (__any_sync(full, o.bad)) {
            // a non-positive 1 - f/12 inside the sweep (grid far too coarse for this energy): generic path
            const LaneOut s = sweep_lane(g, atab, ob.l, E1[0], ob.want);
            o.cfull[0] = s.count_full; o.d_first[0] = s.d_first; o.y0_log2[0] = s.y0_log2; o.y0_pos[0] = s.y0_pos;
        }
        steps += o.steps;
        ++rounds;
        update_bracket(b, E1[0], o.cfull[0] > ob.want + (o.d_first[0] < 0. ? 1 : 0), o.y0_pos[0], o.y0_log2[0]);
    }
    if (lane == 0) {
        SearchState s = ss[k];
        s.bot = b.lo; s.top = b.hi; s.E = b.lo;                              // level.E = BottomEnergy, DFTAtom.cpp:534
        s.y0_log2 = b.ylog;
        s.converged = (b.hi - b.lo < kEnergyTol) && (b.ylog < 49.828921423310435); // DFTAtom.cpp:528
        s.stage = 3;
        s.up_hi = s.pad == 1 ? s.up_lo : 0.;                                 // the last two shifts of the level
        s.up_lo = s.pad == 1 ? b.lo - e_prev : Z * Z;
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
            atomicAdd(work + 8 + min(rounds, 15), 1ULL);                      // histogram (debug aid)
        }
    }
}

void launch_search_fused(const GridDev& g, const double* atab, const AtomDev* atoms, const OrbitalDev* orbs, const AtomState* astate,
                         SearchState* ss, int n_orbs, unsigned long long* work, const int* n_active_orbs, int threshold, int warm_start,
                         cudaStream_t st)
{
    search_fused_kernel<<<(n_orbs + 3) / 4, 128, 0, st>>>(g, atab, atoms, orbs, astate, ss, n_orbs, work, n_active_orbs, threshold, warm_start);
}

}  // namespace dft
