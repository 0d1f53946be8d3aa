// Parallel-in-r Numerov shooting: the energy search of one level by a whole CTA.
//
// Replaces the same reference code as numerov_fast.cu (DFTAtom.cpp:493-541, :566-604; Numerov.h:272-401) with the same
// search (numerov_common.cuh: Bracket), but one thread-block CLUSTER per orbital: warp s of the cluster = radial segment s
// (counted from the far end), lane = trial energy.  The CTAs of a cluster (4 warps each) sit on different SMs, so the
// segments of one orbital get several FP64 pipes (a B200 SM issues one FP64 warp-instruction per 2 cycles and scheduler:
// 32 warps on ONE SM would be throughput-bound again); the per-segment results are exchanged through distributed
// shared memory.  One inward sweep of N nodes is a chain of N dependent steps; cut into S segments it becomes
//   pass 1  every (segment, energy) pushes the two basis states (W, D) = (1, 0), (0, 1) through its segment of the
//           three-term recurrence: the segment's 2x2 transfer matrix.  The segment that contains the far seeds of an
//           energy (Numerov.h:294-303) runs the real solution instead; segments beyond the seeds are idle;
//   scan    the S maps of an energy are applied in order (shared memory), which gives every segment its entry state;
//   pass 2  every segment re-runs from its entry state and counts the sign changes of y (the Sturm count).
// 2.25x the arithmetic of the serial sweep, 2/S of its depth: this is the kernel for few active orbitals (late SCF
// steps, small batches), where the serial sweep leaves the GPU idle and the step time is pure latency.
// The recurrence is the scaled difference form of numerov_fast.cu:
//   g = f/12, d = 1 - g, s_i = 1 - d_i d_{i+1},  D_i = D_{i+1} + 10 g_i W_i + s_i W_{i+1},  W_{i-1} = W_i + D_i,
//   y_i = W_i / (P_i d_i), P_i = prod_{j>i} d_j > 0.
#include "numerov_sweep.cuh"
#include <cooperative_groups.h>
#include <cstdio>

namespace cg = cooperative_groups;

namespace dft {

constexpr int kSegMax = 32;        // segments per orbital = 4 warps x cluster size (<= 8 CTAs: the portable maximum)
constexpr int kSegWarps = 4;       // warps (segments) per CTA

struct SegShared {                 // one per CTA: the results of its 4 segments; y0s / d1 / bad are used in the rank-0 CTA only
    double m[4][kSegWarps][32];    // normal: transfer matrix (ww, wd, dw, dd); seeded: outgoing state (W, D) in [0], [1]
    double pseg[kSegWarps][32];    // product of d_i d_{i+1} over the even nodes of the segment (nodes <= start)
    double y0s[32], d1[32];        // bottom segment: scaled y0 and d_1 per energy
    double ct[4][32];              // the composition of this CTA's 4 segment maps (same encoding), kind in ctk
    int ctk[32];
    int meta[kSegWarps][32];       // kind (bits 0-1: 0 idle, 1 normal, 2 seeded) | sign of y at the lowest node (bit 2) | sign changes inside the segment << 3
                                   // (written once per round, by pass 1; the scan of the other warps reads its kind bits)
    int res[kSegWarps][32];        // normal segments: the same fields for the REAL solution, written after the scan (a separate word: no
                                   // write into meta while another warp's scan may still be reading it)
    int bad[32];
};

struct SegRoundOut { int cfull, y0_pos, start; double d_first, y0_log2; };

// One inward sweep of 32 trial energies (lane = energy) by all S = CL x 4 warps of the cluster (warp = radial segment):
// pass 1, scan, pass 2, totals.  Every warp returns the same result.  Ends with a cluster barrier: the shared results have
// been consumed when it returns.
__device__ __forceinline__ SegRoundOut seg_round(const GridDev& g, const double* __restrict__ atab, double ll1, int l, int want, double E,
                                                 cg::cluster_group& cluster, SegShared& sh, SegShared* sh0, double2* sbuf, int S, int rank,
                                                 int wl, int w, int lane, unsigned long long* work)
{
    const unsigned full = 0xffffffffu;
    const double kappa = sqrt(2. * fabs(E));
    const int start = start_index(g, kappa);
    int imax = start;
#pragma unroll
    for (int o = 16; o; o >>= 1) imax = max(imax, __shfl_xor_sync(full, imax, o));
    // segments are whole 32-node tiles of the staged tables (numerov_sweep.cuh): only the tile with the far seeds, the top tile
    // (partial) and tile 0 (node 0 is not swept) take the general per-node path, every interior boundary is tile-aligned
    const int m_top = imax >> 5;
    const int tps = (m_top + S) / S;                       // ceil((m_top + 1) / S) tiles per segment
    const int m_hi = m_top - w * tps;
    const int m_lo = max(m_hi - tps + 1, 0);
    const bool valid = m_hi >= 0;
    const int top = valid ? min(imax, (m_hi << 5) + 31) : 0;
    const int bot = max(m_lo << 5, 1);
    // this segment's role for this energy: the seeds are nodes start, start - 1; the segment that contains start - 1 runs
    // the real solution from the seeds (w_start itself depends on the tables only)
    const int kind = (!valid || start - 1 < bot) ? 0 : (start - 1 > top ? 1 : 2);
    const bool is_bottom = valid && bot == 1;

    // ---------------- pass 1: transfer matrix (two basis chains) or the seeded real solution ----------------
    FastOut<2> o;
    {
        SweepIn<2> in;
        in.E[0] = E; in.E[1] = E;
        in.running[0] = kind == 1; in.W_in[0] = 1.; in.D_in[0] = 0.;
        in.running[1] = kind == 1; in.W_in[1] = 0.; in.D_in[1] = 1.;
        in.start[0] = kind == 2 ? start : -1;
        in.start[1] = -1;
        o.bad = 0;
        if (valid && __any_sync(full, kind != 0)) range_sweep<2>(g, atab, ll1, in, top, bot, sbuf + wl * 64, o);
        sh.meta[wl][lane] = kind | ((int)o.prev[0] << 2) | (o.count[0] << 3);      // normal segments: count and sign follow in pass 2
        sh.pseg[wl][lane] = kind ? o.P[0] : 1.;
        sh.m[0][wl][lane] = o.W[0]; sh.m[2][wl][lane] = o.D[0];      // normal: (ww, dw); seeded: outgoing state (W, D)
        sh.m[1][wl][lane] = o.W[1]; sh.m[3][wl][lane] = o.D[1];      // normal: (wd, dd)
        if (w == 0) sh.bad[lane] = 0;
        if (kind == 2 && is_bottom) { sh0->y0s[lane] = o.Y0s[0]; sh0->d1[lane] = o.d_first[0]; }
        // the map of the whole CTA (its 4 segments in order): identity / matrix / constant state
        __syncthreads();
        if (wl == 0) {
            int ck = 0;
            double c0 = 1., c1 = 0., c2 = 0., c3 = 1.;
#pragma unroll
            for (int vl = 0; vl < kSegWarps; ++vl) {
                const int kv = sh.meta[vl][lane] & 3;
                const double m0 = sh.m[0][vl][lane], m1 = sh.m[1][vl][lane], m2 = sh.m[2][vl][lane], m3 = sh.m[3][vl][lane];
                if (kv == 2) { ck = 2; c0 = m0; c2 = m2; }
                else if (kv == 1) {
                    if (ck == 2) { const double na = fma(m0, c0, m1 * c2), nb = fma(m2, c0, m3 * c2); c0 = na; c2 = nb; }
                    else if (ck == 1) {
                        const double n0 = fma(m0, c0, m1 * c2), n1 = fma(m0, c1, m1 * c3), n2 = fma(m2, c0, m3 * c2), n3 = fma(m2, c1, m3 * c3);
                        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
                    } else { ck = 1; c0 = m0; c1 = m1; c2 = m2; c3 = m3; }
                }
            }
            sh.ctk[lane] = ck;
            sh.ct[0][lane] = c0; sh.ct[1][lane] = c1; sh.ct[2][lane] = c2; sh.ct[3][lane] = c3;
        }
        cluster.sync();
        if (kind == 2 && o.bad) atomicOr(&sh0->bad[lane], 1);
    }
    // ---------------- scan: entry state of this segment ----------------
    double A = 0., Bd = 0.;
    if (kind == 1) {
        // the CTAs before this one (their composed maps, distributed shared memory), then the warps before this one
#pragma unroll 4
        for (int rk = 0; rk < rank; ++rk) {
            const SegShared* r = cluster.map_shared_rank(&sh, rk);
            const int kv = r->ctk[lane];
            const double m0 = r->ct[0][lane], m1 = r->ct[1][lane], m2 = r->ct[2][lane], m3 = r->ct[3][lane];
            if (kv == 2) { A = m0; Bd = m2; }
            else if (kv == 1) { const double na = fma(m0, A, m1 * Bd), nb = fma(m2, A, m3 * Bd); A = na; Bd = nb; }
        }
        for (int vl = 0; vl < wl; ++vl) {
            const int kv = sh.meta[vl][lane] & 3;
            const double m0 = sh.m[0][vl][lane], m1 = sh.m[1][vl][lane], m2 = sh.m[2][vl][lane], m3 = sh.m[3][vl][lane];
            if (kv == 2) { A = m0; Bd = m2; }
            else if (kv == 1) { const double na = fma(m0, A, m1 * Bd), nb = fma(m2, A, m3 * Bd); A = na; Bd = nb; }
        }
    }
    // ---------------- sign changes of the real solution inside the normal segments ----------------
    // The real solution is W_i = A u_i + B v_i with (u, v) the two basis chains of pass 1.  X_i = (u_i, v_i) turns counter-
    // clockwise as i decreases (its Wronskian keeps its sign: d_i d_{i+1} > 0) by less than pi per node, so W changes sign
    // between two nodes exactly when X crosses the line L perpendicular to (A, B), and u when X crosses the axis u = 0.  Lines
    // through the origin are crossed alternately: #L = #(sign changes of u, counted by pass 1) + [end past L] - [start past L],
    // "past L" = on the far side of L within the half turn that starts at the axis:  [sigma_X sigma_L W(X) <= 0],
    // sigma_X = -sign(u), sigma_L = sign(B).  No second sweep: the Sturm count of a segment costs its transfer matrix only.
    // (Exception: the bottom segment of an l = 3 level, where 1 - f_1/12 < 0 flips y_1, SURVEY fact 6: re-run as before.)
    const bool flip_node1 = kind == 1 && is_bottom && !(o.d_first[0] > 0.);
    if (__any_sync(full, flip_node1)) {
        SweepIn<1> in;
        in.E[0] = E; in.running[0] = kind == 1; in.W_in[0] = A; in.D_in[0] = Bd; in.start[0] = -1;
        FastOut<1> o2;
        range_sweep<1>(g, atab, ll1, in, top, bot, sbuf + wl * 64, o2);
        if (kind == 1) {
            sh.res[wl][lane] = 1 | ((int)o2.prev[0] << 2) | (o2.count[0] << 3);
            if (is_bottom) { sh0->y0s[lane] = o2.Y0s[0]; sh0->d1[lane] = o2.d_first[0]; }
            if (o2.bad) atomicOr(&sh0->bad[lane], 1);
        }
    } else if (kind == 1) {
        const double ue = o.W[0], ve = o.W[1];
        const double We = fma(A, ue, Bd * ve);                       // the real W at the lowest node of the segment
        const double sl = Bd < 0. ? -1. : (Bd > 0. ? 1. : (A >= 0. ? 1. : -1.));
        const int past0 = (sl * A >= 0.) ? 1 : 0;                    // X_start = (1, 0): sigma_X = -1, W = A
        const int past1 = ((ue >= 0. ? 1. : -1.) * sl * We >= 0.) ? 1 : 0;
        const int cnt = o.count[0] + past1 - past0;
        const unsigned sgn = (unsigned)hi32(We) >> 31;
        sh.res[wl][lane] = 1 | ((int)sgn << 2) | (cnt << 3);
        if (is_bottom) { sh0->y0s[lane] = fma(A, o.Y0s[0], Bd * o.Y0s[1]); sh0->d1[lane] = o.d_first[0]; }
        if (o.bad) atomicOr(&sh0->bad[lane], 1);
    }
    cluster.sync();

    // ---------------- totals (every warp redundantly, so that all warps hold the same bracket) ----------------
    int cfull = 0;
    double Ptot = 1.;
    unsigned pbot = 0;
    int have_bottom = 0;
#pragma unroll 4
    for (int v = 0; v < S; ++v) {
        const SegShared* r = cluster.map_shared_rank(&sh, v / kSegWarps);
        const int mk = r->meta[v % kSegWarps][lane];
        const int mv = (mk & 3) == 1 ? r->res[v % kSegWarps][lane] : mk;       // normal segments: the real solution's count and sign
        const double pv = r->pseg[v % kSegWarps][lane];
        if (mv & 3) { cfull += mv >> 3; pbot = (unsigned)(mv >> 2) & 1u; Ptot *= pv; have_bottom = 1; }
    }
    const double Y0s = sh0->y0s[lane], d1v = sh0->d1[lane];
    int y0_pos = Y0s > 0.;
    double y0_log2 = (fabs(Y0s) <= 1.7e308) ? log2(fabs(Y0s)) - log2(fabs(Ptot)) : INFINITY;
    cfull += (((y0_pos ? 0u : 1u) != pbot) ? 1 : 0);
    int lane_bad = sh0->bad[lane] | !(Ptot > 0.) | !have_bottom | (start < 3);
    double d_first = d1v;
    if (__any_sync(full, lane_bad)) {
        if (work && threadIdx.x == 0) atomicAdd(work + DFTATOM_K_POTENTIAL, 1ULL);       // rounds that fell back to the serial sweep
        // a non-positive 1 - f/12 inside the sweep (grid far too coarse for this energy): generic serial path
        const LaneOut so = sweep_lane(g, atab, l, E, want);
        cfull = so.count_full; d_first = so.d_first; y0_log2 = so.y0_log2; y0_pos = so.y0_pos;
    }
    cluster.sync();
    SegRoundOut r;
    r.cfull = cfull; r.y0_pos = y0_pos; r.start = start; r.d_first = d_first; r.y0_log2 = y0_log2;
    return r;
}

__global__ void __launch_bounds__(32 * kSegWarps) search_seg_kernel(GridDev g, const double* __restrict__ atab_all, const AtomDev* atoms,
                                                             const OrbitalDev* orbs, const AtomState* astate, SearchState* ss, int n_orbs,
                                                             unsigned long long* work, const int* n_active_orbs, int threshold,
                                                             int warm_start)
{
    __shared__ SegShared sh;
    __shared__ double2 sbuf[kSegWarps * 64];
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned full = 0xffffffffu;
    const int CL = (int)cluster.num_blocks();             // CTAs per orbital
    const int S = CL * kSegWarps;
    const int rank = (int)cluster.block_rank();
    const int wl = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int w = rank * kSegWarps + wl;                  // segment index inside the orbital
    const int k = blockIdx.x / CL;
    SegShared* sh0 = cluster.map_shared_rank(&sh, 0);     // the rank-0 CTA collects y0 / d1 / bad
    if (k >= n_orbs) return;
    if (n_active_orbs && *n_active_orbs > threshold) return;       // the serial-in-r kernel is still the faster one
    const OrbitalDev ob = orbs[k];
    if (astate[ob.atom].done) return;
    const double* __restrict__ atab = atab_all + (size_t)ob.tab * g.N;
    const double ll1 = (double)(ob.l * (ob.l + 1));
    const double Z = (double)atoms[ob.atom].Z;
    Bracket b;
    b.lo = -Z * Z - 1.; b.hi = kTopEnergy;                // DFTAtom.cpp:407,499
    b.ylog = 0.;
    b.ladder = warm_start && ss[k].pad == 1;
    // first ladder: the levels move geometrically from one SCF step to the next (linear mixing), by tens of Hartree in
    // the first steps; centre = previous eigenvalue + last shift x (ratio of the last two shifts), radius = 1.5 x last shift
    const double e_prev = ss[k].E, s1 = ss[k].up_lo, s2 = ss[k].up_hi;
    const double ratio = (s2 != 0. && fabs(s1) < fabs(s2)) ? s1 / s2 : 0.;
    b.c_est = fmin(fmax(e_prev + s1 * ratio, b.lo), b.hi); b.radius = fmin(fmax(1.5 * fabs(s1), 1e-3), Z * Z + 51.);
    long long steps = 0;
    int rounds = 0;

    for (int round = 0; round < 64 && bracket_open(b.lo, b.hi); ++round) {
        // every warp computes the same 32 trial energies
        const double E = sample_energy(b, lane);
        const SegRoundOut r = seg_round(g, atab, ll1, ob.l, ob.want, E, cluster, sh, sh0, sbuf, S, rank, wl, w, lane, work);
        if (w == 0) steps += r.start - 1;
        ++rounds;
        update_bracket(b, E, r.cfull > ob.want + (r.d_first < 0. ? 1 : 0), r.y0_pos, r.y0_log2);
    }
    if (w == 0 && lane == 0) {
        SearchState s = ss[k];
        s.bot = b.lo; s.top = b.hi; s.E = b.lo;                              // level.E = BottomEnergy, DFTAtom.cpp:534
        s.y0_log2 = b.ylog;
        s.converged = (b.hi - b.lo < kEnergyTol) && (b.ylog < 49.828921423310435); // DFTAtom.cpp:528
        s.stage = 3;
        s.up_hi = s.pad == 1 ? s.up_lo : 0.;                                 // the last two shifts of the level
        s.up_lo = s.pad == 1 ? b.lo - e_prev : Z * Z;
        s.pad = 1;
        ss[k] = s;
    }
    if (work && w == 0) {
#pragma unroll
        for (int o = 16; o; o >>= 1) steps += __shfl_xor_sync(full, steps, o);
        if (lane == 0) {
            atomicAdd(work, (unsigned long long)steps);
            atomicAdd(work + DFTATOM_K_MATCH, 1ULL);                          // orbital solves
            atomicAdd(work + DFTATOM_K_DENSITY, (unsigned long long)rounds);  // search rounds
            atomicAdd(work + 8 + min(rounds, 15), 1ULL);                      // histogram (debug aid)
            atomicAdd(work + 24 + min(ob.l, 3), (unsigned long long)rounds);  // rounds and solves with >= 4 rounds per l (debug aid)
            if (rounds >= 4) atomicAdd(work + 28 + min(ob.l, 3), 1ULL);
        }
    }
}

void launch_search_seg(const GridDev& g, const double* atab, const AtomDev* atoms, const OrbitalDev* orbs, const AtomState* astate,
                       SearchState* ss, int n_orbs, unsigned long long* work, int segments, const int* n_active_orbs, int threshold,
                       int warm_start, cudaStream_t st)
{
    int CL = 1;
    while (CL * 2 * kSegWarps <= segments && CL * 2 * kSegWarps <= kSegMax) CL *= 2;      // 1, 2, 4 or 8 CTAs per orbital
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(n_orbs * CL));
    cfg.blockDim = dim3(32 * kSegWarps);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, search_seg_kernel, g, atab, atoms, orbs, astate, ss, n_orbs, work, n_active_orbs, threshold, warm_start);
}

// ---------------------------------------------------------------------------------------------------------
// lanes kernel (component entry point / C5b microbench): one cluster per group of 32 lanes that share (tab, l)
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32 * kSegWarps) numerov_lanes_seg_kernel(GridDev g, NumerovLaneArgs a)
{
    __shared__ SegShared sh;
    __shared__ double2 sbuf[kSegWarps * 64];
    cg::cluster_group cluster = cg::this_cluster();
    const int CL = (int)cluster.num_blocks();
    const int S = CL * kSegWarps;
    const int rank = (int)cluster.block_rank();
    const int wl = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int w = rank * kSegWarps + wl;
    const int grp = blockIdx.x / CL;
    SegShared* sh0 = cluster.map_shared_rank(&sh, 0);
    if (grp * 32 >= a.n_lanes) return;
    const int k = grp * 32 + lane, kk = min(k, a.n_lanes - 1);
    const int l = a.l[kk];
    const SegRoundOut r = seg_round(g, a.atab + (size_t)a.tab[kk] * g.N, (double)(l * (l + 1)), l, a.limit ? a.limit[kk] : 0, a.E[kk], cluster, sh,
                                    sh0, sbuf, S, rank, wl, w, lane, nullptr);
    if (w == 0 && k < a.n_lanes) {
        if (a.y0_sign) a.y0_sign[k] = r.y0_pos;
        if (a.y0_log2) a.y0_log2[k] = r.y0_log2;
        if (a.count) a.count[k] = r.cfull;
    }
}

static int seg_cluster_size(int segments)
{
    int CL = 1;
    while (CL * 2 * kSegWarps <= segments && CL * 2 * kSegWarps <= kSegMax) CL *= 2;      // 1, 2, 4 or 8 CTAs per orbital
    return CL;
}

void launch_numerov_lanes_seg(const GridDev& g, const NumerovLaneArgs& a, int segments, cudaStream_t st)
{
    const int CL = seg_cluster_size(segments);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(((a.n_lanes + 31) / 32) * CL));
    cfg.blockDim = dim3(32 * kSegWarps);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, numerov_lanes_seg_kernel, g, a);
}

}  // namespace dft
