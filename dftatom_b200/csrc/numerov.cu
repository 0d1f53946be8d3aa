// Numerov shooting on the logarithmic grid, batched over (orbital, trial-energy) lanes.
//
// Replaces (reference DFTAtom/):  Numerov.h:272-349 SolveSchrodingerCountNodes, :351-401
// SolveSchrodingerSolutionInZero, :403-504 SolveSchrodingerMatchSolutionCompletely, the function object
// NumerovFunctionNonUniformGrid :73-196, and the per-level energy search of DFTAtom.cpp:493-541 / :566-604.
//
// Formulation.  With y'' = f y on the unit-step index grid and w_i = d_i y_i, d_i = 1 - f_i/12 (Numerov.h:510-513)
// the reference's step  w_{i-1} = 2 w_i - w_{i+1} + y_i f_i  is  w_{i-1} = (12 - 10 d_i)/d_i w_i - w_{i+1}.
// Multiplying by P_i = prod_{j>i} d_j gives the division-free form in W_i = w_i P_i
//        W_{i-1} = (12 - 10 d_i) W_i - d_i d_{i+1} W_{i+1}
// and sign(y_i) = sign(W_i) sign(P_{i-1}).  d_i(l,E) = a_i - l(l+1) b_i + E c_i comes from three tables
// (a: per potential; b, c: grid only) so there is no exp() in the loop (the reference calls it 2-3x per node).
#include "numerov_common.cuh"

namespace dft {

__global__ void numerov_lanes_kernel(GridDev g, NumerovLaneArgs a)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int kk = min(k, a.n_lanes - 1);                   // keep whole warps alive for the shuffles
    LaneOut o = sweep_lane(g, a.atab + (size_t)a.tab[kk] * g.N, a.l[kk], a.E[kk], a.limit[kk]);
    if (k < a.n_lanes) {
        if (a.y0_sign) a.y0_sign[k] = o.y0_pos;
        if (a.y0_log2) a.y0_log2[k] = o.y0_log2;
        if (a.count) a.count[k] = o.count;
    }
}

// SolveSchrodingerCountNodesFromNucleus (Numerov.h:204-270; public in the reference, without a caller - SURVEY 8(f) rank 4): the
// outward sweep from y_0 = 0, y_1 = r_1^(l+1) e^(-delta/2) (uniform grid: h^(l+1)) to the cut-off index, counting the sign changes
// of y; returns on overflow, on count > limit, and at the outer classical turning point.  Reference-shaped (one division per node):
// a component entry point (dftatom_numerov_lanes, impl = 3), not on the SCF path.
__global__ void numerov_lanes_outward_kernel(GridDev g, NumerovLaneArgs a)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= a.n_lanes) return;
    const double* __restrict__ atab = a.atab + (size_t)a.tab[k] * g.N;
    const int l = a.l[k], limit = a.limit[k];
    const double E = a.E[k];
    const double ll1 = (double)(l * (l + 1));
    const double kappa = sqrt(2. * fabs(E));
    const int steps = start_index(g, kappa);
    const double thr = 1. - g.delta * g.delta * (1. / 48.);       // d_i >= thr  <=>  Veff_i <= E
    auto gval = [&](int i) { return fma(-E, __ldg(g.c6 + i), fma(ll1, __ldg(g.b12 + i), __ldg(atab + i))); };
    double y = g.uniform ? pow(g.h, (double)l + 1.) : pow(__ldg(g.r + 1), (double)l + 1.) * exp(-0.5 * g.delta);
    double gq = gval(1), d = 1. - gq, f = 12. * gq;
    double wprev = 0., w = d * y;
    bool positive = y > 0., seen = d >= thr;
    int count = 0;
    for (int i = 2; i <= steps; ++i) {
        const double wn = 2. * w - wprev + y * f;
        wprev = w; w = wn;
        gq = gval(i); d = 1. - gq; f = 12. * gq;
        y = w / d;
        if (fabs(y) == INFINITY) break;
        if ((y > 0.) != positive) {
            if (++count > limit) break;
            positive = !positive;
        }
        if (d >= thr) seen = true;
        else if (seen) break;
    }
    if (a.count) a.count[k] = count;
    if (a.y0_sign) a.y0_sign[k] = 0;
    if (a.y0_log2) a.y0_log2[k] = 0.;
}

void launch_numerov_lanes_outward(const GridDev& g, const NumerovLaneArgs& a, cudaStream_t st)
{
    numerov_lanes_outward_kernel<<<(a.n_lanes + 63) / 64, 64, 0, st>>>(g, a);
}

void launch_numerov_lanes(const GridDev& g, const NumerovLaneArgs& a, cudaStream_t st)
{
    const int threads = 32;
    numerov_lanes_kernel<<<(a.n_lanes + threads - 1) / threads, threads, 0, st>>>(g, a);
}

// ---------------------------------------------------------------------------------------------------------
// Energy search: K-section instead of the reference's bisection, same predicates, same tolerances.
// ---------------------------------------------------------------------------------------------------------

__global__ void search_init_kernel(GridDev g, const AtomDev* atoms, const AtomState* astate, const OrbitalDev* orbs, SearchState* ss, int n_orbs)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_orbs) return;
    const OrbitalDev ob = orbs[k];
    if (astate[ob.atom].done) return;
    const double Z = (double)atoms[ob.atom].Z;
    SearchState s;
    const double bottom = -Z * Z - 1.;          // DFTAtom.cpp:407
    s.up_lo = bottom; s.up_hi = kTopEnergy;     // DFTAtom.cpp:499,568-569
    s.dn_lo = bottom; s.dn_hi = kTopEnergy;
    s.bot = bottom; s.top = kTopEnergy;
    s.y0_log2 = 0.; s.E = 0.;
    s.stage = 0; s.sgn_bottom = 0; s.converged = 0; s.pad = 0;
    ss[k] = s;
}

void launch_search_init(const GridDev& g, const AtomDev* atoms, const AtomState* astate, const OrbitalDev* orbs, SearchState* ss,
                        int n_orbs, cudaStream_t st)
{
    search_init_kernel<<<(n_orbs + 127) / 128, 128, 0, st>>>(g, atoms, astate, orbs, ss, n_orbs);
}

// one warp per orbital
__global__ void __launch_bounds__(32) search_round_kernel(GridDev g, const double* __restrict__ atab_all, const OrbitalDev* orbs,
                                                          const AtomState* astate, SearchState* ss, int n_orbs,
                                                          unsigned long long* work)
{
    const int k = blockIdx.x;
    if (k >= n_orbs) return;
    const OrbitalDev ob = orbs[k];
    if (astate[ob.atom].done) return;
    SearchState s = ss[k];
    if (s.stage == 3) return;
    const int lane = threadIdx.x;
    const double* atab = atab_all + (size_t)ob.tab * g.N;
    const unsigned full = 0xffffffffu;
    int lane_steps = 0;

    if (s.stage == 0) {
        const bool up_open = bracket_open(s.up_lo, s.up_hi);
        const bool dn_open = bracket_open(s.dn_lo, s.dn_hi);
        const bool same = (s.up_lo == s.dn_lo) && (s.up_hi == s.dn_hi);
        int K_up, K_dn, j; bool mine_up;
        if (same || !dn_open) { K_up = 32; K_dn = same ? 32 : 0; j = lane; mine_up = true; }
        else if (!up_open) { K_up = 0; K_dn = 32; j = lane; mine_up = false; }
        else { K_up = 16; K_dn = 16; j = lane & 15; mine_up = lane < 16; }
        const double lo = mine_up ? s.up_lo : s.dn_lo, hi = mine_up ? s.up_hi : s.dn_hi;
        const int K = mine_up ? K_up : K_dn;
        const double E = lo + (hi - lo) * ((double)(j + 1) / (double)(K + 1));
        const LaneOut o = sweep_lane(g, atab, ob.l, E, ob.want);
        lane_steps = o.steps;
        // A sweep that never met a classically allowed node lies below the bottom of the well: it has no physical
        // node.  The reference still reports 1 there for l = 3 (the sign flip of 1 - f_1/12 < 0 at the first grid
        // node) and only avoids that regime because it starts each level's bisection at
        // E_previous_level - 3 (DFTAtom.cpp:541).  All levels are searched concurrently here, from -Z^2-1, so the
        // predicate is made monotone instead.
        const int cnt = o.seen ? o.count : 0;
        const unsigned m_gt = __ballot_sync(full, cnt > ob.want);          // LocateInterval first loop, DFTAtom.cpp:578
        const unsigned m_ge = __ballot_sync(full, !(cnt < ob.want));       // second loop, :596
        int lo_i, hi_i, lm;
        if (K_up) {
            const unsigned m = (K_up == 32) ? m_gt : (m_gt & 0xffffu);
            virtual_bisect(m, K_up, lo_i, hi_i, lm);
            const double e_lo = __shfl_sync(full, E, max(lo_i, 0)), e_hi = __shfl_sync(full, E, min(hi_i, K_up - 1));
            if (lo_i >= 0) s.up_lo = e_lo;
            if (hi_i < K_up) s.up_hi = e_hi;
        }
        if (K_dn) {
            const int base = (K_dn == 32) ? 0 : 16;
            const unsigned m = (K_dn == 32) ? m_ge : (m_ge >> 16);
            virtual_bisect(m, K_dn, lo_i, hi_i, lm);
            const double e_lo = __shfl_sync(full, E, base + max(lo_i, 0)), e_hi = __shfl_sync(full, E, base + min(hi_i, K_dn - 1));
            if (lo_i >= 0) s.dn_lo = e_lo;
            if (hi_i < K_dn) s.dn_hi = e_hi;
        }
        if (!bracket_open(s.up_lo, s.up_hi) && !bracket_open(s.dn_lo, s.dn_hi)) {
            s.top = s.up_hi;        // TopEnergy = toe, DFTAtom.cpp:585
            s.bot = s.dn_hi;        // BottomEnergy = toe, :603
            if (s.bot > s.top) s.bot = s.top;
            s.stage = 1;
        }
    } else {
        // stage B: sign of y(0) against the sign at the bottom of the window, DFTAtom.cpp:513-533
        const bool first = (s.stage == 1);
        const int K = first ? 31 : 32;
        const double t = first ? (double)lane / 32. : (double)(lane + 1) / 33.;
        const double E = s.bot + (s.top - s.bot) * t;
        const LaneOut o = sweep_lane(g, atab, ob.l, E, ob.want);
        lane_steps = o.steps;
        if (first) s.sgn_bottom = __shfl_sync(full, o.y0_pos, 0);
        // "(delta > 0) == sgnBottom ? Bottom = E : Top = E": high side = sign differs
        unsigned m_hi = __ballot_sync(full, o.y0_pos != s.sgn_bottom);
        if (first) m_hi >>= 1;      // drop lane 0 (the bottom itself)
        int lo_i, hi_i, lm;
        virtual_bisect(m_hi, K, lo_i, hi_i, lm);
        const int off = first ? 1 : 0;
        const double e_lo = __shfl_sync(full, E, off + max(lo_i, 0)), e_hi = __shfl_sync(full, E, off + min(hi_i, K - 1));
        const double ylog = __shfl_sync(full, o.y0_log2, off + max(lm, 0));
        if (lo_i >= 0) s.bot = e_lo;
        if (hi_i < K) s.top = e_hi;
        s.y0_log2 = ylog;
        s.stage = 2;
        const bool small = !(s.top - s.bot >= kEnergyTol);
        const bool guard = (ylog < 49.828921423310435);        // |y0| < 1e15 and not NaN, DFTAtom.cpp:527-528
        if (small && guard) { s.converged = 1; s.stage = 3; }
        s.E = s.bot;                                            // level.E = BottomEnergy, :534
    }
    if (lane == 0) ss[k] = s;
    if (work) {
        int st = lane_steps;
#pragma unroll
        for (int o = 16; o; o >>= 1) st += __shfl_xor_sync(full, st, o);
        if (lane == 0) atomicAdd(work, (unsigned long long)st);
    }
}

void launch_search_round(const GridDev& g, const double* atab, const OrbitalDev* orbs, const AtomState* astate, SearchState* ss,
                         int n_orbs, unsigned long long* work, cudaStream_t st)
{
    search_round_kernel<<<n_orbs, 32, 0, st>>>(g, atab, orbs, astate, ss, n_orbs, work);
}

// FP64 FMA peak microbenchmark: 8 independent dependent-chains per thread, 2 flops per DFMA.
__global__ void dfma_peak_kernel(double* out, int iters)
{
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1., a2 = a0 + 2., a3 = a0 + 3., a4 = a0 + 4., a5 = a0 + 5., a6 = a0 + 6., a7 = a0 + 7.;
    const double m = 0.999999, c = 1e-7;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
        a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

void launch_dfma_peak(double* out, int blocks, int threads, int iters, cudaStream_t st)
{
    dfma_peak_kernel<<<blocks, threads, 0, st>>>(out, iters);
}

int search_rounds_needed(int Zmax)
{
    const double width = 50. + (double)Zmax * Zmax + 1.;
    // stage A: first round 33-section (both brackets equal), then 17-section; stage B: 32- then 33-section
    const int ra = 1 + (int)std::ceil(std::log(width / 33. / kEnergyTol) / std::log(17.)) + 1;
    const int rb = 1 + (int)std::ceil(std::log(width / 32. / kEnergyTol) / std::log(33.)) + 1;
    return ra + rb;
}

// ---------------------------------------------------------------------------------------------------------
// Two-sided solution (Numerov.h:403-504).  v1: one thread per orbital, serial in r, reference arithmetic.
// The outer part (i > match) is left unscaled; factor is returned in psi_scale and applied by the
// normalisation kernel (density_update) together with the y -> u conversion.
// ---------------------------------------------------------------------------------------------------------

__global__ void match_kernel(GridDev g, const double* __restrict__ atab_all, const OrbitalDev* orbs, const AtomState* astate,
                             SearchState* ss, double* psi_all, int* match_pt, int n_orbs)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_orbs) return;
    const OrbitalDev ob = orbs[k];
    if (astate[ob.atom].done) return;
    SearchState s = ss[k];
    if (s.stage != 3) {                 // search budget exhausted: didNotConverge (DFTAtom.cpp:516,538)
        s.converged = 0;
        s.E = (s.stage == 0) ? s.dn_hi : s.bot;
        s.stage = 3;
        ss[k] = s;
    }
    const double E = s.E;
    const double* atab = atab_all + (size_t)ob.tab * g.N;
    double* psi = psi_all + (size_t)k * g.N;
    const double ll1 = (double)(ob.l * (ob.l + 1));
    const double kappa = sqrt(2. * fabs(E));
    const int start = start_index(g, kappa);
    const int n_steps = g.N - 1;

    const MatchScale msc = match_scale(g, kappa, start, ob.l);
    double f12 = 0.;      // f_i of the node last evaluated by dval
    auto dval = [&](int i) { const double gq = match_g(g, atab, ll1, E, msc.rho2, i); f12 = 12. * gq; return 1. - gq; };

    for (int i = start + 1; i <= n_steps; ++i) psi[i] = 0.;
    double y = msc.y_s0;
    psi[start] = y;
    double d = dval(start);
    double wprev = d * y;
    y = msc.y_s1;
    psi[start - 1] = y;
    d = dval(start - 1);
    double w = d * y;
    double f = f12;
    double ynext = y;
    int match = 2;
    for (int i = start - 2; i > 0; --i) {
        const double wn = 2. * w - wprev + y * f;
        wprev = w; w = wn;
        d = dval(i);
        f = f12;
        ynext = y;
        y = w / d;
        psi[i] = y;
        if (y < ynext || fabs(y) > 1e15) { match = i; break; }
    }
    const double y_in_match = psi[match];
    // outward
    psi[0] = 0.;
    y = msc.y1;
    psi[1] = y;
    d = dval(1);
    f = f12;
    w = d * y; wprev = 0.;
    for (int i = 2; i < match; ++i) {
        const double wn = 2. * w - wprev + y * f;
        wprev = w; w = wn;
        d = dval(i);
        f = f12;
        y = w / d;
        psi[i] = y;
    }
    w = 2. * w - wprev + y * f;
    d = dval(match);
    y = w / d;
    const double factor = y / y_in_match;
    psi[match] = y;
    for (int i = match + 1; i <= start; ++i) psi[i] *= factor;
    match_pt[k] = match;
}

void launch_match(const GridDev& g, const double* atab, const OrbitalDev* orbs, const AtomState* astate, const SearchState* ss,
                  double* psi, int* match_pt, int n_orbs, cudaStream_t st)
{
    match_kernel<<<(n_orbs + 31) / 32, 32, 0, st>>>(g, atab, orbs, astate, const_cast<SearchState*>(ss), psi, match_pt, n_orbs);
}

__global__ void build_atab_kernel(GridDev g, const double* __restrict__ vpot, double* __restrict__ atab, int n_tabs)
{
    const size_t total = (size_t)n_tabs * g.N;
    const double q = g.delta * g.delta * 0.25;
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        const int i = (int)(t % g.N);
        // a_i = (2 K_i V_i + δ²/4) / 12 = f_i/12 at l = 0, E = 0 (Numerov.h:96-101), kept at full relative precision
        atab[t] = (g.k2[i] * vpot[t] + q) * (1. / 12.);
    }
}

void launch_build_atab(const GridDev& g, const double* vpot, double* atab, int n_tabs, cudaStream_t st)
{
    const size_t total = (size_t)n_tabs * g.N;
    const int blocks = (int)std::min<size_t>((total + 255) / 256, 148 * 8);
    build_atab_kernel<<<blocks, 256, 0, st>>>(g, vpot, atab, n_tabs);
}

}  // namespace dft
