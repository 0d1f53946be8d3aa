// Two-sided matched Numerov solution, parallel in r: one warp per orbital, its 32 lanes own 32 consecutive
// radial segments.  Replaces Numerov<...>::SolveSchrodingerMatchSolutionCompletely (reference DFTAtom/Numerov.h:403-504).
//
// Each sweep direction is done in two passes.  Pass 1: every lane pushes two basis states through its segment of
// the three-term recurrence, which gives the segment's 2x2 transfer matrix; the 32 matrices are applied in order
// along the warp (shuffles) to get each segment's true entry state.  Pass 2: every lane re-runs its segment from
// that entry state and stores y_i.
//
// The recurrence is the scaled difference form of numerov_fast.cu (g = f/12, d = 1 - g, s_i = 1 - d_i d_{i+1}):
//   inward   W_i = w_i prod_{j>i} d_j :   D_i = D_{i+1} + 10 g_i W_i + s_i W_{i+1},      W_{i-1} = W_i + D_i
//   outward  W_i = w_i prod_{j<i} d_j :   D_i = D_{i-1} + 10 g_i W_i + s_{i-1} W_{i-1},  W_{i+1} = W_i + D_i
// The state carried between segments is (W, D) rather than two consecutive W: the second component is small, so
// combining transfer matrices does not cancel leading digits.  One division per node, where y_i = w_i / d_i is stored.
#include "numerov_common.cuh"
#include <cstdio>

namespace dft {

struct Mat2 { double ww, wd, dw, dd; };      // (W', D') = (ww W + wd D, dw W + dd D)

__global__ void __launch_bounds__(128) match_seg_kernel(GridDev g, const double* __restrict__ atab_all, const OrbitalDev* orbs,
                                                        const AtomState* astate, SearchState* ss, double* psi_all, int* match_pt, int n_orbs)
{
    const unsigned full = 0xffffffffu;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int k = blockIdx.x * 4 + warp;
    if (k >= n_orbs) return;
    const OrbitalDev ob = orbs[k];
    if (astate[ob.atom].done) return;
    SearchState s = ss[k];
    if (s.stage != 3) {                 // search budget exhausted: didNotConverge (DFTAtom.cpp:516,538)
        s.converged = 0;
        s.E = (s.stage == 0) ? s.dn_hi : s.bot;
        s.stage = 3;
        if (lane == 0) ss[k] = s;
    }
    const double E = s.E;
    const double* __restrict__ atab = atab_all + (size_t)ob.tab * g.N;
    double* __restrict__ psi = psi_all + (size_t)k * g.N;
    const double ll1 = (double)(ob.l * (ob.l + 1));
    const double kappa = sqrt(2. * fabs(E));
    const int start = start_index_fast(g, kappa);
    const int N = g.N;
    const MatchScale msc = match_scale(g, kappa, start, ob.l);
    auto gval = [&](int i) { return match_g(g, atab, ll1, E, msc.rho2, i); };   // f_i / 12

    // zero tail, far seeds (Numerov.h:427-447)
    for (int i = start + 1 + lane; i < N; i += 32) psi[i] = 0.;
    const double y_s0 = msc.y_s0, y_s1 = msc.y_s1;
    const double g_s0 = gval(start), g_s1 = gval(start - 1);
    const double d_s0 = 1. - g_s0, d_s1 = 1. - g_s1;
    if (lane == 0) { psi[start] = y_s0; psi[start - 1] = y_s1; }

    // ------------------------------------------------------------------------------------------------
    // inward: nodes i = start-2 ... 1, lane s owns [bot, top] counted from the top
    // state entering a segment: (W_{top+1}, D_{top+2} = W_{top+1} - W_{top+2})
    // ------------------------------------------------------------------------------------------------
    int match = 2;
    double y_in_match = 0.;
    {
        const int n_in = start - 2;
        const int len = (n_in + 31) / 32;
        const int top = start - 2 - lane * len;
        const int bot = max(top - len + 1, 1);
        const bool have = top >= 1 && n_in > 0;
        Mat2 M = { 1., 0., 0., 1. };
        double prod = 1.;
        if (have) {
            double g1 = gval(top + 1), g2 = (top + 2 <= start) ? gval(top + 2) : 0.;
            // basis a: (W, D) = (1, 0) -> W_{top+1} = W_{top+2} = 1;   basis b: (0, 1) -> W_{top+1} = 0, W_{top+2} = -1
            double aW1 = 1., aW2 = 1., aD = 0., bW1 = 0., bW2 = -1., bD = 1.;
            for (int i = top; i >= bot; --i) {
                const double s1 = fma(-g1, g2, g1 + g2), t1 = 10. * g1;
                const double aDn = fma(t1, aW1, fma(s1, aW2, aD)), bDn = fma(t1, bW1, fma(s1, bW2, bD));
                aW2 = aW1; aW1 += aDn; aD = aDn;
                bW2 = bW1; bW1 += bDn; bD = bDn;
                prod *= (1. - g1);
                g2 = g1; g1 = gval(i);
            }
            M.ww = aW1; M.wd = bW1; M.dw = aD; M.dd = bD;      // out = (W_bot, D_{bot+1})
        }
        // entry of segment 0: W_{start-1} = d_{s1} y_{s1} d_{s0}, W_start = d_{s0} y_{s0}; P_{start-1} = d_{s0}
        const double Ws1 = d_s1 * y_s1 * d_s0, Ws0 = d_s0 * y_s0;
        double A = Ws1, B = Ws1 - Ws0, Pin = d_s0;
        for (int sgm = 0; sgm < 31; ++sgm) {
            const double oa = fma(M.ww, A, M.wd * B), ob_ = fma(M.dw, A, M.dd * B), op = Pin * prod;
            const double na = __shfl_sync(full, oa, sgm), nb = __shfl_sync(full, ob_, sgm), np = __shfl_sync(full, op, sgm);
            if (lane > sgm) { A = na; B = nb; Pin = np; }
        }
        // pass 2: y_i = W_i / (P_i d_i), P_i = P_{i+1} d_{i+1}; first node (descending) with y_i < y_{i+1} or |y_i| > 1e15
        int cand = 0;
        double ycand = 0., y2 = 0.;
        if (have) {
            double g1 = gval(top + 1), g2 = (top + 2 <= start) ? gval(top + 2) : 0.;
            double W1 = A, W2 = A - B, D = B, P = Pin;     // P = P_{top+1}
            double ynext = W1 / (P * (1. - g1));
            for (int i = top; i >= bot; --i) {
                const double s1 = fma(-g1, g2, g1 + g2), t1 = 10. * g1;
                const double Dn = fma(t1, W1, fma(s1, W2, D));
                const double W = W1 + Dn;
                P *= (1. - g1);
                const double gi = gval(i);
                const double y = W / (P * (1. - gi));
                psi[i] = y;
                if (!cand && (y < ynext || fabs(y) > 1e15)) { cand = i; ycand = y; }
                if (i == 2) y2 = y;
                ynext = y;
                W2 = W1; W1 = W; D = Dn; g2 = g1; g1 = gi;
            }
        }
        const unsigned mc = __ballot_sync(full, cand != 0);
        if (mc) {
            const int src = __ffs(mc) - 1;                 // segments are ordered from the top: lowest lane = first hit
            match = __shfl_sync(full, cand, src);
            y_in_match = __shfl_sync(full, ycand, src);
        } else {
            const unsigned m2 = __ballot_sync(full, have && bot <= 2 && top >= 2);
            const int src = m2 ? __ffs(m2) - 1 : 0;
            y_in_match = __shfl_sync(full, y2, src);       // matchPoint stays 2 (Numerov.h:449)
            if (!m2) y_in_match = (start - 1 == 2) ? y_s1 : y_s0;
        }
    }
    __syncwarp();

    // ------------------------------------------------------------------------------------------------
    // outward: y_0 = 0, y_1 = r_1^{l+1} e^{-δ/2} (Numerov.h:110-116, :470-477); nodes i = 2 ... match
    // state entering a segment: (W_{bot-1}, D_{bot-2} = W_{bot-1} - W_{bot-2});  Q_i = prod_{j<i} d_j
    // ------------------------------------------------------------------------------------------------
    double y_out_match;
    {
        const double y1 = msc.y1;
        const double gn1 = gval(1);
        const int n_out = match - 1;                       // nodes 2..match
        const int len = (n_out + 31) / 32;
        const int bot = 2 + lane * len;
        const int top = min(bot + len - 1, match);
        const bool have = bot <= match;
        Mat2 M = { 1., 0., 0., 1. };
        double prod = 1.;
        if (have) {
            double g1 = gval(bot - 1), g2 = (bot - 2 >= 1) ? gval(bot - 2) : 0.;      // g_{i-1}, g_{i-2}; d_0 := 1
            double aW1 = 1., aW2 = 1., aD = 0., bW1 = 0., bW2 = -1., bD = 1.;
            for (int i = bot; i <= top; ++i) {
                const double s1 = fma(-g1, g2, g1 + g2), t1 = 10. * g1;
                const double aDn = fma(t1, aW1, fma(s1, aW2, aD)), bDn = fma(t1, bW1, fma(s1, bW2, bD));
                aW2 = aW1; aW1 += aDn; aD = aDn;
                bW2 = bW1; bW1 += bDn; bD = bDn;
                prod *= (1. - g1);
                g2 = g1; g1 = gval(i);
            }
            M.ww = aW1; M.wd = bW1; M.dw = aD; M.dd = bD;
        }
        // entry of segment 0: W_1 = d_1 y_1 (Q_1 = 1), W_0 = 0  ->  (W, D) = (W_1, W_1)
        const double Wn1 = (1. - gn1) * y1;
        double A = Wn1, B = Wn1, Qin = 1.;
        for (int sgm = 0; sgm < 31; ++sgm) {
            const double oa = fma(M.ww, A, M.wd * B), ob_ = fma(M.dw, A, M.dd * B), oq = Qin * prod;
            const double na = __shfl_sync(full, oa, sgm), nb = __shfl_sync(full, ob_, sgm), nq = __shfl_sync(full, oq, sgm);
            if (lane > sgm) { A = na; B = nb; Qin = nq; }
        }
        double ylast = 0.;
        if (have) {
            double g1 = gval(bot - 1), g2 = (bot - 2 >= 1) ? gval(bot - 2) : 0.;
            double W1 = A, W2 = A - B, D = B, Q = Qin;     // Q = Q_{bot-1}
            for (int i = bot; i <= top; ++i) {
                const double s1 = fma(-g1, g2, g1 + g2), t1 = 10. * g1;
                const double Dn = fma(t1, W1, fma(s1, W2, D));
                const double W = W1 + Dn;
                Q *= (1. - g1);                            // Q_i = Q_{i-1} d_{i-1}
                const double gi = gval(i);
                const double y = W / (Q * (1. - gi));
                psi[i] = y;                                // includes psi[match] = outward value (Numerov.h:499)
                ylast = y;
                W2 = W1; W1 = W; D = Dn; g2 = g1; g1 = gi;
            }
        }
        const unsigned mm = __ballot_sync(full, have && top == match);
        y_out_match = __shfl_sync(full, ylast, __ffs(mm) - 1);
        if (lane == 0) { psi[0] = 0.; psi[1] = y1; }
    }
    __syncwarp();
    // scale the outer part so that both pieces meet at the match point (Numerov.h:497-501)
    const double factor = y_out_match / y_in_match;
    for (int i = match + 1 + lane; i <= start; i += 32) psi[i] *= factor;
    if (lane == 0) match_pt[k] = match;
}

// ---------------------------------------------------------------------------------------------------------
// The same two-sided matched solution by a whole CTA: kMT threads = kMT radial segments per direction (64 nodes each at
// 16385 nodes instead of 512), the 2x2 transfer matrices combined by a log-depth scan (Kogge-Stone in the warp, one
// shared-memory hop across warps).  Also returns 1 / integral u^2 dr (NormalizeNonUniform, DFTAtom.cpp:36-56), so the
// density kernel reads every orbital once.  This is the latency-optimised shape: ~30 us per orbital instead of ~500.
// ---------------------------------------------------------------------------------------------------------
constexpr int kMT = 256;

struct MatP { double ww, wd, dw, dd, p; };      // transfer matrix of (W, D) and the product of d over the segment

__device__ __forceinline__ MatP matp_mul(const MatP& later, const MatP& earlier)
{   // apply `earlier` first, then `later`
    MatP r;
    r.ww = fma(later.ww, earlier.ww, later.wd * earlier.dw);
    r.wd = fma(later.ww, earlier.wd, later.wd * earlier.dd);
    r.dw = fma(later.dw, earlier.ww, later.dd * earlier.dw);
    r.dd = fma(later.dw, earlier.wd, later.dd * earlier.dd);
    r.p = later.p * earlier.p;
    return r;
}
__device__ __forceinline__ MatP matp_shfl_up(const MatP& m, int o)
{
    const unsigned full = 0xffffffffu;
    MatP r;
    r.ww = __shfl_up_sync(full, m.ww, o); r.wd = __shfl_up_sync(full, m.wd, o);
    r.dw = __shfl_up_sync(full, m.dw, o); r.dd = __shfl_up_sync(full, m.dd, o);
    r.p = __shfl_up_sync(full, m.p, o);
    return r;
}

struct MatchShared {
    MatP wtot[kMT / 32];
    int cand[kMT / 32]; double ycand[kMT / 32];
    double red[kMT / 32];
    double y2, ylast, bcast;
    double xW, xD, xP;        // windowed kernel: state leaving a window
    int match;
};

// entry state of every thread's segment: (A, B, Pin) = [product of the maps of all earlier segments] applied to (A0, B0, P0);
// segments are ordered by thread index.  Block-collective.
__device__ __forceinline__ void segment_entries(const MatP& mine, double A0, double B0, double P0, MatchShared& sh,
                                                double& A, double& B, double& Pin)
{
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    MatP t = mine;                                   // inclusive scan: maps of lanes 0 .. lane, composed in order
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const MatP prevm = matp_shfl_up(t, o);
        if (lane >= o) t = matp_mul(t, prevm);
    }
    if (lane == 31) sh.wtot[w] = t;
    __syncthreads();
    double a = A0, b = B0, p = P0;                   // state entering this warp
    for (int v = 0; v < w; ++v) {
        const MatP m = sh.wtot[v];
        const double na = fma(m.ww, a, m.wd * b), nb = fma(m.dw, a, m.dd * b);
        a = na; b = nb; p *= m.p;
    }
    MatP ex = matp_shfl_up(t, 1);                    // exclusive prefix inside the warp
    if (lane == 0) { ex.ww = 1.; ex.wd = 0.; ex.dw = 0.; ex.dd = 1.; ex.p = 1.; }
    A = fma(ex.ww, a, ex.wd * b);
    B = fma(ex.dw, a, ex.dd * b);
    Pin = p * ex.p;
    __syncthreads();                                 // wtot may be reused
}

// 1 / x for the normalisation denominators P_i d_i of the hot loops (O(1) numbers: products of 1 - f/12): hardware reciprocal estimate
// (rcp.approx.ftz.f64, ~20 bits) + two Newton steps = ~1 ulp in 5 FP64 instructions, where the IEEE division is a ~30-instruction call that
// ptxas evaluates on every node.  The wave functions are compared at 1e-10; the match point is decided by sign tests that do not divide.
__device__ __forceinline__ double fast_rcp(double x)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    r = fma(fma(-x, r, 1.), r, r);
    r = fma(fma(-x, r, 1.), r, r);
    return r;
}

// Node i of the staged arrays lives at slot i + i / 32: a thread walking its own contiguous chunk and a warp reading 32
// consecutive nodes are both free of bank conflicts as long as the chunk length is a multiple of 32 (or the chunk stride
// in slots is odd); the launcher rounds the chunk length accordingly.
__device__ __forceinline__ int pslot(int i) { return i + (i >> 5); }

// SMEM = true: g_i = f_i / 12 of the whole orbital is staged in shared memory (one coalesced pass over the tables) and
// the solution is built in place of it, then copied out coalesced; per-thread chunk walks through global memory would
// touch 32 cache lines per warp instruction (measured: 800 cycles per node, L1-bound).  SMEM = false: the same algorithm
// on global memory (kept for comparison only: grids too large for one shared-memory stage take match_win_kernel below).
template <bool SMEM>
__global__ void __launch_bounds__(kMT) match_cta_kernel(GridDev g, const double* __restrict__ atab_all, const OrbitalDev* orbs,
                                                        const AtomState* astate, SearchState* ss, double* psi_all, int* match_pt,
                                                        double* inv_norm, int n_orbs, int step_min, int step_max)
{
    DFT_PDL_WAIT();
    __shared__ MatchShared sh;
    extern __shared__ double gy[];                       // SMEM: g_i, later y_i, slot pslot(i)
    const unsigned full = 0xffffffffu;
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    const int k = blockIdx.x;
    if (k >= n_orbs) return;
    const OrbitalDev ob = orbs[k];
    if (astate[ob.atom].done) return;
    { const int sc = astate[ob.atom].n_steps; if (sc < step_min || sc >= step_max) return; }     // (the two shapes of the kernel share an SCF by step index)
#ifdef DFT_MATCH_DEBUG
    long long tclk[8]; int nclk = 0;
#define MCLK() do { if (t == 0 && nclk < 8) tclk[nclk++] = clock64(); } while (0)
    MCLK();
#else
#define MCLK() do { } while (0)
#endif
    SearchState s = ss[k];
    if (s.stage != 3) {                 // search budget exhausted: didNotConverge (DFTAtom.cpp:516,538)
        s.converged = 0;
        s.E = (s.stage == 0) ? s.dn_hi : s.bot;
        s.stage = 3;
        if (t == 0) ss[k] = s;
    }
    const double E = s.E;
    const double* __restrict__ atab = atab_all + (size_t)ob.tab * g.N;
    double* __restrict__ psi = psi_all + (size_t)k * g.N;
    const double ll1 = (double)(ob.l * (ob.l + 1));
    const double kappa = sqrt(2. * fabs(E));
    const int start = start_index_fast(g, kappa);
    const int N = g.N;
    const MatchScale msc = match_scale(g, kappa, start, ob.l);
    auto gtab = [&](int i) { return match_g(g, atab, ll1, E, msc.rho2, i); };   // f_i / 12
    if (SMEM) {
        // batches of 16 nodes per thread: 48 independent table loads (~100 KB per CTA) in flight instead of one L2 round trip per node
        // (one CTA of 256 threads per SM: the register file is free)
        for (int i0 = t; i0 <= start; i0 += 16 * kMT) {
            double gv[16];
#pragma unroll
            for (int u = 0; u < 16; ++u) { const int i = i0 + u * kMT; gv[u] = gtab(min(i, start)); }
#pragma unroll
            for (int u = 0; u < 16; ++u) { const int i = i0 + u * kMT; if (i <= start) gy[pslot(i)] = gv[u]; }
        }
        __syncthreads();
    }
    auto gval = [&](int i) { return SMEM ? gy[pslot(i)] : gtab(i); };
    auto put = [&](int i, double y) { if (SMEM) gy[pslot(i)] = y; else psi[i] = y; };

    MCLK();
    // far seeds (Numerov.h:427-447)
    const double y_s0 = msc.y_s0, y_s1 = msc.y_s1;
    const double g_s0 = gval(start), g_s1 = gval(start - 1);
    const double d_s0 = 1. - g_s0, d_s1 = 1. - g_s1;
    if (t == 0) { sh.y2 = 0.; sh.ylast = 0.; }

    // ------------------------------------------------------------------------------------------------
    // inward: nodes i = start-2 ... 1, thread t owns [bot, top] counted from the top
    // state entering a segment: (W_{top+1}, D_{top+2} = W_{top+1} - W_{top+2})
    // ------------------------------------------------------------------------------------------------
    int match = 2;
    double y_in_match = 0.;
    {
        const int n_in = start - 2;
        const int len = (((n_in + kMT - 1) / kMT) + 31) & ~31;      // chunk length: a multiple of 32 (bank-conflict-free walks)
        const int top = start - 2 - t * len;
        const int bot = max(top - len + 1, 1);
        const bool have = top >= 1 && n_in > 0;
        MatP M = { 1., 0., 0., 1., 1. };
        double g1e = 0., g2e = 0.;                             // g_{top+1}, g_{top+2}: read before anything is overwritten
        if (have) { g1e = gval(top + 1); g2e = (top + 2 <= start) ? gval(top + 2) : 0.; }
        if (have) {
            double g1 = g1e, g2 = g2e;
            // basis a: (W, D) = (1, 0) -> W_{top+1} = W_{top+2} = 1;   basis b: (0, 1) -> W_{top+1} = 0, W_{top+2} = -1
            double aW1 = 1., aW2 = 1., aD = 0., bW1 = 0., bW2 = -1., bD = 1., prod = 1.;
            for (int i = top; i >= bot; --i) {
                const double s1 = fma(-g1, g2, g1 + g2), t1 = 10. * g1;
                const double aDn = fma(t1, aW1, fma(s1, aW2, aD)), bDn = fma(t1, bW1, fma(s1, bW2, bD));
                aW2 = aW1; aW1 += aDn; aD = aDn;
                bW2 = bW1; bW1 += bDn; bD = bDn;
                prod *= (1. - g1);
                g2 = g1; g1 = gval(i);
            }
            M.ww = aW1; M.wd = bW1; M.dw = aD; M.dd = bD; M.p = prod;      // out = (W_bot, D_{bot+1})
        }
        // entry of segment 0: W_{start-1} = d_{s1} y_{s1} d_{s0}, W_start = d_{s0} y_{s0}; P_{start-1} = d_{s0}
        const double Ws1 = d_s1 * y_s1 * d_s0, Ws0 = d_s0 * y_s0;
        double A, B, Pin;
        MCLK();
        segment_entries(M, Ws1, Ws1 - Ws0, d_s0, sh, A, B, Pin);
        MCLK();
        // pass 2a: y_i = W_i / (P_i d_i), P_i = P_{i+1} d_{i+1}; first node (descending) with y_i < y_{i+1} or |y_i| > 1e15.
        // Nothing is stored yet: the nodes below the match point still need their g for the outward solution.
        int cand = 0;
        double ycand = 0.;
        if (have) {
            double g1 = g1e, g2 = g2e;
            double W1 = A, W2 = A - B, D = B, P = Pin;     // P = P_{top+1}
            double Wc = 0., denc = 1., Wy2 = 0., deny2 = 1.;
            bool has2 = false;
            for (int i = top; i >= bot; --i) {
                const double s1 = fma(-g1, g2, g1 + g2), t1 = 10. * g1;
                const double Dn = fma(t1, W1, fma(s1, W2, D));
                const double W = W1 + Dn;
                P *= (1. - g1);                              // P_i = P_{i+1} d_{i+1}
                const double gi = gval(i);
                const double den = P * (1. - gi);            // y_i = W_i / den,  y_{i+1} = W_{i+1} / P_i
                // y_i < y_{i+1}  <=>  (W_i P_i - W_{i+1} den) sign(den P_i) < 0: no division inside the loop
                const bool drop = (W * P - W1 * den) * (den * P) < 0.;
                const bool big = fabs(W) > 1e15 * fabs(den);
                if (i == 2) { Wy2 = W; deny2 = den; has2 = true; }
                W2 = W1; W1 = W; D = Dn; g2 = g1; g1 = gi;
                if (drop || big) { cand = i; Wc = W; denc = den; break; }       // everything below belongs to the outward solution
            }
            if (cand) ycand = Wc / denc;
            if (has2) sh.y2 = Wy2 / deny2;
        }
        // the first candidate from the top = the candidate of the lowest thread index that has one
        const unsigned mc = __ballot_sync(full, cand != 0);
        const int srcl = mc ? __ffs(mc) - 1 : 0;
        const int wc = __shfl_sync(full, cand, srcl);
        const double wy = __shfl_sync(full, ycand, srcl);
        if (lane == 0) { sh.cand[w] = mc ? wc : 0; sh.ycand[w] = wy; }
        __syncthreads();
        bool found = false;
        for (int v = 0; v < kMT / 32; ++v)
            if (!found && sh.cand[v]) { match = sh.cand[v]; y_in_match = sh.ycand[v]; found = true; }
        if (!found) {                                      // matchPoint stays 2 (Numerov.h:449)
            y_in_match = (start - 2 >= 2) ? sh.y2 : ((start - 1 == 2) ? y_s1 : y_s0);
        }
        // pass 2b: the inward solution above the match point, stored (in place of g: every thread only overwrites nodes
        // whose g it has already consumed; the entry values g1e, g2e of the neighbours are in registers)
        __syncthreads();
        MCLK();
        if (have && top > match) {
            double g1 = g1e, g2 = g2e;
            double W1 = A, W2 = A - B, D = B;
            double P = Pin;
            for (int i = top; i >= bot && i > match; --i) {
                const double s1 = fma(-g1, g2, g1 + g2), t1 = 10. * g1;
                const double Dn = fma(t1, W1, fma(s1, W2, D));
                const double W = W1 + Dn;
                P *= (1. - g1);
                const double gi = gval(i);
                put(i, W * fast_rcp(P * (1. - gi)));
                W2 = W1; W1 = W; D = Dn; g2 = g1; g1 = gi;
            }
        }
    }
    __syncthreads();
    MCLK();

    // ------------------------------------------------------------------------------------------------
    // outward: y_0 = 0, y_1 = r_1^{l+1} e^{-δ/2} (Numerov.h:110-116, :470-477); nodes i = 2 ... match
    // state entering a segment: (W_{bot-1}, D_{bot-2} = W_{bot-1} - W_{bot-2});  Q_i = prod_{j<i} d_j
    // ------------------------------------------------------------------------------------------------
    double y_out_match;
    const double y1 = msc.y1;
    {
        const double gn1 = gval(1);
        const int n_out = match - 1;                       // nodes 2..match
        const int len = (((n_out + kMT - 1) / kMT) + 31) & ~31;
        const int bot = 2 + t * len;
        const int top = min(bot + len - 1, match);
        const bool have = bot <= match;
        MatP M = { 1., 0., 0., 1., 1. };
        double g1e = 0., g2e = 0.;                             // g_{bot-1}, g_{bot-2}; d_0 := 1
        if (have) { g1e = gval(bot - 1); g2e = (bot - 2 >= 1) ? gval(bot - 2) : 0.; }
        if (have) {
            double g1 = g1e, g2 = g2e;
            double aW1 = 1., aW2 = 1., aD = 0., bW1 = 0., bW2 = -1., bD = 1., prod = 1.;
            for (int i = bot; i <= top; ++i) {
                const double s1 = fma(-g1, g2, g1 + g2), t1 = 10. * g1;
                const double aDn = fma(t1, aW1, fma(s1, aW2, aD)), bDn = fma(t1, bW1, fma(s1, bW2, bD));
                aW2 = aW1; aW1 += aDn; aD = aDn;
                bW2 = bW1; bW1 += bDn; bD = bDn;
                prod *= (1. - g1);
                g2 = g1; g1 = gval(i);
            }
            M.ww = aW1; M.wd = bW1; M.dw = aD; M.dd = bD; M.p = prod;
        }
        // entry of segment 0: W_1 = d_1 y_1 (Q_1 = 1), W_0 = 0  ->  (W, D) = (W_1, W_1)
        const double Wn1 = (1. - gn1) * y1;
        double A, B, Qin;
        segment_entries(M, Wn1, Wn1, 1., sh, A, B, Qin);      // (its barriers also order the entry reads before the stores below)
        if (have) {
            double g1 = g1e, g2 = g2e;
            double W1 = A, W2 = A - B, D = B;
            double Q = Qin;                                // Q = Q_{bot-1}
            double ylast = 0.;
            for (int i = bot; i <= top; ++i) {
                const double s1 = fma(-g1, g2, g1 + g2), t1 = 10. * g1;
                const double Dn = fma(t1, W1, fma(s1, W2, D));
                const double W = W1 + Dn;
                Q *= (1. - g1);                            // Q_i = Q_{i-1} d_{i-1}
                const double gi = gval(i);
                const double y = W * fast_rcp(Q * (1. - gi));
                put(i, y);                                 // includes node `match` = outward value (Numerov.h:499)
                ylast = y;
                W2 = W1; W1 = W; D = Dn; g2 = g1; g1 = gi;
            }
            if (top == match) sh.ylast = ylast;
        }
        __syncthreads();
        y_out_match = sh.ylast;
    }
    // the seeds and the nodes 0, 1; then scale the outer part so that both pieces meet at the match point
    // (Numerov.h:497-501), zero the tail, and the norm integral of u^2 dr, u_i = y_i e^{i δ/2}, dr = Rp δ e^{δ i} di
    // (DFTAtom.cpp:36-56)
    if (t == 0) { put(start, y_s0); put(start - 1, y_s1); put(0, 0.); put(1, y1); }
    __syncthreads();
    MCLK();
    const double factor = y_out_match / y_in_match;
    double acc = 0.;
    for (int i0 = t; i0 < N; i0 += 16 * kMT) {              // batches of 16 nodes per thread: the table loads of a batch are independent
        double sq[16], wj[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) { const int i = min(i0 + u * kMT, N - 1); sq[u] = __ldg(g.sqex + i); wj[u] = __ldg(g.wjac + i); }
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            const int i = i0 + u * kMT;
            if (i < N) {
                double y = 0.;
                if (i <= start) {
                    y = SMEM ? gy[pslot(i)] : psi[i];
                    if (i > match) y *= factor;
                    const double uu = y * sq[u];
                    acc = fma(wj[u], uu * uu, acc);
                }
                if (SMEM || i > match) psi[i] = y;
            }
        }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(full, acc, o);
    if (lane == 0) sh.red[w] = acc;
    __syncthreads();
    if (t == 0) {
        double tot = 0.;
        for (int v = 0; v < kMT / 32; ++v) tot += sh.red[v];
        inv_norm[k] = 1. / tot;
        match_pt[k] = match;
    }
#ifdef DFT_MATCH_DEBUG
    MCLK();
    if (t == 0 && k == 0 && SMEM) printf("match clk (orb 0, start %d match %d): stage %lld in-pass1 %lld scan %lld 2a+cand %lld 2b %lld outward %lld final %lld\n", start, match,
                                  tclk[1] - tclk[0], tclk[2] - tclk[1], tclk[3] - tclk[2], tclk[4] - tclk[3], tclk[5] - tclk[4], tclk[6] - tclk[5], tclk[7] - tclk[6]);
#endif
}

// ---------------------------------------------------------------------------------------------------------
// The same algorithm for grids whose g = f/12 does not fit in shared memory (65537, 131073 nodes ...): the orbital is
// processed in windows of kWinNodes nodes, each staged in shared memory by coalesced loads, solved exactly like the
// single-window kernel above (256 segments, transfer matrices, log-depth scan) with the state leaving one window as the
// entry state of the next, and copied out coalesced.  Inward windows run from the far seeds down to the match point,
// outward windows from the nucleus up to it.  (Per-thread chunk walks through global memory - the old large-grid path -
// touch 32 cache lines per warp instruction: measured 1.15 ms per Rn orbital against ~0.2 ms here.)
// ---------------------------------------------------------------------------------------------------------
constexpr int kWinNodes = 24576;

__global__ void __launch_bounds__(kMT) match_win_kernel(GridDev g, const double* __restrict__ atab_all, const OrbitalDev* orbs,
                                                        const AtomState* astate, SearchState* ss, double* psi_all, int* match_pt,
                                                        double* inv_norm, int n_orbs, int win_nodes, int step_min, int step_max)
{
    DFT_PDL_WAIT();
    __shared__ MatchShared sh;
    extern __shared__ double gy[];                       // g_i of the window, later y_i; node i at pslot(i - base)
    const unsigned full = 0xffffffffu;
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    const int k = blockIdx.x;
    if (k >= n_orbs) return;
    const OrbitalDev ob = orbs[k];
    if (astate[ob.atom].done) return;
    { const int sc = astate[ob.atom].n_steps; if (sc < step_min || sc >= step_max) return; }
    SearchState s = ss[k];
    if (s.stage != 3) {                 // search budget exhausted: didNotConverge (DFTAtom.cpp:516,538)
        s.converged = 0;
        s.E = (s.stage == 0) ? s.dn_hi : s.bot;
        s.stage = 3;
        if (t == 0) ss[k] = s;
    }
    const double E = s.E;
    const double* __restrict__ atab = atab_all + (size_t)ob.tab * g.N;
    double* __restrict__ psi = psi_all + (size_t)k * g.N;
    const double ll1 = (double)(ob.l * (ob.l + 1));
    const double kappa = sqrt(2. * fabs(E));
    const int start = start_index_fast(g, kappa);
    const int N = g.N;
    const MatchScale msc = match_scale(g, kappa, start, ob.l);
    auto gtab = [&](int i) { return match_g(g, atab, ll1, E, msc.rho2, i); };   // f_i / 12

    // far seeds (Numerov.h:427-447)
    const double y_s0 = msc.y_s0, y_s1 = msc.y_s1;
    const double d_s0 = 1. - gtab(start), d_s1 = 1. - gtab(start - 1);
    if (t == 0) { sh.y2 = 0.; sh.ylast = 0.; }

    // ---------------- inward windows: nodes start-2 ... down to the match point ----------------
    int match = 2;
    double y_in_match = 0.;
    bool found = false;
    double eW = d_s1 * y_s1 * d_s0, eD = d_s1 * y_s1 * d_s0 - d_s0 * y_s0, eP = d_s0;      // (W_{hi+1}, W_{hi+1} - W_{hi+2}, P_{hi+1})
    for (int hi = start - 2; hi >= 1 && !found; hi -= win_nodes) {
        const int lo = max(hi - win_nodes + 1, 1);
        const int base = lo;
        const int sth = min(hi + 2, start);
        __syncthreads();
        for (int i0 = lo + t; i0 <= sth; i0 += 16 * kMT) {     // 16 nodes per thread in flight (see match_cta_kernel)
            double gv[16];
#pragma unroll
            for (int u = 0; u < 16; ++u) gv[u] = gtab(min(i0 + u * kMT, sth));
#pragma unroll
            for (int u = 0; u < 16; ++u) { const int i = i0 + u * kMT; if (i <= sth) gy[pslot(i - base)] = gv[u]; }
        }
        __syncthreads();
        auto gval = [&](int i) { return gy[pslot(i - base)]; };
        const int n_in = hi - lo + 1;
        const int len = (((n_in + kMT - 1) / kMT) + 31) & ~31;
        const int top = hi - t * len;
        const int bot = max(top - len + 1, lo);
        const bool have = top >= lo;
        MatP M = { 1., 0., 0., 1., 1. };
        double g1e = 0., g2e = 0.;
        if (have) { g1e = gval(top + 1); g2e = (top + 2 <= start) ? gval(top + 2) : 0.; }
        if (have) {
            double g1 = g1e, g2 = g2e;
            double aW1 = 1., aW2 = 1., aD = 0., bW1 = 0., bW2 = -1., bD = 1., prod = 1.;
            for (int i = top; i >= bot; --i) {
                const double s1 = fma(-g1, g2, g1 + g2), t1 = 10. * g1;
                const double aDn = fma(t1, aW1, fma(s1, aW2, aD)), bDn = fma(t1, bW1, fma(s1, bW2, bD));
                aW2 = aW1; aW1 += aDn; aD = aDn;
                bW2 = bW1; bW1 += bDn; bD = bDn;
                prod *= (1. - g1);
                g2 = g1; g1 = gval(i);
            }
            M.ww = aW1; M.wd = bW1; M.dw = aD; M.dd = bD; M.p = prod;
        }
        double A, B, Pin;
        segment_entries(M, eW, eD, eP, sh, A, B, Pin);
        // first node (descending) with y_i < y_{i+1} or |y_i| > 1e15 (Numerov.h:449-468)
        int cand = 0;
        double ycand = 0.;
        if (have) {
            double g1 = g1e, g2 = g2e;
            double W1 = A, W2 = A - B, D = B, P = Pin;
            double Wc = 0., denc = 1., Wy2 = 0., deny2 = 1.;
            bool has2 = false;
            for (int i = top; i >= bot; --i) {
                const double s1 = fma(-g1, g2, g1 + g2), t1 = 10. * g1;
                const double Dn = fma(t1, W1, fma(s1, W2, D));
                const double W = W1 + Dn;
                P *= (1. - g1);
                const double gi = gval(i);
                const double den = P * (1. - gi);            // y_i = W_i / den, y_{i+1} = W_{i+1} / P_i: compared without a division (match_cta_kernel)
                const bool drop = (W * P - W1 * den) * (den * P) < 0.;
                const bool big = fabs(W) > 1e15 * fabs(den);
                if (i == 2) { Wy2 = W; deny2 = den; has2 = true; }
                W2 = W1; W1 = W; D = Dn; g2 = g1; g1 = gi;
                if (drop || big) { cand = i; Wc = W; denc = den; break; }
            }
            if (cand) ycand = Wc / denc;
            if (has2) sh.y2 = Wy2 / deny2;
        }
        const unsigned mc = __ballot_sync(full, cand != 0);
        const int srcl = mc ? __ffs(mc) - 1 : 0;
        const int wc = __shfl_sync(full, cand, srcl);
        const double wy = __shfl_sync(full, ycand, srcl);
        if (lane == 0) { sh.cand[w] = mc ? wc : 0; sh.ycand[w] = wy; }
        __syncthreads();
        for (int v = 0; v < kMT / 32; ++v)
            if (!found && sh.cand[v]) { match = sh.cand[v]; y_in_match = sh.ycand[v]; found = true; }
        __syncthreads();
        // the inward solution above the match point, stored in place of g, and the state leaving the window
        const int stop = found ? match : 0;
        if (have && top > stop) {
            double g1 = g1e, g2 = g2e;
            double W1 = A, W2 = A - B, D = B, P = Pin;
            for (int i = top; i >= bot && i > stop; --i) {
                const double s1 = fma(-g1, g2, g1 + g2), t1 = 10. * g1;
                const double Dn = fma(t1, W1, fma(s1, W2, D));
                const double W = W1 + Dn;
                P *= (1. - g1);
                const double gi = gval(i);
                gy[pslot(i - base)] = W * fast_rcp(P * (1. - gi));
                W2 = W1; W1 = W; D = Dn; g2 = g1; g1 = gi;
            }
            if (bot == lo && !found) { sh.xW = W1; sh.xD = D; sh.xP = P; }
        }
        __syncthreads();
        for (int i = max(lo, stop + 1) + t; i <= hi; i += kMT) psi[i] = gy[pslot(i - base)];
        if (!found) { eW = sh.xW; eD = sh.xD; eP = sh.xP; }
    }
    if (!found) y_in_match = (start - 2 >= 2) ? sh.y2 : ((start - 1 == 2) ? y_s1 : y_s0);      // matchPoint stays 2 (Numerov.h:449)

    // ---------------- outward windows: nodes 2 ... match ----------------
    const double y1 = msc.y1;
    double oW = (1. - gtab(1)) * y1, oD = oW, oQ = 1.;          // (W_{lo-1}, W_{lo-1} - W_{lo-2}, Q_{lo-1}); W_0 = 0
    for (int lo = 2; lo <= match; lo += win_nodes) {
        const int hi = min(lo + win_nodes - 1, match);
        const int base = lo - 2;
        __syncthreads();
        for (int i0 = max(lo - 2, 1) + t; i0 <= hi; i0 += 16 * kMT) {
            double gv[16];
#pragma unroll
            for (int u = 0; u < 16; ++u) gv[u] = gtab(min(i0 + u * kMT, hi));
#pragma unroll
            for (int u = 0; u < 16; ++u) { const int i = i0 + u * kMT; if (i <= hi) gy[pslot(i - base)] = gv[u]; }
        }
        __syncthreads();
        auto gval = [&](int i) { return gy[pslot(i - base)]; };
        const int n_out = hi - lo + 1;
        const int len = (((n_out + kMT - 1) / kMT) + 31) & ~31;
        const int bot = lo + t * len;
        const int top = min(bot + len - 1, hi);
        const bool have = bot <= hi;
        MatP M = { 1., 0., 0., 1., 1. };
        double g1e = 0., g2e = 0.;                             // g_{bot-1}, g_{bot-2}; d_0 := 1
        if (have) { g1e = gval(bot - 1); g2e = (bot - 2 >= 1) ? gval(bot - 2) : 0.; }
        if (have) {
            double g1 = g1e, g2 = g2e;
            double aW1 = 1., aW2 = 1., aD = 0., bW1 = 0., bW2 = -1., bD = 1., prod = 1.;
            for (int i = bot; i <= top; ++i) {
                const double s1 = fma(-g1, g2, g1 + g2), t1 = 10. * g1;
                const double aDn = fma(t1, aW1, fma(s1, aW2, aD)), bDn = fma(t1, bW1, fma(s1, bW2, bD));
                aW2 = aW1; aW1 += aDn; aD = aDn;
                bW2 = bW1; bW1 += bDn; bD = bDn;
                prod *= (1. - g1);
                g2 = g1; g1 = gval(i);
            }
            M.ww = aW1; M.wd = bW1; M.dw = aD; M.dd = bD; M.p = prod;
        }
        double A, B, Qin;
        segment_entries(M, oW, oD, oQ, sh, A, B, Qin);
        if (have) {
            double g1 = g1e, g2 = g2e;
            double W1 = A, W2 = A - B, D = B, Q = Qin;
            double ylast = 0.;
            for (int i = bot; i <= top; ++i) {
                const double s1 = fma(-g1, g2, g1 + g2), t1 = 10. * g1;
                const double Dn = fma(t1, W1, fma(s1, W2, D));
                const double W = W1 + Dn;
                Q *= (1. - g1);
                const double gi = gval(i);
                const double y = W * fast_rcp(Q * (1. - gi));
                gy[pslot(i - base)] = y;
                ylast = y;
                W2 = W1; W1 = W; D = Dn; g2 = g1; g1 = gi;
            }
            if (top == hi) { sh.xW = W1; sh.xD = D; sh.xP = Q; sh.ylast = ylast; }
        }
        __syncthreads();
        for (int i = lo + t; i <= hi; i += kMT) psi[i] = gy[pslot(i - base)];
        oW = sh.xW; oD = sh.xD; oQ = sh.xP;
    }
    __syncthreads();
    const double y_out_match = sh.ylast;
    if (t == 0) { psi[start] = y_s0; psi[start - 1] = y_s1; psi[0] = 0.; psi[1] = y1; }
    __syncthreads();
    // scale the outer part so that both pieces meet at the match point (Numerov.h:497-501), zero the tail, norm integral
    const double factor = y_out_match / y_in_match;
    double acc = 0.;
    for (int i0 = t; i0 < N; i0 += 16 * kMT) {               // 16 nodes per thread in flight
        double sq[16], wj[16], yv[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) { const int i = min(i0 + u * kMT, N - 1); sq[u] = __ldg(g.sqex + i); wj[u] = __ldg(g.wjac + i); yv[u] = (i <= start) ? psi[i] : 0.; }
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            const int i = i0 + u * kMT;
            if (i < N) {
                double y = 0.;
                if (i <= start) {
                    y = yv[u];
                    if (i > match) y *= factor;
                    const double uu = y * sq[u];
                    acc = fma(wj[u], uu * uu, acc);
                }
                if (i > match) psi[i] = y;
            }
        }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(full, acc, o);
    if (lane == 0) sh.red[w] = acc;
    __syncthreads();
    if (t == 0) {
        double tot = 0.;
        for (int v = 0; v < kMT / 32; ++v) tot += sh.red[v];
        inv_norm[k] = 1. / tot;
        match_pt[k] = match;
    }
}

// Opt-in to large dynamic shared memory is per-device state: set once per device from dftatom_create (under cudaSetDevice), not
// cached in a process-wide static (a second context on another GPU would otherwise never get it).
constexpr size_t kMatchCtaMaxBytes = 200 * 1024;
int match_init_device()
{
    const size_t wb = ((size_t)kWinNodes + 2 + (size_t)(kWinNodes + 2) / 32 + 8) * sizeof(double);
    DFT_CHECK(cudaFuncSetAttribute(match_cta_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMatchCtaMaxBytes));
    DFT_CHECK(cudaFuncSetAttribute(match_win_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wb));
    return 0;
}

static size_t match_win_bytes(int win_nodes) { return ((size_t)win_nodes + 2 + (size_t)(win_nodes + 2) / 32 + 8) * sizeof(double); }

// win_until_step > 0 (grids that fit one window): while an atom's SCF step counter is below it, its orbitals are solved by the windowed kernel
// with small windows (several CTAs per SM: throughput while most atoms are still iterating), afterwards by the one-window kernel (latency);
// both are launched, the atom's own step counter decides - its records do not depend on what else is in the batch
int launch_match_cta(const GridDev& g, const double* atab, const OrbitalDev* orbs, const AtomState* astate, const SearchState* ss,
                     double* psi, int* match_pt, double* inv_norm, int n_orbs, int win_until_step, int win_nodes, int step_lo, int step_hi, cudaStream_t st)
{
    int n_launch = 0;
    // [step_lo, step_hi): the SCF steps this launch can be executed at; a shape whose step window misses it is not launched
    const size_t bytes = ((size_t)g.N + (size_t)g.N / 32 + 8) * sizeof(double);
    if (bytes <= kMatchCtaMaxBytes) {
        int lo = 0;
        if (win_until_step > 0 && win_nodes >= 1024 && win_nodes < g.N) {
            if (step_lo < win_until_step) {
                launch_step_kernel(match_win_kernel, dim3(n_orbs), dim3(kMT), match_win_bytes(win_nodes), st, g, atab, orbs, astate, const_cast<SearchState*>(ss), psi, match_pt,
                                   inv_norm, n_orbs, win_nodes, 0, win_until_step);
                ++n_launch;
            }
            lo = win_until_step;
        }
        if (step_hi > lo) {
            launch_step_kernel(match_cta_kernel<true>, dim3(n_orbs), dim3(kMT), bytes, st, g, atab, orbs, astate, const_cast<SearchState*>(ss), psi, match_pt, inv_norm, n_orbs, lo, 1 << 30);
            ++n_launch;
        }
    } else {
        launch_step_kernel(match_win_kernel, dim3(n_orbs), dim3(kMT), match_win_bytes(kWinNodes), st, g, atab, orbs, astate, const_cast<SearchState*>(ss), psi, match_pt, inv_norm,
                           n_orbs, kWinNodes, 0, 1 << 30);
        ++n_launch;
    }
    return n_launch;
}

void launch_match_seg(const GridDev& g, const double* atab, const OrbitalDev* orbs, const AtomState* astate, const SearchState* ss,
                      double* psi, int* match_pt, int n_orbs, cudaStream_t st)
{
    match_seg_kernel<<<(n_orbs + 3) / 4, 128, 0, st>>>(g, atab, orbs, astate, const_cast<SearchState*>(ss), psi, match_pt, n_orbs);
}

}  // namespace dft
