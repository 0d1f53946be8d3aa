// Two-sided matched Numerov solution, parallel in r: one warp per orbital, its 32 lanes own 32 consecutive
// radial segments.  Replaces Numerov<...>::SolveSchrodingerMatchSolutionCompletely (reference DFTAtom/Numerov.h:403-504).
//
// Each sweep direction is done in two passes.  Pass 1: every lane pushes the two unit vectors through its
// segment of the three-term recurrence, which gives the segment's 2x2 transfer matrix; the 32 matrices are
// applied in order along the warp (shuffles) to get each segment's true entry vector.  Pass 2: every lane
// re-runs its segment from that entry vector and stores y_i.  The recurrence is the division-free scaled form
// (see numerov_fast.cu): inward  W_i = w_i prod_{j>i} d_j,  outward  W_i = w_i prod_{j<i} d_j,
// with one division per node only where y_i = w_i / d_i is written out.
#include "numerov_common.cuh"

namespace dft {

struct Mat2 { double a, b, c, d; };      // (x', y') = (a x + b y, c x + d y)

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
    const double nll1 = -(double)(ob.l * (ob.l + 1));
    const double kappa = sqrt(2. * fabs(E));
    const int start = start_index(g, kappa);
    const int N = g.N;
    auto dval = [&](int i) { return fma(E, __ldg(g.c6 + i), fma(nll1, __ldg(g.b12 + i), __ldg(atab + i))); };

    // zero tail, far seeds (Numerov.h:427-447)
    for (int i = start + 1 + lane; i < N; i += 32) psi[i] = 0.;
    const double y_s0 = far_value(g, kappa, start), y_s1 = far_value(g, kappa, start - 1);
    const double d_s0 = dval(start), d_s1 = dval(start - 1);
    if (lane == 0) { psi[start] = y_s0; psi[start - 1] = y_s1; }

    // ------------------------------------------------------------------------------------------------
    // inward: nodes i = start-2 ... 1, lane s owns [bot, top] counted from the top
    // ------------------------------------------------------------------------------------------------
    int match = 2;
    double y_in_match = 0.;
    {
        const int n_in = start - 2;
        const int len = (n_in + 31) / 32;
        const int top = start - 2 - lane * len;
        const int bot = max(top - len + 1, 1);
        const bool have = top >= 1 && n_in > 0;
        // pass 1: transfer matrix of (W_{top+1}, W_{top+2}) -> (W_bot, W_{bot+1}); prod = product of d_{i+1}, i in segment
        Mat2 M = { 1., 0., 0., 1. };
        double prod = 1.;
        if (have) {
            double d1 = dval(top + 1), d2 = (top + 2 <= start) ? dval(top + 2) : 1.;
            double u1 = 1., u2 = 0., v1 = 0., v2 = 1.;     // u: W_{i+1} column, v: W_{i+2} column (as coefficients of the entry vector)
            for (int i = top; i >= bot; --i) {
                const double n1 = fma(-10., d1, 12.), dd = d1 * d2;
                const double un = fma(n1, u1, -(dd * u2)), vn = fma(n1, v1, -(dd * v2));
                u2 = u1; u1 = un; v2 = v1; v1 = vn;
                prod *= d1;
                d2 = d1; d1 = dval(i);
            }
            M.a = u1; M.b = v1; M.c = u2; M.d = v2;
        }
        // entry vectors: segment 0 enters with (W_{start-1}, W_start) = (d_{s1} y_{s1} d_{s0}, d_{s0} y_{s0}), P_{start-1} = d_{s0}
        double A = d_s1 * y_s1 * d_s0, B = d_s0 * y_s0, Pin = d_s0;
        for (int sgm = 0; sgm < 31; ++sgm) {
            // lane sgm holds the entry of segment sgm; produce the entry of segment sgm+1
            const double oa = fma(M.a, A, M.b * B), ob_ = fma(M.c, A, M.d * B), op = Pin * prod;
            const double na = __shfl_sync(full, oa, sgm), nb = __shfl_sync(full, ob_, sgm), np = __shfl_sync(full, op, sgm);
            if (lane > sgm) { A = na; B = nb; Pin = np; }
        }
        // pass 2: y_i = W_i / (P_i d_i), P_i = P_{i+1} d_{i+1}; first node (descending) with y_i < y_{i+1} or |y_i| > 1e15
        int cand = 0;
        double ycand = 0., y2 = 0.;
        if (have) {
            double d1 = dval(top + 1), d2 = (top + 2 <= start) ? dval(top + 2) : 1.;
            double W1 = A, W2 = B, P = Pin;                // P = P_{top+1}
            double ynext = W1 / (P * d1);
            for (int i = top; i >= bot; --i) {
                const double n1 = fma(-10., d1, 12.), dd = d1 * d2;
                const double W = fma(n1, W1, -(dd * W2));
                P *= d1;
                const double d = dval(i);
                const double y = W / (P * d);
                psi[i] = y;
                if (!cand && (y < ynext || fabs(y) > 1e15)) { cand = i; ycand = y; }
                if (i == 2) y2 = y;
                ynext = y;
                W2 = W1; W1 = W; d2 = d1; d1 = d;
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
    // ------------------------------------------------------------------------------------------------
    double y_out_match;
    {
        const double y1 = pow(__ldg(g.r + 1), (double)ob.l + 1.) * exp(-0.5 * g.delta);
        const double dn1 = dval(1);
        const int n_out = match - 1;                       // nodes 2..match
        const int len = (n_out + 31) / 32;
        const int bot = 2 + lane * len;
        const int top = min(bot + len - 1, match);
        const bool have = bot <= match;
        Mat2 M = { 1., 0., 0., 1. };
        double prod = 1.;
        if (have) {
            double d1 = dval(bot - 1), d2 = (bot - 2 >= 1) ? dval(bot - 2) : 1.;      // d_{i-1}, d_{i-2}; d_0 := 1
            double u1 = 1., u2 = 0., v1 = 0., v2 = 1.;
            for (int i = bot; i <= top; ++i) {
                const double n1 = fma(-10., d1, 12.), dd = d1 * d2;
                const double un = fma(n1, u1, -(dd * u2)), vn = fma(n1, v1, -(dd * v2));
                u2 = u1; u1 = un; v2 = v1; v1 = vn;
                prod *= d1;
                d2 = d1; d1 = dval(i);
            }
            M.a = u1; M.b = v1; M.c = u2; M.d = v2;
        }
        // entry of segment 0: (W_1, W_0) = (d_1 y_1, 0), Q_1 = 1
        double A = dn1 * y1, B = 0., Qin = 1.;
        for (int sgm = 0; sgm < 31; ++sgm) {
            const double oa = fma(M.a, A, M.b * B), ob_ = fma(M.c, A, M.d * B), oq = Qin * prod;
            const double na = __shfl_sync(full, oa, sgm), nb = __shfl_sync(full, ob_, sgm), nq = __shfl_sync(full, oq, sgm);
            if (lane > sgm) { A = na; B = nb; Qin = nq; }
        }
        double ylast = 0.;
        if (have) {
            double d1 = dval(bot - 1), d2 = (bot - 2 >= 1) ? dval(bot - 2) : 1.;
            double W1 = A, W2 = B, Q = Qin;                // Q = Q_{bot-1}
            for (int i = bot; i <= top; ++i) {
                const double n1 = fma(-10., d1, 12.), dd = d1 * d2;
                const double W = fma(n1, W1, -(dd * W2));
                Q *= d1;                                   // Q_i = Q_{i-1} d_{i-1}
                const double d = dval(i);
                const double y = W / (Q * d);
                psi[i] = y;                                // includes psi[match] = outward value (Numerov.h:499)
                ylast = y;
                W2 = W1; W1 = W; d2 = d1; d1 = d;
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

void launch_match_seg(const GridDev& g, const double* atab, const OrbitalDev* orbs, const AtomState* astate, const SearchState* ss,
                      double* psi, int* match_pt, int n_orbs, cudaStream_t st)
{
    match_seg_kernel<<<(n_orbs + 3) / 4, 128, 0, st>>>(g, atab, orbs, astate, const_cast<SearchState*>(ss), psi, match_pt, n_orbs);
}

}  // namespace dft
