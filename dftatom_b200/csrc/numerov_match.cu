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
    const int start = start_index(g, kappa);
    const int N = g.N;
    auto gval = [&](int i) { return fma(-E, __ldg(g.c6 + i), fma(ll1, __ldg(g.b12 + i), __ldg(atab + i))); };   // f_i / 12

    // zero tail, far seeds (Numerov.h:427-447)
    for (int i = start + 1 + lane; i < N; i += 32) psi[i] = 0.;
    const double y_s0 = far_value(g, kappa, start), y_s1 = far_value(g, kappa, start - 1);
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
        const double y1 = pow(__ldg(g.r + 1), (double)ob.l + 1.) * exp(-0.5 * g.delta);
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

void launch_match_seg(const GridDev& g, const double* atab, const OrbitalDev* orbs, const AtomState* astate, const SearchState* ss,
                      double* psi, int* match_pt, int n_orbs, cudaStream_t st)
{
    match_seg_kernel<<<(n_orbs + 3) / 4, 128, 0, st>>>(g, atab, orbs, astate, const_cast<SearchState*>(ss), psi, match_pt, n_orbs);
}

}  // namespace dft
