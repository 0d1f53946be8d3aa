// Radial Poisson multigrid, one CTA per density, all levels of the 2^L+1 hierarchy L2-resident.
//
// Replaces (reference DFTAtom/) PoissonSolver.h:51-81 SolvePoissonNonUniform, :89-124 FullCycle, :155-159 VCycle and
// PoissonSolver.cpp:40-64 GaussSeidel, :66-77 IterateGaussSeidel, :80-106 Initialize, :110-123 Prolong, :126-157
// Restrict, :162-197 Ascend/Descend.
//
// The reference's smoother is a lexicographic in-place Gauss-Seidel sweep, i.e. the first-order recurrence
//     Phi_i <- a Phi_{i-1} + c_i ,   a = (1 + d_l/2)/2 ,   c_i = (S_i + (1 - d_l/2) Phi_{i+1}^old)/2 ,  d_l = δ 2^l .
// Every thread owns NPT consecutive nodes: it runs the recurrence locally with zero carry-in, the carries are
// resolved by an associative scan of the affine maps x -> a^m x + p (warp shuffles, then one shared-memory hop
// across warps), and the result is patched in.  That is the same sweep (same operator, same ordering), evaluated
// in O(NPT + log T) depth instead of O(N).
//
// Level visits keep a thread's nodes of Phi and Source in registers for the three sweeps (levels up to 16384 nodes);
// the six coarsest levels (<= 32 nodes) are run by warp 0 alone with no block barrier.  The update norm of the
// reference's IterateGaussSeidel only drives its early exits; with a fixed number of sweeps it is evaluated once,
// for the last fine-grid sweep of the solve.
#include "internal.h"
#include <cmath>

namespace dft {

PoissonLevels make_levels(int L)
{
    PoissonLevels lv{};
    lv.L = L;
    int off = 0;
    for (int l = 0; l < L; ++l) {
        lv.size[l] = (1 << (L - l)) + 1;
        lv.off[l] = off;
        off += (lv.size[l] + 3) & ~3;      // keep every level 32-byte aligned
    }
    lv.total = off;
    return lv;
}

constexpr int kPT = 512;     // threads per CTA
constexpr int kPM = 16;      // nodes per thread per pass of the generic (global-memory) sweep
constexpr int kMaxNpt = 32;  // register-resident visits handle levels with up to kPT * kMaxNpt nodes
constexpr int kWarpLevelNodes = 32;

struct PoissonSmem {
    double scanA[32], scanP[32];
    double edge[32];
    double red[32];
    double carry;        // last new value of the previous pass
    double bcast;
    unsigned long long updates;   // Gauss-Seidel node-updates performed by this CTA (work counter)
    int soff[24];                 // offsets of the shared-memory-resident levels inside the dynamic smem arrays
};

__device__ __forceinline__ double block_sum(double v, PoissonSmem& sm)
{
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if (lane == 0) sm.red[w] = v;
    __syncthreads();
    if (w == 0) {
        double t = (lane < (blockDim.x >> 5)) ? sm.red[lane] : 0.;
#pragma unroll
        for (int o = 16; o; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (lane == 0) sm.bcast = t;
    }
    __syncthreads();
    return sm.bcast;
}

// ---------------------------------------------------------------------------------------------------------
// generic sweep on global-memory level arrays (any size, multi-pass); returns sqrt(sum (old-new)^2)
// ---------------------------------------------------------------------------------------------------------
__device__ __noinline__ double gs_sweep(double* __restrict__ phi, const double* __restrict__ src, int size, double d, PoissonSmem& sm)
{
    const int T = blockDim.x, t = threadIdx.x, lane = t & 31, w = t >> 5;
    const double a = 0.5 * (1. + 0.5 * d), bcoef = 0.5 * (1. - 0.5 * d);
    const int n_int = size - 2;
    double err2 = 0.;
    if (t == 0) { sm.carry = phi[0]; sm.updates += (unsigned long long)n_int; }
    const int per_pass = T * kPM;
    for (int base = 0; base < n_int; base += per_pass) {
        const int i0 = 1 + base + t * kPM;
        int m = n_int + 1 - i0;                 // valid nodes of this thread in this pass
        m = m < 0 ? 0 : (m > kPM ? kPM : m);
        double p[kPM];
        double A = 1., x = 0.;
        if (m > 0) {
#pragma unroll
            for (int k = 0; k < kPM; ++k) {
                if (k < m) {
                    const double c = fma(bcoef, phi[i0 + k + 1], 0.5 * src[i0 + k]);
                    x = fma(a, x, c);
                    p[k] = x;
                    A *= a;
                }
            }
        }
        double sA = A, sP = x;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const double pa = __shfl_up_sync(0xffffffffu, sA, o), pp = __shfl_up_sync(0xffffffffu, sP, o);
            if (lane >= o) { sP = fma(sA, pp, sP); sA *= pa; }
        }
        __syncthreads();                           // all loads of old values done; smem from previous pass consumed
        if (lane == 31) { sm.scanA[w] = sA; sm.scanP[w] = sP; }
        __syncthreads();
        if (w == 0) {
            const int nw = T >> 5;
            double wa = lane < nw ? sm.scanA[lane] : 1., wp = lane < nw ? sm.scanP[lane] : 0.;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const double pa = __shfl_up_sync(0xffffffffu, wa, o), pp = __shfl_up_sync(0xffffffffu, wp, o);
                if (lane >= o) { wp = fma(wa, pp, wp); wa *= pa; }
            }
            sm.scanA[lane] = wa; sm.scanP[lane] = wp;   // inclusive over warps
        }
        __syncthreads();
        double eA = __shfl_up_sync(0xffffffffu, sA, 1), eP = __shfl_up_sync(0xffffffffu, sP, 1);
        if (lane == 0) { eA = 1.; eP = 0.; }
        if (w > 0) { const double wa = sm.scanA[w - 1], wp = sm.scanP[w - 1]; eP = fma(eA, wp, eP); eA *= wa; }
        const double left = sm.carry;
        const double cin = fma(eA, left, eP);       // new value of node i0-1
        if (m > 0) {
            double q = a;
#pragma unroll
            for (int k = 0; k < kPM; ++k) {
                if (k < m) {
                    const double v = fma(q, cin, p[k]);
                    const double dif = phi[i0 + k] - v;
                    err2 = fma(dif, dif, err2);
                    phi[i0 + k] = v;
                    q *= a;
                }
            }
        }
        __syncthreads();
        if (t == T - 1) sm.carry = fma(sm.scanA[(T >> 5) - 1], left, sm.scanP[(T >> 5) - 1]);
        __syncthreads();
    }
    return sqrt(block_sum(err2, sm));
}

// ---------------------------------------------------------------------------------------------------------
// register-resident level visit: `sweeps` Gauss-Seidel sweeps on a level with n = size-1 <= kPT*NPT owned nodes
// (thread t owns nodes [t NPT, (t+1) NPT); node 0 is the fixed left boundary, node n the fixed right boundary)
// ---------------------------------------------------------------------------------------------------------
template <int NPT, bool SRC_REGS>
__device__ __noinline__ void visit_regs(double* __restrict__ phi_g, const double* __restrict__ src_g, int size, double d, int sweeps,
                                           int pad, PoissonSmem& sm)
{
    // pad = 1: the level lives in shared memory with one padding slot per thread chunk (chunk stride NPT + 1 doubles is odd,
    // so the 32 lanes of a warp hit 32 different bank pairs); pad = 0: plain layout in global memory
    const unsigned full = 0xffffffffu;
    const int t = threadIdx.x, lane = t & 31, w = t >> 5, nw = blockDim.x >> 5;
    const int n = size - 1;
    const int i0 = t * NPT;
    const int p0 = t * (NPT + pad);
    const bool active = i0 < n;
    const double a = 0.5 * (1. + 0.5 * d), bcoef = 0.5 * (1. - 0.5 * d);
    const double right_bc = phi_g[n + (pad ? n / NPT : 0)];
    double phi[NPT], src[SRC_REGS ? NPT : 1];
    const double* __restrict__ sp = src_g + (active ? p0 : 0);      // Source is read-only during the visit
#pragma unroll
    for (int k = 0; k < NPT; ++k) {
        phi[k] = active ? phi_g[p0 + k] : 0.;
        if (SRC_REGS) src[k] = active ? 0.5 * src_g[p0 + k] : 0.;
    }
    double aN = a;
#pragma unroll
    for (int k = 1; k < NPT; ++k) aN *= a;
    if (t == 0) sm.updates += (unsigned long long)sweeps * (unsigned long long)(n - 1);

    for (int sw = 0; sw < sweeps; ++sw) {
        // old value of the right neighbour's first node
        double nb = __shfl_down_sync(full, phi[0], 1);
        if (lane == 0) sm.edge[w] = phi[0];
        __syncthreads();
        if (lane == 31) nb = (w + 1 < nw) ? sm.edge[w + 1] : right_bc;
        if (i0 + NPT >= n) nb = right_bc;
        // local recurrence with zero carry-in, in place
        double x = 0.;
#pragma unroll
        for (int k = 0; k < NPT; ++k) {
            const double nxt = (k + 1 < NPT) ? phi[k + 1] : nb;
            const double c = fma(bcoef, nxt, SRC_REGS ? src[k] : 0.5 * sp[k]);
            x = (t == 0 && k == 0) ? phi[0] : fma(a, x, c);      // node 0 keeps its boundary value
            phi[k] = x;
        }
        double sA = active ? (t == 0 ? 0. : aN) : 1., sP = active ? x : 0.;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const double pa = __shfl_up_sync(full, sA, o), pp = __shfl_up_sync(full, sP, o);
            if (lane >= o) { sP = fma(sA, pp, sP); sA *= pa; }
        }
        if (lane == 31) { sm.scanA[w] = sA; sm.scanP[w] = sP; }
        __syncthreads();
        // carry into this warp = composition of the total maps of the warps before it, applied to 0: every warp scans
        // the (<= 32) warp totals itself with shuffles (no second barrier, no serial chain)
        double carry;
        {
            double wa = lane < nw ? sm.scanA[lane] : 1., wp = lane < nw ? sm.scanP[lane] : 0.;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const double pa = __shfl_up_sync(full, wa, o), pp = __shfl_up_sync(full, wp, o);
                if (lane >= o) { wp = fma(wa, pp, wp); wa *= pa; }
            }
            carry = __shfl_sync(full, wp, (w + 31) & 31);       // inclusive prefix of warp w-1 (map applied to 0 = its P)
            if (w == 0) carry = 0.;
        }
        double eA = __shfl_up_sync(full, sA, 1), eP = __shfl_up_sync(full, sP, 1);
        if (lane == 0) { eA = 1.; eP = 0.; }
        const double cin = fma(eA, carry, eP);                    // new value of node i0-1
        double q = a;
#pragma unroll
        for (int k = 0; k < NPT; ++k) { phi[k] = fma(q, cin, phi[k]); q *= a; }
    }
    if (active) {
#pragma unroll
        for (int k = 0; k < NPT; ++k) phi_g[p0 + k] = phi[k];
    }
    __syncthreads();
}

// same for a level of <= 32 owned nodes, executed by one warp, no block barrier
__device__ __noinline__ void visit_warp(double* __restrict__ phi_g, const double* __restrict__ src_g, int size, double d, int sweeps)
{
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int n = size - 1;
    const bool active = lane < n;
    const double a = 0.5 * (1. + 0.5 * d), bcoef = 0.5 * (1. - 0.5 * d);
    const double right_bc = phi_g[n];
    double phi = active ? phi_g[lane] : 0.;
    const double src = active ? 0.5 * src_g[lane] : 0.;
    for (int sw = 0; sw < sweeps; ++sw) {
        double nb = __shfl_down_sync(full, phi, 1);
        if (lane + 1 >= n) nb = right_bc;
        const double c = fma(bcoef, nb, src);
        double sA = active ? (lane == 0 ? 0. : a) : 1., sP = active ? (lane == 0 ? phi : c) : 0.;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const double pa = __shfl_up_sync(full, sA, o), pp = __shfl_up_sync(full, sP, o);
            if (lane >= o) { sP = fma(sA, pp, sP); sA *= pa; }
        }
        phi = sP;      // inclusive scan value = new Phi of this node (carry-in of node 0 is its own fixed value)
    }
    if (active) phi_g[lane] = phi;
    __syncwarp();
}

// physical slot of node i of a level whose chunks of 2^sh nodes carry one padding slot (sh = 31: no padding)
__device__ __forceinline__ int slot(int i, int sh) { return i + (i >> sh); }

__device__ __forceinline__ void restrict_nodes(const double* pf, const double* sf, int shf, double* pc, double* sc, int shc, int nc,
                                               double dc, int tid, int nthr)
{   // Restrict, PoissonSolver.cpp:126-157
    for (int i = tid; i < nc; i += nthr) {
        double v = 0.;
        if (i > 0 && i < nc - 1) {
            const int k = 2 * i;
            const double lft = pf[slot(k - 1, shf)], mid = pf[slot(k, shf)], rgt = pf[slot(k + 1, shf)];
            v = 4. * (sf[slot(k, shf)] + lft - 2. * mid + rgt) - dc * (rgt - lft);
        }
        pc[slot(i, shc)] = 0.;
        sc[slot(i, shc)] = v;
    }
}

__device__ __forceinline__ void prolong_nodes(const double* pc, int shc, double* pf, int shf, int nc, int tid, int nthr)
{   // Prolong, PoissonSolver.cpp:110-123
    for (int i = tid; i < nc; i += nthr) {
        const double c = pc[slot(i, shc)];
        pf[slot(2 * i, shf)] += c;
        if (i > 0) pf[slot(2 * i - 1, shf)] += 0.5 * (pc[slot(i - 1, shc)] + c);
    }
}

// One multigrid hierarchy of one density; every method is called by all threads of the CTA.
// Levels with at most kSmemLevelNodes nodes live in (padded) shared memory, the finer ones in global memory (L2).
constexpr int kSmemLevelNodes = 4096;

__host__ __device__ inline int level_npt(int n) { return n > kPT ? n / kPT : 1; }          // nodes per thread of a level
__host__ __device__ inline int level_slots(int n) { return n + 1 + (level_npt(n) > 1 ? n / level_npt(n) : 0); }

struct Hierarchy {
    double* phi; double* src;       // global arrays (all levels, plain layout)
    double* sphi; double* ssrc;     // shared-memory arrays of the coarse levels (padded layout), may be null
    const PoissonLevels& lv;
    double delta;
    PoissonSmem& sm;
    bool pending;        // warp 0 wrote coarse levels that the other warps have not synchronised with yet

    __device__ __forceinline__ double dl(int l) const { return delta * (double)(1 << l); }
    __device__ __forceinline__ bool warp_level(int l) const { return lv.size[l] - 1 <= kWarpLevelNodes; }
    __device__ __forceinline__ bool in_smem(int l) const { return sphi != nullptr && l >= 1 && lv.size[l] - 1 <= kSmemLevelNodes; }
    __device__ __forceinline__ int shift(int l) const
    {
        const int npt = level_npt(lv.size[l] - 1);
        return (in_smem(l) && npt > 1) ? 31 - __clz(npt) : 31;
    }
    __device__ __forceinline__ double* P(int l) const { return in_smem(l) ? sphi + sm.soff[l] : phi + lv.off[l]; }
    __device__ __forceinline__ double* S(int l) const { return in_smem(l) ? ssrc + sm.soff[l] : src + lv.off[l]; }
    __device__ __forceinline__ void block_begin() { if (pending) { __syncthreads(); pending = false; } }

    // IterateGaussSeidel(l, ., sweeps) without the early exit
    __device__ __forceinline__ void smooth(int l, int sweeps)
    {
        double* p = P(l);
        const double* s = S(l);
        const int size = lv.size[l], n = size - 1;
        if (warp_level(l)) {
            if (threadIdx.x < 32) {
                if (threadIdx.x == 0) sm.updates += (unsigned long long)sweeps * (unsigned long long)(n - 1);
                visit_warp(p, s, size, dl(l), sweeps);
            }
            pending = true;
            return;
        }
        block_begin();
        const int pad = (in_smem(l) && n > kPT) ? 1 : 0;
        if (n > kPT * kMaxNpt) { for (int k = 0; k < sweeps; ++k) gs_sweep(p, s, size, dl(l), sm); }
        else if (n > kPT * 16) visit_regs<32, false>(p, s, size, dl(l), sweeps, pad, sm);
        else if (n > kPT * 8) visit_regs<16, true>(p, s, size, dl(l), sweeps, pad, sm);
        else if (n > kPT * 4) visit_regs<8, true>(p, s, size, dl(l), sweeps, pad, sm);
        else if (n > kPT * 2) visit_regs<4, true>(p, s, size, dl(l), sweeps, pad, sm);
        else if (n > kPT) visit_regs<2, true>(p, s, size, dl(l), sweeps, pad, sm);
        else visit_regs<1, true>(p, s, size, dl(l), sweeps, 0, sm);
    }
    __device__ __forceinline__ void restrict_to(int l)       // level l-1 -> l
    {
        if (warp_level(l)) {                                  // <= 33 coarse nodes: warp 0 alone (level l-1 is complete: either
            if (threadIdx.x < 32) {                           //  a block op ended with a barrier or warp 0 wrote it itself)
                restrict_nodes(P(l - 1), S(l - 1), shift(l - 1), P(l), S(l), shift(l), lv.size[l], dl(l), threadIdx.x, 32);
                __syncwarp();
            }
            pending = true;
            return;
        }
        block_begin();
        restrict_nodes(P(l - 1), S(l - 1), shift(l - 1), P(l), S(l), shift(l), lv.size[l], dl(l), threadIdx.x, blockDim.x);
        __syncthreads();
    }
    __device__ __forceinline__ void prolong_from(int l)      // level l -> l-1
    {
        if (warp_level(l - 1)) {
            if (threadIdx.x < 32) { prolong_nodes(P(l), shift(l), P(l - 1), shift(l - 1), lv.size[l], threadIdx.x, 32); __syncwarp(); }
            pending = true;
            return;
        }
        block_begin();
        prolong_nodes(P(l), shift(l), P(l - 1), shift(l - 1), lv.size[l], threadIdx.x, blockDim.x);
        __syncthreads();
    }
    // shared-memory offsets of the coarse levels (thread 0 fills sm.soff; returns doubles needed per array)
    __device__ __forceinline__ void layout_smem()
    {
        if (threadIdx.x == 0) {
            int off = 0;
            for (int l = 0; l < lv.L; ++l) {
                sm.soff[l] = off;
                if (l >= 1 && lv.size[l] - 1 <= kSmemLevelNodes) off += (level_slots(lv.size[l] - 1) + 3) & ~3;
            }
        }
        __syncthreads();
    }
    __device__ __forceinline__ void to_coarse(int from, int to)      // "Ascend", PoissonSolver.cpp:162-171
    {
        for (int l = from; l < to; ++l) { smooth(l, 3); restrict_to(l + 1); }
        smooth(to, 3);
    }
    __device__ __forceinline__ void to_fine(int from, int to)        // "Descend", PoissonSolver.cpp:173-186
    {
        for (int l = from; l > to; --l) { prolong_from(l); smooth(l - 1, 3); }
    }
    // the last fine-grid visit of a solve: two register sweeps + one generic sweep that also returns the update norm
    __device__ __forceinline__ double to_fine_with_norm(int from)
    {
        for (int l = from; l > 1; --l) { prolong_from(l); smooth(l - 1, 3); }
        if (from >= 1) prolong_from(1);
        smooth(0, 2);
        block_begin();
        return gs_sweep(P(0), const_cast<const double*>(S(0)), lv.size[0], dl(0), sm);
    }
};

// error-free transformations (Dekker / Knuth); the intrinsics keep nvcc from contracting or re-associating them
__device__ __forceinline__ void two_sum(double a, double b, double& s, double& e)
{
    s = __dadd_rn(a, b);
    const double bb = __dadd_rn(s, -a);
    e = __dadd_rn(__dadd_rn(a, -__dadd_rn(s, -bb)), __dadd_rn(b, -bb));
}
__device__ __forceinline__ void two_prod(double a, double b, double& p, double& e)
{
    p = __dmul_rn(a, b);
    e = __fma_rn(a, b, -p);
}
__device__ __forceinline__ void dd_add(double& hi, double& lo, double x, double xe)
{
    double s, e;
    two_sum(hi, x, s, e);
    hi = s;
    lo = __dadd_rn(lo, __dadd_rn(e, xe));
}

// residual of the fine-grid equation  U_{i-1}(1+δ/2) - 2 U_i + U_{i+1}(1-δ/2) = -S_i  in double-double.
// In FP64 the three U terms cancel to ~1e-7 of their size, which is what floors the plain iteration at ~1e-9 in U
// (SURVEY fact 3); evaluated with error-free transformations the residual is exact to ~1e-30.
__device__ __forceinline__ double dd_residual(double S, double um, double u0, double up, double cl, double cr)
{
    double hi = S, lo = 0., p, e;
    two_prod(cl, um, p, e); dd_add(hi, lo, p, e);
    dd_add(hi, lo, -2. * u0, 0.);
    two_prod(cr, up, p, e); dd_add(hi, lo, p, e);
    return __dadd_rn(hi, lo);
}

__global__ void __launch_bounds__(kPT) poisson_full_kernel(GridDev g, PoissonLevels lv, PoissonArgs a)
{
    __shared__ PoissonSmem sm;
    extern __shared__ double dyn_smem[];
    const int k = blockIdx.x;
    if (threadIdx.x == 0) sm.updates = 0;
    if (a.skip && *reinterpret_cast<const int*>(reinterpret_cast<const char*>(a.skip) + (size_t)k * a.skip_stride_bytes)) return;
    double* phi = a.phi + (size_t)k * lv.total;
    double* src = a.src + (size_t)k * lv.total;
    const int N = g.N, L = lv.L, c = L - 1;
    Hierarchy h{ phi, src, a.smem_doubles ? dyn_smem : nullptr, a.smem_doubles ? dyn_smem + a.smem_doubles : nullptr, lv, g.delta, sm, false };
    h.layout_smem();

    // Source_0 (PoissonSolver.h:55-74) and Initialize (PoissonSolver.cpp:80-106)
    if (a.rho) {
        const double* rho = a.rho + (size_t)k * N;
        for (int i = threadIdx.x; i < N; i += blockDim.x) { src[i] = g.psrc[i] * rho[i]; phi[i] = 0.; }
    } else {
        for (int i = threadIdx.x; i < N; i += blockDim.x) phi[i] = 0.;
    }
    __syncthreads();
    for (int l = 1; l < L; ++l) {
        const double* sf = h.S(l - 1);
        double* sc = h.S(l);
        double* pc = h.P(l);
        const int n = lv.size[l], shf = h.shift(l - 1), shc = h.shift(l);
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            sc[slot(i, shc)] = (i > 0 && i < n - 1) ? 4. * sf[slot(2 * i, shf)] : 0.;
            pc[slot(i, shc)] = 0.;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        h.P(c)[0] = 0.;                                               // SetBoundaries(0, Z), PoissonSolver.h:76
        h.P(c)[lv.size[c] - 1] = a.Zbc ? (double)a.Zbc[k] : 0.;
    }
    __syncthreads();
    h.smooth(c, 2);        // the coarsest level has one interior node: the reference's <= 15 sweeps converge in one

    // FullCycle, PoissonSolver.h:89-124
    for (int l = L - 2; l > 0; --l) {
        h.to_fine(c, l);
        h.to_coarse(l, c);
    }
    h.to_fine(c, 0);
    double err = 0., prev = 1e300;
    int used = 0, stagnant = 0;
    for (int it = 0; it < a.max_vcycles; ++it) {
        h.to_coarse(0, c);
        const bool last = (it == a.max_vcycles - 1);
        ++used;
        if (last || a.floor_stop) {
            err = h.to_fine_with_norm(c);
            if (err < 1e-14) break;                                   // PoissonSolver.h:120
            // the update norm contracts ~25x per cycle until it reaches its FP64 rounding floor
            if (a.floor_stop) { if (err > 0.25 * prev) { if (++stagnant >= 2) break; } else stagnant = 0; }
            prev = err;
        } else {
            h.to_fine(c, 0);
        }
    }
    h.block_begin();
    // Defect correction (beyond the reference): one residual in double-double, then the error equation A e = r is solved
    // by the same V-cycles from e = 0 (its own rounding floor is ~1e-9 |e|, i.e. negligible) and U <- U + e.  The result
    // is the discrete solution to FP64 representation accuracy instead of ~1e-9, which removes the rounding-noise floor
    // of the SCF energies (the reference's |dE/E| wanders at 2e-11..1e-10 before it randomly dips below 1e-11).
    if (a.refine_vcycles > 0 && a.u0) {
        double* u0 = a.u0 + (size_t)k * N;
        const double cl = 1. + 0.5 * g.delta, cr = 1. - 0.5 * g.delta;
        // r into a register, U0 saved, then Source_0 <- r, Phi_0 <- 0 (all levels' Phi are re-zeroed by restriction)
        for (int i = threadIdx.x; i < N; i += blockDim.x)
            u0[i] = (i > 0 && i < N - 1) ? dd_residual(src[i], phi[i - 1], phi[i], phi[i + 1], cl, cr) : 0.;
        __syncthreads();
        for (int i = threadIdx.x; i < N; i += blockDim.x) {      // same thread owns node i in both passes
            const double r = u0[i];
            u0[i] = phi[i]; src[i] = r; phi[i] = 0.;
        }
        __syncthreads();
        for (int it = 0; it < a.refine_vcycles; ++it) { h.to_coarse(0, c); h.to_fine(c, 0); }
        h.block_begin();
        for (int i = threadIdx.x; i < N; i += blockDim.x) phi[i] += u0[i];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        if (a.work) atomicAdd(a.work, sm.updates);
        if (a.vcycles_used) a.vcycles_used[k] = used;
        if (a.last_err) a.last_err[k] = err;
    }
}

// doubles per shared-memory array (phi or src) for the coarse levels of an L-level hierarchy
static int smem_doubles_for(const PoissonLevels& lv)
{
    int off = 0;
    for (int l = 1; l < lv.L; ++l)
        if (lv.size[l] - 1 <= kSmemLevelNodes) off += (level_slots(lv.size[l] - 1) + 3) & ~3;
    return off;
}

void launch_poisson_full(const GridDev& g, const PoissonLevels& lv, const PoissonArgs& a_in, cudaStream_t st)
{
    PoissonArgs a = a_in;
    a.smem_doubles = smem_doubles_for(lv);
    const size_t bytes = (size_t)a.smem_doubles * 2 * sizeof(double);
    static bool attr_set = false;
    if (!attr_set) { cudaFuncSetAttribute(poisson_full_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); attr_set = true; }
    poisson_full_kernel<<<a.n_dens, kPT, bytes, st>>>(g, lv, a);
}

__global__ void __launch_bounds__(kPT) poisson_vcycles_kernel(double delta, PoissonLevels lv, double* phi_all, double* src_all,
                                                             int n_cycles, double* last_err)
{
    __shared__ PoissonSmem sm;
    const int k = blockIdx.x;
    if (threadIdx.x == 0) sm.updates = 0;
    double* phi = phi_all + (size_t)k * lv.total;
    double* src = src_all + (size_t)k * lv.total;
    const int c = lv.L - 1;
    Hierarchy h{ phi, src, nullptr, nullptr, lv, delta, sm, false };      // microbench / parity entry point: all levels in global memory
    h.layout_smem();
    double err = 0.;
    for (int it = 0; it < n_cycles; ++it) {
        h.to_coarse(0, c);
        if (it == n_cycles - 1) err = h.to_fine_with_norm(c); else h.to_fine(c, 0);
    }
    h.block_begin();
    if (threadIdx.x == 0 && last_err) last_err[k] = err;
}

void launch_poisson_vcycles(int L, double delta, const PoissonLevels& lv, int n_dens, double* phi, double* src, int n_cycles,
                            double* last_err, cudaStream_t st)
{
    (void)L;
    poisson_vcycles_kernel<<<n_dens, kPT, 0, st>>>(delta, lv, phi, src, n_cycles, last_err);
}

}  // namespace dft
