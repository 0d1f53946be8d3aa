// Streaming level visits of the radial Poisson multigrid for grids that do not fit on chip and MANY densities
// (config C5a: 2^20 + 1 nodes x 1024 densities; SURVEY section 8d: HBM-bound, 112 N bytes per V-cycle per density).
//
// Replaces, for the levels above 16384 nodes, (reference DFTAtom/) PoissonSolver.cpp:40-64 GaussSeidel (x3 per visit,
// :66-77), :126-157 Restrict and :110-123 Prolong, driven as VCycle (PoissonSolver.h:155-159, .cpp:162-197).
//
// One launch = one level visit of ALL densities: grid (slabs of the level, densities).  A CTA owns a slab of consecutive
// nodes and works on a window = slab + halo held in shared memory: because the lexicographic Gauss-Seidel sweep is the
// recurrence Phi_i <- a Phi_{i-1} + c_i with a ~ 1/2, a new value depends on new values at most ~64 nodes to its left
// (a^64 < 1e-19) and on old values `sweeps` nodes to its right, so a window whose ends are held fixed reproduces the sweep
// of the whole level inside the slab to FP64 resolution (halo: 128 / 192 nodes on the left for 3 / 6 sweeps, 8 on the
// right) and the CTAs need no carry exchange.  Every level is read once and written once per visit:
//   down-leg visit : [read Phi_l (top level only)] read Source_l, 3 sweeps, write Phi_l, write Source_{l+1} (restriction)
//   up-leg visit   : read Phi_l, Source_l, Phi_{l+1} (prolongation), 3 sweeps, write Phi_l
//   fused top      : the up-leg visit of cycle k and the down-leg visit of cycle k+1 of level 0 are ONE visit with 6 sweeps.
// Global -> shared copies are cp.async (16 B, L1 bypass) straight into a padded layout (row of NPT nodes + 2 pad: the
// per-thread 16-byte accesses are bank-conflict free); results go back through shared memory as coalesced 16-byte stores.
// Inside the window the sweep is evaluated like everywhere else in this library: every thread runs the recurrence over its
// NPT nodes with zero carry-in, the carries are resolved by a truncated scan of the affine maps, and patched in.
// All levels handled here are in natural node order; the levels <= 16384 nodes are run by poisson_mid_kernel (poisson.cu).
#include "internal.h"
#include <algorithm>

namespace dft {

namespace {

constexpr int kHaloRight = 8;
constexpr double kTinyCarry = 1e-19;

template <int T, int NPT> struct StreamGeom {
    static constexpr int W = T * NPT;                 // owned nodes of the window; node W is its fixed right end
    static constexpr int NC = NPT / 2;
    static constexpr int PW = W + 2 * T + 2;          // padded fine window (doubles)
    static constexpr int CW = W / 2 + 2 * T + 2;      // padded coarse window
    static constexpr size_t smem = sizeof(double) * (size_t)(2 * PW + CW);
};

__device__ __forceinline__ void cp_async16(double* smem_dst, const double* gsrc)
{
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all()
{
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

template <int T, int NPT, int MINB>
__global__ void __launch_bounds__(T, MINB) stream_visit_kernel(StreamVisitArgs v)
{
    using G = StreamGeom<T, NPT>;
    constexpr int W = G::W, NC = G::NC;
    constexpr int NW = T / 32;
    extern __shared__ __align__(16) double sm[];
    double* phiL = sm;
    double* srcL = sm + G::PW;          // Source window; reused as staging of the restricted residual
    double* cL = sm + 2 * G::PW;        // coarse Phi window (prolongation)
    __shared__ double s_edge[NW], s_wtot[NW];
    const unsigned full = 0xffffffffu;
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    if (v.skip && *reinterpret_cast<const int*>(reinterpret_cast<const char*>(v.skip) + (size_t)blockIdx.y * v.skip_stride_bytes)) return;
    const int n = v.n;
    const int a0 = blockIdx.x * v.slab;
    const int b0 = min(a0 + v.slab, n);
    int wa = max(a0 - v.HL, 0);
    if (wa + W > n) wa = n - W;
    const size_t kd = blockIdx.y;
    double* gp = v.phi_f + kd * (size_t)v.stride_f + wa;
    const double* gs = v.src_f + kd * (size_t)v.stride_f + wa;
    const double* gpc = v.phi_c + kd * (size_t)v.stride_c + (wa >> 1);
    const bool load_phi = v.flags & kVisitLoadPhi, prolong = v.flags & kVisitProlongIn, restr = v.flags & kVisitRestrictOut;

    // ---- global -> shared (node j of the window at j + 2 (j / NPT)) ----
#pragma unroll
    for (int u = 0; u < NC; ++u) {
        const int j = 2 * (t + u * T);
        const int sj = j + 2 * (j / NPT);
        if (load_phi) cp_async16(phiL + sj, gp + j);
        cp_async16(srcL + sj, gs + j);
    }
    if (prolong) {
#pragma unroll
        for (int u = 0; u < NC / 2; ++u) {
            const int q = 2 * (t + u * T);
            cp_async16(cL + q + 2 * (q / NC), gpc + q);
        }
    }
    double right = load_phi ? __ldcg(gp + W) : 0.;             // window node W (old value, fixed)
    const double cright = prolong ? __ldcg(gpc + (W >> 1)) : 0.;
    cp_async_wait_all();
    __syncthreads();

    // ---- registers: thread t owns window nodes [t NPT, (t+1) NPT) ----
    double phi[NPT], src[NPT];
    {
        const double2* ps = reinterpret_cast<const double2*>(srcL + t * (NPT + 2));
#pragma unroll
        for (int m = 0; m < NC; ++m) { const double2 x = ps[m]; src[2 * m] = 0.5 * x.x; src[2 * m + 1] = 0.5 * x.y; }
        if (load_phi) {
            const double2* pp = reinterpret_cast<const double2*>(phiL + t * (NPT + 2));
#pragma unroll
            for (int m = 0; m < NC; ++m) { const double2 x = pp[m]; phi[2 * m] = x.x; phi[2 * m + 1] = x.y; }
        } else {
#pragma unroll
            for (int k = 0; k < NPT; ++k) phi[k] = 0.;
        }
    }
    if (prolong) {                                             // Prolong, PoissonSolver.cpp:110-123
        double corr[NC + 1];
        const double2* pc = reinterpret_cast<const double2*>(cL + t * (NC + 2));
#pragma unroll
        for (int m = 0; m < NC / 2; ++m) { const double2 x = pc[m]; corr[2 * m] = x.x; corr[2 * m + 1] = x.y; }
        corr[NC] = (t == T - 1) ? cright : cL[(t + 1) * (NC + 2)];
#pragma unroll
        for (int m = 0; m < NC; ++m) {
            phi[2 * m] += corr[m];
            phi[2 * m + 1] += 0.5 * (corr[m] + corr[m + 1]);
        }
        right += cright;
    }

    // ---- sweeps (GaussSeidel, PoissonSolver.cpp:40-64) ----
    const double a = v.a, bcoef = v.bcoef;
    double Ap[5];
    {
        double A = a;
#pragma unroll
        for (int k = 1; k < NPT; k <<= 1) A *= A;
        Ap[0] = A;
#pragma unroll
        for (int j = 1; j < 5; ++j) Ap[j] = Ap[j - 1] * Ap[j - 1];
    }
    double Am[5];
    double Alane = 1.;
    int nsteps = 5;
#pragma unroll
    for (int j = 4; j >= 0; --j) if (Ap[j] < kTinyCarry) nsteps = j;
#pragma unroll
    for (int j = 0; j < 5; ++j) {
        Am[j] = (lane >= (1 << j) && j < nsteps) ? Ap[j] : 0.;
        if ((lane >> j) & 1) Alane *= Ap[j];
    }
    const bool cross = Ap[4] * Ap[4] >= kTinyCarry;            // does a carry survive a whole warp?  (never for NPT >= 4)
    double cin = 0.;
    for (int sw = 0; sw < v.sweeps; ++sw) {
        double nb = __shfl_down_sync(full, phi[0], 1);
        if (lane == 0) s_edge[w] = phi[0];
        __syncthreads();
        if (lane == 31 && w + 1 < NW) nb = s_edge[w + 1];
        if (t == T - 1) nb = right;
        double x = 0.;
#pragma unroll
        for (int k = 0; k < NPT; ++k) {
            const double c = fma(bcoef, (k + 1 < NPT) ? phi[k + 1] : nb, src[k]);
            x = (t == 0 && k == 0) ? phi[0] : fma(a, x, c);      // window node 0 is held fixed
            phi[k] = x;
        }
        double Pw = x;
        Pw = fma(Am[0], __shfl_up_sync(full, Pw, 1), Pw);
        if (nsteps > 1) {
            Pw = fma(Am[1], __shfl_up_sync(full, Pw, 2), Pw);
            if (nsteps > 2) {
                Pw = fma(Am[2], __shfl_up_sync(full, Pw, 4), Pw);
                Pw = fma(Am[3], __shfl_up_sync(full, Pw, 8), Pw);
                Pw = fma(Am[4], __shfl_up_sync(full, Pw, 16), Pw);
            }
        }
        if (lane == 31) s_wtot[w] = Pw;
        __syncthreads();
        double carry = (w > 0) ? s_wtot[w - 1] : 0.;            // new value of the last node of the previous warp
        if (cross) {
            double bp = Ap[4] * Ap[4];
            for (int k = 2; k <= w && bp >= kTinyCarry; ++k) { carry = fma(bp, s_wtot[w - k], carry); bp *= Ap[4] * Ap[4]; }
        }
        double Pex = __shfl_up_sync(full, Pw, 1);
        if (lane == 0) Pex = 0.;
        cin = fma(Alane, carry, Pex);                           // new value of the node before this thread's first node
        if (t == 0) cin = 0.;
        double q = a;
#pragma unroll
        for (int k = 0; k < NPT; ++k) { phi[k] = fma(q, cin, phi[k]); q *= a; }
    }

    // ---- results -> shared -> global (the slab only) ----
    {
        double2* pp = reinterpret_cast<double2*>(phiL + t * (NPT + 2));
#pragma unroll
        for (int m = 0; m < NC; ++m) pp[m] = make_double2(phi[2 * m], phi[2 * m + 1]);
    }
    if (restr) {                                                // Restrict, PoissonSolver.cpp:126-157
        const double dc = v.dc;
        double rv[NC];
#pragma unroll
        for (int m = 0; m < NC; ++m) {
            const int k = 2 * m;
            const double lft = (m == 0) ? cin : phi[k - 1], mid = phi[k], rgt = phi[k + 1];
            rv[m] = 4. * (2. * src[k] + lft - 2. * mid + rgt) - dc * (rgt - lft);
        }
        if (wa == 0 && t == 0) rv[0] = 0.;
        double2* pr = reinterpret_cast<double2*>(srcL + t * (NC + 2));
#pragma unroll
        for (int m = 0; m < NC / 2; ++m) pr[m] = make_double2(rv[2 * m], rv[2 * m + 1]);
    }
    __syncthreads();
    for (int i = a0 + 2 * t; i < b0; i += 2 * T) {
        const int j = i - wa;
        *reinterpret_cast<double2*>(gp + j) = *reinterpret_cast<const double2*>(phiL + j + 2 * (j / NPT));
    }
    if (b0 == n && t == 0) gp[W] = right;                       // the level's right boundary (wa + W == n for the last slab)
    if (restr) {
        double* gsc = v.src_c + kd * (size_t)v.stride_c;
        const int wc = wa >> 1;
        for (int i = (a0 >> 1) + 2 * t; i < (b0 >> 1); i += 2 * T) {
            const int q = i - wc;
            *reinterpret_cast<double2*>(gsc + i) = *reinterpret_cast<const double2*>(srcL + q + 2 * (q / NC));
        }
        if (b0 == n && t == 0) gsc[n >> 1] = 0.;
    }
}

// Initialize (PoissonSolver.cpp:80-106) for the streamed levels: Source_{l+1} = 4 x injection of Source_l, zero at the ends
__global__ void __launch_bounds__(256) stream_inject_kernel(const double* __restrict__ src_f, long long stride_f, double* __restrict__ src_c,
                                                           long long stride_c, int nc, const int* skip, int skip_stride_bytes)
{
    const size_t k = blockIdx.y;
    if (skip && *reinterpret_cast<const int*>(reinterpret_cast<const char*>(skip) + k * skip_stride_bytes)) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > nc) return;
    src_c[k * (size_t)stride_c + i] = (i > 0 && i < nc) ? 4. * __ldg(src_f + k * (size_t)stride_f + 2 * (size_t)i) : 0.;
}

// Source_0 = r 4 pi K rho (PoissonSolver.h:55-74; psrc is zero at both ends)
__global__ void __launch_bounds__(256) stream_source_kernel(const double* __restrict__ psrc, const double* __restrict__ rho, long long rho_stride,
                                                           double* __restrict__ src0, long long ld0, int N, const int* skip, int skip_stride_bytes)
{
    const size_t k = blockIdx.y;
    if (skip && *reinterpret_cast<const int*>(reinterpret_cast<const char*>(skip) + k * skip_stride_bytes)) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N) src0[k * (size_t)ld0 + i] = __ldg(psrc + i) * __ldg(rho + k * (size_t)rho_stride + i);
}

template <int T, int NPT, int MINB>
void launch_variant(const StreamVisitArgs& v_in, int n_dens, cudaStream_t st)
{
    using G = StreamGeom<T, NPT>;
    StreamVisitArgs v = v_in;
    v.HL = v.sweeps > 3 ? 192 : 128;
    v.slab = G::W - v.HL - kHaloRight;
    const int slabs = (v.n + v.slab - 1) / v.slab;
    stream_visit_kernel<T, NPT, MINB><<<dim3(slabs, n_dens), T, G::smem, st>>>(v);
}

}  // namespace

int stream_window_nodes(int variant) { return variant == 1 ? 256 * 8 : (variant == 2 ? 512 * 8 : 256 * 16); }

// One level visit of all densities.  variant 0: 256 threads x 16 nodes (window 4096, 2 CTAs per SM); 1: 256 x 8 (window 2048,
// 3-4 CTAs per SM); 2: 512 x 8 (window 4096, 2 CTAs per SM)
void launch_stream_visit(const StreamVisitArgs& v, int n_dens, int variant, cudaStream_t st)
{
    if (variant == 1) launch_variant<256, 8, 3>(v, n_dens, st);
    else if (variant == 2) launch_variant<512, 8, 2>(v, n_dens, st);
    else launch_variant<256, 16, 2>(v, n_dens, st);
}

int stream_init_device()
{   // per-device opt-in to the windows' dynamic shared memory (dftatom_create, under cudaSetDevice)
    DFT_CHECK(cudaFuncSetAttribute(stream_visit_kernel<256, 8, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)StreamGeom<256, 8>::smem));
    DFT_CHECK(cudaFuncSetAttribute(stream_visit_kernel<512, 8, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)StreamGeom<512, 8>::smem));
    DFT_CHECK(cudaFuncSetAttribute(stream_visit_kernel<256, 16, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)StreamGeom<256, 16>::smem));
    return 0;
}

StreamPlan make_stream_plan(int L, int n_dens, int mid_levels)
{
    StreamPlan sp{};
    sp.L = L;
    sp.lv = make_levels(L);
    sp.K = L - mid_levels;                           // size[K] - 1 == 2^mid_levels (16384 for 14, 2048 for 11)
    long long off = 0;
    for (int l = 1; l <= sp.K; ++l) { sp.coff[l] = (int)off; off += (sp.lv.size[l] + 3) & ~3; }
    sp.cstride = off;
    sp.mid_total = sp.lv.total - sp.lv.off[sp.K];
    sp.off_cphi = 0;
    sp.off_csrc = sp.off_cphi + sp.cstride * n_dens;
    sp.off_mphi = sp.off_csrc + sp.cstride * n_dens;
    sp.off_msrc = sp.off_mphi + (long long)sp.mid_total * n_dens;
    sp.total = sp.off_msrc + (long long)sp.mid_total * n_dens;
    return sp;
}

void launch_poisson_stream_vcycles(const StreamPlan& sp, double delta, int n_dens, double* phi0, const double* src0, long long ld0,
                                   double* scratch, const double* coarse_op, int n_cycles, int fuse_tops, int variant, cudaStream_t st,
                                   long long* launches)
{
    const int K = sp.K;
    double* cphi = scratch + sp.off_cphi;
    double* csrc = scratch + sp.off_csrc;
    long long nl = 0;
    auto visit = [&](int l, int flags, int sweeps) {
        StreamVisitArgs v{};
        if (l == 0) { v.phi_f = phi0; v.src_f = src0; v.stride_f = ld0; }
        else { v.phi_f = cphi + sp.coff[l]; v.src_f = csrc + sp.coff[l]; v.stride_f = sp.cstride; }
        v.phi_c = cphi + sp.coff[l + 1]; v.src_c = csrc + sp.coff[l + 1]; v.stride_c = sp.cstride;
        v.n = sp.lv.size[l] - 1;
        v.flags = flags; v.sweeps = sweeps;
        const double d = delta * (double)(1 << l);
        v.a = 0.5 * (1. + 0.5 * d); v.bcoef = 0.5 * (1. - 0.5 * d); v.dc = 2. * d;
        launch_stream_visit(v, n_dens, variant, st);
        ++nl;
    };
    bool top_done = false;      // the down-visit of level 0 was already part of the previous fused top
    for (int c = 0; c < n_cycles; ++c) {
        for (int l = 0; l < K; ++l) {
            if (l == 0 && top_done) continue;
            visit(l, (l == 0 ? kVisitLoadPhi : 0) | kVisitRestrictOut, 3);
        }
        launch_poisson_mid(sp.lv, delta, K, n_dens, cphi + sp.coff[K], csrc + sp.coff[K], sp.cstride, scratch + sp.off_mphi,
                           scratch + sp.off_msrc, sp.mid_total, coarse_op, nullptr, 0, st);
        ++nl;
        for (int l = K - 1; l > 0; --l) visit(l, kVisitLoadPhi | kVisitProlongIn, 3);
        if (fuse_tops && c + 1 < n_cycles) { visit(0, kVisitLoadPhi | kVisitProlongIn | kVisitRestrictOut, 6); top_done = true; }
        else { visit(0, kVisitLoadPhi | kVisitProlongIn, 3); top_done = false; }
    }
    if (launches) *launches = nl;
}

// The chain of cycles  to_coarse(top, c); to_fine(c, next top)  (PoissonSolver.h:89-124) over the streamed levels.  An "arrival"
// is the visit that ends an ascent at its top level b: prolongation in, 3 sweeps - fused with the 3 sweeps and the restriction
// that start the next descent unless it is the last visit of the solve.  fresh: Phi_b is still zero (first visit of the ramp).
void launch_poisson_stream_solve(const StreamPlan& sp, double delta, const StreamSolveArgs& a, cudaStream_t st, long long* launches)
{
    const int K = sp.K, nd = a.n_dens;
    double* cphi = a.scratch + sp.off_cphi;
    double* csrc = a.scratch + sp.off_csrc;
    double* mphi = a.scratch + sp.off_mphi;
    double* msrc = a.scratch + sp.off_msrc;
    long long nl = 0;
    auto visit = [&](int l, int flags, int sweeps) {
        StreamVisitArgs v{};
        if (l == 0) { v.phi_f = a.U; v.src_f = a.src0; v.stride_f = a.ld0; }
        else { v.phi_f = cphi + sp.coff[l]; v.src_f = csrc + sp.coff[l]; v.stride_f = sp.cstride; }
        v.phi_c = cphi + sp.coff[l + 1]; v.src_c = csrc + sp.coff[l + 1]; v.stride_c = sp.cstride;
        v.n = sp.lv.size[l] - 1;
        v.flags = flags; v.sweeps = sweeps;
        const double d = delta * (double)(1 << l);
        v.a = 0.5 * (1. + 0.5 * d); v.bcoef = 0.5 * (1. - 0.5 * d); v.dc = 2. * d;
        v.skip = a.skip; v.skip_stride_bytes = a.skip_stride_bytes;
        launch_stream_visit(v, nd, a.variant, st);
        ++nl;
    };
    auto descend_mid_ascend = [&](int from, int to) {      // down-visits from+1 .. K-1, the levels below, up-visits K-1 .. to+1
        for (int l = from + 1; l < K; ++l) visit(l, kVisitRestrictOut, 3);
        launch_poisson_mid(sp.lv, delta, K, nd, cphi + sp.coff[K], csrc + sp.coff[K], sp.cstride, mphi, msrc, sp.mid_total, a.coarse_op,
                           a.skip, a.skip_stride_bytes, st);
        ++nl;
        for (int l = K - 1; l > to; --l) visit(l, kVisitLoadPhi | kVisitProlongIn, 3);
    };
    const int N = sp.lv.size[0];
    if (a.rho) {
        stream_source_kernel<<<dim3((N + 255) / 256, nd), 256, 0, st>>>(a.psrc, a.rho, a.rho_stride, a.src0, a.ld0, N, a.skip, a.skip_stride_bytes);
        ++nl;
    }
    int remaining;        // arrivals still to come
    if (a.warm) {
        visit(0, kVisitLoadPhi | kVisitRestrictOut, 3);
        remaining = a.n_v;
        for (int q = 0; q < a.n_v; ++q) {
            descend_mid_ascend(0, 0);
            --remaining;
            if (remaining) visit(0, kVisitLoadPhi | kVisitProlongIn | kVisitRestrictOut, 6); else visit(0, kVisitLoadPhi | kVisitProlongIn, 3);
        }
    } else {
        // Initialize: Source_l of every level by injection; the levels <= 16384 nodes run their part of the ramp - every cycle whose
        // top is below the streamed levels, and the descent + ascent back to level K of the cycle whose top is K - as one
        // full-multigrid solve with one V-cycle on the 14-level grid of spacing delta 2^K (same operators, same order)
        for (int l = 0; l < K; ++l) {
            const int nc = sp.lv.size[l + 1] - 1;
            const double* sf = l == 0 ? a.src0 : csrc + sp.coff[l];
            stream_inject_kernel<<<dim3((nc + 256) / 256, nd), 256, 0, st>>>(sf, l == 0 ? a.ld0 : sp.cstride, csrc + sp.coff[l + 1], sp.cstride, nc,
                                                                           a.skip, a.skip_stride_bytes);
            ++nl;
        }
        GridDev gm{};
        gm.N = sp.lv.size[K]; gm.L = sp.L - K; gm.delta = delta * (double)(1 << K);
        const PoissonLevels lvm = make_levels(gm.L);
        PoissonArgs pa{};
        pa.n_dens = nd; pa.src_nat = csrc + sp.coff[K]; pa.u_out = cphi + sp.coff[K]; pa.nat_stride = sp.cstride; pa.Zbc = a.Zbc;
        pa.phi = mphi; pa.src = msrc; pa.max_vcycles = 1; pa.coarse_op = a.coarse_op; pa.skip = a.skip; pa.skip_stride_bytes = a.skip_stride_bytes;
        launch_poisson_full(gm, lvm, pa, st);
        ++nl;
        remaining = K + a.n_v;
        int prev = K;
        for (int q = 0; q < K + a.n_v; ++q) {
            const int b = prev > 0 ? prev - 1 : 0;
            const bool fresh = prev > 0;
            if (q > 0) descend_mid_ascend(prev, b);
            --remaining;
            const int ld = fresh ? 0 : kVisitLoadPhi;
            if (remaining) visit(b, ld | kVisitProlongIn | kVisitRestrictOut, 6); else visit(b, ld | kVisitProlongIn, 3);
            prev = b;
        }
    }
    if (launches) *launches = nl;
}

}  // namespace dft
