// Radial Poisson multigrid, one CTA per density, all levels of the 2^L+1 hierarchy L2-resident.
//
// Replaces (reference DFTAtom/) PoissonSolver.h:51-81 SolvePoissonNonUniform, :89-124 FullCycle, :155-159 VCycle and
// PoissonSolver.cpp:40-64 GaussSeidel, :66-77 IterateGaussSeidel, :80-106 Initialize, :110-123 Prolong, :126-157
// Restrict, :162-197 Ascend/Descend.
//
// The reference's smoother is a lexicographic in-place Gauss-Seidel sweep, i.e. the first-order recurrence
//     Phi_i <- a Phi_{i-1} + c_i ,   a = (1 + d_l/2)/2 ,   c_i = (S_i + (1 - d_l/2) Phi_{i+1}^old)/2 ,  d_l = δ 2^l .
// Here every thread owns M consecutive nodes: it runs the recurrence locally with zero carry-in, the carries are
// resolved by an associative block scan of the affine maps x -> a^m x + p, and the result is patched in.  That is
// the same sweep (same operator, same ordering), evaluated in O(M + log T) depth instead of O(N).
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
constexpr int kPM = 16;      // nodes per thread per pass

struct PoissonSmem {
    double scanA[32], scanP[32];
    double red[32];
    double carry;        // last new value of the previous pass
    double bcast;
    unsigned long long updates;   // Gauss-Seidel node-updates performed by this CTA (work counter)
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

// One lexicographic Gauss-Seidel sweep over level arrays (phi, src) of `size` nodes; returns sqrt(sum (old-new)^2).
__device__ double gs_sweep(double* __restrict__ phi, const double* __restrict__ src, int size, double d, PoissonSmem& sm)
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
            double nxt = phi[i0];
#pragma unroll
            for (int k = 0; k < kPM; ++k) {
                if (k < m) {
                    nxt = phi[i0 + k + 1];
                    const double c = fma(bcoef, nxt, 0.5 * src[i0 + k]);
                    x = fma(a, x, c);
                    p[k] = x;
                    A *= a;
                }
            }
        }
        // inclusive scan of the affine maps (A, x) across the block
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
        // exclusive prefix for this thread = (warps before) o (lanes before)
        double eA = __shfl_up_sync(0xffffffffu, sA, 1), eP = __shfl_up_sync(0xffffffffu, sP, 1);
        if (lane == 0) { eA = 1.; eP = 0.; }
        if (w > 0) { const double wa = sm.scanA[w - 1], wp = sm.scanP[w - 1]; eP = fma(eA, wp, eP); eA *= wa; }
        const double left = sm.carry;
        double cin = fma(eA, left, eP);             // new value of node i0-1
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
        // (visible to all after the next pass's first __syncthreads; thread 0 re-reads it only after that)
        __syncthreads();
    }
    return sqrt(block_sum(err2, sm));
}

__device__ double gs_smooth(double* phi, const double* src, int size, double d, double tol, int sweeps, PoissonSmem& sm)
{   // IterateGaussSeidel, PoissonSolver.cpp:66-77
    double err = 1e10;
    for (int k = 0; k < sweeps; ++k) {
        err = gs_sweep(phi, src, size, d, sm);
        if (err < tol) break;
    }
    return err;
}

__device__ void mg_restrict(const double* __restrict__ pf, const double* __restrict__ sf, double* __restrict__ pc,
                            double* __restrict__ sc, int nc, double dc)
{   // Restrict, PoissonSolver.cpp:126-157
    for (int i = threadIdx.x; i < nc; i += blockDim.x) {
        pc[i] = 0.;
        double v = 0.;
        if (i > 0 && i < nc - 1) {
            const int k = 2 * i;
            const double lft = pf[k - 1], mid = pf[k], rgt = pf[k + 1];
            v = 4. * (sf[k] + lft - 2. * mid + rgt) - dc * (rgt - lft);
        }
        sc[i] = v;
    }
    __syncthreads();
}

__device__ void mg_prolong(const double* __restrict__ pc, double* __restrict__ pf, int nc)
{   // Prolong, PoissonSolver.cpp:110-123
    for (int i = threadIdx.x; i < nc; i += blockDim.x) {
        const double c = pc[i];
        pf[2 * i] += c;
        if (i > 0) pf[2 * i - 1] += 0.5 * (pc[i - 1] + c);
    }
    __syncthreads();
}

struct LevelPtrs { double* phi; double* src; };

__device__ __forceinline__ void to_coarse(double* phi, double* src, const PoissonLevels& lv, double delta, int from, int to,
                                          double tol, PoissonSmem& sm)
{   // "Ascend", PoissonSolver.cpp:162-171
    for (int l = from; l < to; ++l) {
        gs_smooth(phi + lv.off[l], src + lv.off[l], lv.size[l], delta * (double)(1 << l), tol, 3, sm);
        mg_restrict(phi + lv.off[l], src + lv.off[l], phi + lv.off[l + 1], src + lv.off[l + 1], lv.size[l + 1], delta * (double)(1 << (l + 1)));
    }
    gs_smooth(phi + lv.off[to], src + lv.off[to], lv.size[to], delta * (double)(1 << to), tol, 3, sm);
}

__device__ __forceinline__ double to_fine(double* phi, double* src, const PoissonLevels& lv, double delta, int from, int to,
                                          double tol, PoissonSmem& sm)
{   // "Descend", PoissonSolver.cpp:173-186
    double err = 1e10;
    for (int l = from; l > to; --l) {
        mg_prolong(phi + lv.off[l], phi + lv.off[l - 1], lv.size[l]);
        err = gs_smooth(phi + lv.off[l - 1], src + lv.off[l - 1], lv.size[l - 1], delta * (double)(1 << (l - 1)), tol, 3, sm);
    }
    return err;
}

__global__ void __launch_bounds__(kPT) poisson_full_kernel(GridDev g, PoissonLevels lv, PoissonArgs a)
{
    __shared__ PoissonSmem sm;
    const int k = blockIdx.x;
    if (threadIdx.x == 0) sm.updates = 0;
    if (a.skip && *reinterpret_cast<const int*>(reinterpret_cast<const char*>(a.skip) + (size_t)k * a.skip_stride_bytes)) return;
    double* phi = a.phi + (size_t)k * lv.total;
    double* src = a.src + (size_t)k * lv.total;
    const int N = g.N, L = lv.L, c = L - 1;
    const double delta = g.delta;

    // Source_0 (PoissonSolver.h:55-74) and Initialize (PoissonSolver.cpp:80-106)
    if (a.rho) {
        const double* rho = a.rho + (size_t)k * N;
        for (int i = threadIdx.x; i < N; i += blockDim.x) { src[i] = g.psrc[i] * rho[i]; phi[i] = 0.; }
    } else {
        for (int i = threadIdx.x; i < N; i += blockDim.x) phi[i] = 0.;
    }
    __syncthreads();
    for (int l = 1; l < L; ++l) {
        const double* sf = src + lv.off[l - 1];
        double* sc = src + lv.off[l];
        double* pc = phi + lv.off[l];
        const int n = lv.size[l];
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            sc[i] = (i > 0 && i < n - 1) ? 4. * sf[2 * i] : 0.;
            pc[i] = 0.;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        phi[lv.off[c]] = 0.;                                          // SetBoundaries(0, Z), PoissonSolver.h:76
        phi[lv.off[c] + lv.size[c] - 1] = a.Zbc ? (double)a.Zbc[k] : 0.;
    }
    __syncthreads();
    gs_smooth(phi + lv.off[c], src + lv.off[c], lv.size[c], delta * (double)(1 << c), 1e-3, 15, sm);

    // FullCycle, PoissonSolver.h:89-124
    for (int l = L - 2; l > 0; --l) {
        to_fine(phi, src, lv, delta, c, l, 1e-3, sm);
        to_coarse(phi, src, lv, delta, l, c, 1e-3, sm);
    }
    to_fine(phi, src, lv, delta, c, 0, 1e-14, sm);
    double err = 0., prev = 1e300;
    int used = 0, stagnant = 0;
    for (int it = 0; it < a.max_vcycles; ++it) {
        to_coarse(phi, src, lv, delta, 0, c, 1e-14, sm);
        err = to_fine(phi, src, lv, delta, c, 0, 1e-14, sm);
        ++used;
        if (err < 1e-14) break;                                       // PoissonSolver.h:120
        if (a.floor_stop) {
            // the update norm contracts ~25x per cycle until it reaches its FP64 rounding floor;
            // once it stops contracting, further cycles only re-roll the rounding noise.
            if (err > 0.25 * prev) { if (++stagnant >= 2) break; } else stagnant = 0;
        }
        prev = err;
    }
    if (threadIdx.x == 0) {
        if (a.work) atomicAdd(a.work, sm.updates);
        if (a.vcycles_used) a.vcycles_used[k] = used;
        if (a.last_err) a.last_err[k] = err;
    }
}

void launch_poisson_full(const GridDev& g, const PoissonLevels& lv, const PoissonArgs& a, cudaStream_t st)
{
    poisson_full_kernel<<<a.n_dens, kPT, 0, st>>>(g, lv, a);
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
    double err = 0.;
    for (int it = 0; it < n_cycles; ++it) {
        to_coarse(phi, src, lv, delta, 0, c, 1e-14, sm);
        err = to_fine(phi, src, lv, delta, c, 0, 1e-14, sm);
    }
    if (threadIdx.x == 0 && last_err) last_err[k] = err;
}

void launch_poisson_vcycles(int L, double delta, const PoissonLevels& lv, int n_dens, double* phi, double* src, int n_cycles,
                            double* last_err, cudaStream_t st)
{
    (void)L;
    poisson_vcycles_kernel<<<n_dens, kPT, 0, st>>>(delta, lv, phi, src, n_cycles, last_err);
}

}  // namespace dft
