// Radial Poisson V-cycles for grids of 2049 .. 16385 nodes: ONE THREAD-BLOCK CLUSTER (8 CTAs on 8 SMs) PER DENSITY, the whole
// multigrid hierarchy resident in distributed shared memory.
//
// Replaces (reference DFTAtom/) PoissonSolver.h:155-159 VCycle, PoissonSolver.cpp:40-64 GaussSeidel, :110-123 Prolong, :126-157
// Restrict, :162-197 Ascend/Descend for the warm-started solves of the SCF (from SCF step `warm_after` on the solve is
// `warm_vcycles` V-cycles from the previous step's U; the cold full-multigrid solves of the first steps stay with
// poisson_full_kernel).  Same operators, same order of sweeps / restriction / prolongation as poisson.cu; what changes is where
// the data lives and how many SMs work on one density:
//  * the levels with >= 2048 nodes are cut into 8 slabs, one per CTA of the cluster, Phi and Source of every slab in that CTA's
//    shared memory.  Nothing of the hierarchy touches L2 / HBM between the import of (rho, U_prev) and the export of U: the
//    compulsory 24 N bytes per solve;
//  * a level visit (3 or 6 lexicographic Gauss-Seidel sweeps) needs old values <= `sweeps` nodes to the right of a node and,
//    because a = (1 + d_l/2)/2 ~ 1/2, new values <= 64 nodes to its left per sweep (a^64 < 1e-19): a CTA sweeps a WINDOW = its
//    slab + a halo (64 / 96 nodes left for 3 / 6 sweeps, 8 right) read from the neighbour CTAs through distributed shared
//    memory, with both window ends held fixed; inside the slab the result is the sweep of the whole level to FP64 resolution.
//    No carry crosses a CTA inside a visit: the cluster only meets at two split barriers per visit (arrive after the window
//    is loaded / wait before the slab is stored; arrive after the stores / wait before the next visit loads), both overlapped
//    with work;
//  * inside the window the sweep is the zero-carry local recurrence + truncated affine scan + patch of poisson.cu, with the
//    window in registers and ONE block barrier per sweep;
//  * the levels with <= 1024 nodes belong to CTA 0: 1024 nodes by the whole CTA, 512 .. 64 by warp 0 alone (no block barrier),
//    the sub-cycle below the 32-node level is the precomputed dense operator of the grid (coarse_op_kernel).
#include "internal.h"
#include "poisson_tri.cuh"
#include <cooperative_groups.h>
#include <algorithm>
#include <cmath>

namespace cg = cooperative_groups;

namespace dft {
namespace {

constexpr int kCL = 8;          // CTAs per density
constexpr int kCT = 256;        // threads per CTA
constexpr int kHR = 8;          // right halo (>= sweeps - 1)
constexpr int kHB = 96;         // left halo buffer in front of every slab array (the largest left halo: 6-sweep visits)
constexpr double kTinyC = 1e-19;
enum { kCLoad = 1, kCProlong = 2, kCRestrict = 4 };

struct CLevel {
    int n;                  // owned nodes; node n is the right boundary (Z on level 0, 0 on the correction levels)
    int m;                  // distributed levels: slab nodes per CTA
    int offP, offS;         // offsets (doubles) inside the CTA's dynamic shared memory.  Distributed levels: [kHB left-halo buffer][m slab][kHR right-halo
                            // buffer] - the halo buffers are written by the NEIGHBOUR CTAs (remote stores at the end of their visits), so a window is one
                            // contiguous piece of this CTA's own shared memory
    double d, a, b;         // d_l = delta 2^l; a = (1 + d/2)/2; b = (1 - d/2)/2       (PoissonSolver.cpp:56-57)
    int npt, nsteps;        // nodes per thread of this level's visits; warp-scan steps that still matter
    double Ap[5], B;        // A^(2^j), A = a^npt (one thread's affine map); A^32 (one warp)
    double apow[16];        // a^(k+1): the carry patch of a thread's k-th node
};

struct CShared {
    CLevel lv[16];
    int L, n_dist, lb, md;              // levels; number of distributed levels; block-local level (1024 nodes); dense level (32 nodes)
    int offCb;                          // every CTA's copy of its share (+ halos) of Phi of the block-local level, pushed by CTA 0: [kHB][m_last/2][kHR]
    double wtot[2][kCT / 32];           // per-warp scan totals, double buffered by sweep parity
    double ufirst[2][kCT / 32 + 1];     // unpatched first value of every warp
    double uinit[kCT / 32 + 1];         // first value of every warp at the start of a visit
    double left_adj;                    // new value of the node left of the slab (restriction of the slab's first coarse node)
    double right_bc;                    // Phi_0[n]
    unsigned long long updates;
    long long dbg[32];                  // development aid (thread 0 of every CTA): [0] total [1] local sub-cycle [2] first barrier wait of the visits
                                        // [3] second barrier wait [4] window load [5] sweeps [6] store + restrict [8 + l] visits of level l
    int dbg_on;
};
__shared__ CShared cs;
extern __shared__ double c_dyn[];

__device__ __forceinline__ double pow_int(double a, int e) { double r = 1.; for (int k = 0; k < e; ++k) r *= a; return r; }

// owner-major slots of the local levels of CTA 0: 1024 nodes on 256 threads x 4, 512 .. 32 nodes on 32 lanes x n/32
__device__ __forceinline__ int slot_local(int n, int i)
{
    if (i >= n) return n;
    if (n >= 1024) return (i & 3) * kCT + (i >> 2);
    const int npt = n >> 5;
    return npt <= 1 ? i : (i % npt) * 32 + i / npt;
}

struct ScanK { double Am[5]; double Alane; double B; int nsteps; };
__device__ __forceinline__ ScanK scan_constants(const CLevel& c, int lane)      // per-lane view of the level's precomputed constants
{
    ScanK s;
    s.B = c.B;
    s.nsteps = c.nsteps;
    s.Alane = 1.;
#pragma unroll
    for (int j = 0; j < 5; ++j) {
        s.Am[j] = (lane >= (1 << j) && j < s.nsteps) ? c.Ap[j] : 0.;
        if ((lane >> j) & 1) s.Alane *= c.Ap[j];
    }
    return s;
}
__device__ __forceinline__ void level_constants(CLevel& c)
{
    c.Ap[0] = pow_int(c.a, c.npt);
    for (int j = 1; j < 5; ++j) c.Ap[j] = c.Ap[j - 1] * c.Ap[j - 1];
    c.B = c.Ap[4] * c.Ap[4];
    c.nsteps = 5;
    for (int j = 4; j >= 0; --j) if (c.Ap[j] < kTinyC) c.nsteps = j;
    double q = c.a;
    for (int k = 0; k < 16; ++k) { c.apow[k] = q; q *= c.a; }
}
__device__ __forceinline__ double warp_scan(double P, const ScanK& s)
{
    const unsigned full = 0xffffffffu;
    P = fma(s.Am[0], __shfl_up_sync(full, P, 1), P);
    if (s.nsteps > 1) {
        P = fma(s.Am[1], __shfl_up_sync(full, P, 2), P);
        if (s.nsteps > 2) {
            P = fma(s.Am[2], __shfl_up_sync(full, P, 4), P);
            if (s.nsteps > 3) {
                P = fma(s.Am[3], __shfl_up_sync(full, P, 8), P);
                P = fma(s.Am[4], __shfl_up_sync(full, P, 16), P);
            }
        }
    }
    return P;
}

// ---------------------------------------------------------------------------------------------------------
// Visit of a distributed level by the whole cluster: every CTA sweeps its window (registers), stores its slab.
// ---------------------------------------------------------------------------------------------------------
template <int NPT>
__device__ __noinline__ void visit_dist(int l, int flags, int sweeps)
{
    const unsigned full = 0xffffffffu;
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const CLevel& c = cs.lv[l];
    const int m = c.m, s = rank * m, n = c.n;
    const int HL = sweeps > 3 ? kHB : 64;
    const int i0 = max(s - HL, 0);                  // frozen left end of the window (rank 0: the boundary node 0)
    const int iR = min(s + m + kHR, n);             // frozen right end (last rank: the boundary node n)
    const int jR = iR - i0;
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    const int j0 = t * NPT;
    const int q0 = i0 + j0 - s;                     // slab index of this thread's first node (< 0: left halo, >= m: right halo)
    const double a = c.a, b = c.b;
    double* Pa = c_dyn + c.offP + kHB;              // slab node q at Pa[q], halos at negative q / q >= m
    double* Sa = c_dyn + c.offS + kHB;
    const double rbc = (l == 0) ? cs.right_bc : 0.;
    const int qmax = m + kHR - 1;

    double phi[NPT], hs[NPT];
    const bool dbg = cs.dbg_on && t == 0;
    long long tq0 = dbg ? clock64() : 0;
    cluster.barrier_wait();                         // the stores and halo pushes of the previous visit (all CTAs) are visible
    long long tq1 = dbg ? clock64() : 0;
    // ---- window: Phi, Source / 2 - one contiguous piece of this CTA's own shared memory
#pragma unroll
    for (int k = 0; k < NPT; ++k) {
        const int q = q0 + k, i = s + q;
        const bool live = i < n && i <= iR;
        const int qq = min(q, qmax);
        const double pv = (flags & kCLoad) ? Pa[qq] : 0.;
        const double sv = Sa[qq];
        phi[k] = live ? pv : (i == n ? rbc : 0.);
        hs[k] = live ? 0.5 * sv : 0.;
    }
    if (flags & kCProlong) {                        // Phi_l += P Phi_{l+1}   (Prolong, PoissonSolver.cpp:110-123)
        const CLevel& cc = cs.lv[l + 1];
        const bool cdist = (l + 1 < cs.n_dist);
        const int mc = m >> 1, sc = rank * mc;
        const double* Ca = c_dyn + (cdist ? cc.offP : cs.offCb) + kHB;       // coarse slab node qc at Ca[qc]
        constexpr int NCW = NPT / 2 + 2;            // coarse nodes under this thread's nodes
        const int ifirst = i0 + j0, par = ifirst & 1, icf = ifirst >> 1;
        double cv[NCW];
#pragma unroll
        for (int q = 0; q < NCW; ++q) {
            const int ic = icf + q;
            const int qc = min(ic - sc, mc + kHR - 1);
            const double v = Ca[qc];
            cv[q] = ic < cc.n ? v : 0.;             // the correction vanishes on the boundary (and beyond)
        }
#pragma unroll
        for (int k = 0; k < NPT; ++k) {
            const int i = ifirst + k;
            const int e0 = k >> 1, e1 = (k + 1) >> 1;
            const double even0 = cv[e0], odd0 = 0.5 * (cv[e0] + cv[e0 + 1]);        // first node even: node k sits on coarse e0 (k even) / between e0, e0+1
            const double even1 = cv[e1], odd1 = 0.5 * (cv[e1] + cv[e1 + 1]);        // first node odd
            const double corr = par ? ((k & 1) ? even1 : odd1) : ((k & 1) ? odd0 : even0);
            if (i < n && i <= iR) phi[k] += corr;
        }
    }
    cluster.barrier_arrive();                       // window loaded: the neighbours may overwrite their slabs / push into the halo buffers once everyone got here
    long long tq2 = dbg ? clock64() : 0;

    // ---- sweeps in registers
    const ScanK sk = scan_constants(c, lane);
    double apow[NPT];
#pragma unroll
    for (int k = 0; k < NPT; ++k) apow[k] = c.apow[k];
    if (t == 0 && rank == 0) cs.updates += (unsigned long long)sweeps * (unsigned long long)(n - 1);
    // The two window ends stay fixed.  Left end (thread 0, node 0): the chain starts from zero, so the node keeps its value when its
    // "c" is that value.  Right end (node jR) and the padding behind it: swept like any node (what they hold only flows to the right,
    // out of the window) and the end itself is restored after every sweep.
    const int kR = jR - j0;                         // index of the frozen right end inside this thread (0 <= kR < NPT for one thread)
    const bool hasR = kR >= 0 && kR < NPT;
    double fixR = 0.;
    if (hasR) {
#pragma unroll
        for (int k = 0; k < NPT; ++k) if (k == kR) fixR = phi[k];
    }
    const double fix0 = phi[0];
    // old value of the node after this thread's last one: lane + 1's first node; across warps through shared memory
    if (lane == 0) cs.uinit[w] = phi[0];
    __syncthreads();
    double nb_cross = (w + 1 < kCT / 32) ? cs.uinit[w + 1] : 0.;
    for (int sw = 0; sw < sweeps; ++sw) {
        const int pb = sw & 1;
        double nb = __shfl_down_sync(full, phi[0], 1);
        if (lane == 31) nb = nb_cross;
        double cc_[NPT];
#pragma unroll
        for (int k = 0; k < NPT; ++k) cc_[k] = fma(b, (k + 1 < NPT) ? phi[k + 1] : nb, hs[k]);
        if (t == 0) cc_[0] = fix0;
        double x = 0.;
#pragma unroll
        for (int k = 0; k < NPT; ++k) { x = fma(a, x, cc_[k]); phi[k] = x; }
        const double P = warp_scan(x, sk);
        if (lane == 31) cs.wtot[pb][w] = P;
        if (lane == 0) cs.ufirst[pb][w] = phi[0];
        __syncthreads();
        const double carry = (w > 0) ? cs.wtot[pb][w - 1] : 0.;        // a^(32 NPT) < 1e-19: only the previous warp matters
        double Pex = __shfl_up_sync(full, P, 1);
        if (lane == 0) Pex = 0.;
        double cin = fma(sk.Alane, carry, Pex);
        if (t == 0) cin = 0.;
#pragma unroll
        for (int k = 0; k < NPT; ++k) phi[k] = fma(apow[k], cin, phi[k]);
        if (hasR) {
#pragma unroll
            for (int k = 0; k < NPT; ++k) if (k == kR) phi[k] = fixR;
        }
        if (lane == 31 && w + 1 < kCT / 32) {
            // next sweep's "old" right neighbour = the next warp's first node after ITS patch: its carry-in is this warp's total
            const double u = cs.ufirst[pb][w + 1];
            nb_cross = (j0 + NPT == jR) ? cs.uinit[w + 1] : fma(a, cs.wtot[pb][w], u);
        }
    }
    long long tq3 = dbg ? clock64() : 0;
    cluster.barrier_wait();                         // every CTA has loaded its window: slabs and halo buffers may be overwritten
    long long tq4 = dbg ? clock64() : 0;
    // ---- store the slab; push its two ends into the neighbours' halo buffers (remote stores: nobody waits for them)
    {
        double* Pl = rank > 0 ? cluster.map_shared_rank(Pa, rank - 1) : nullptr;
        double* Pr = rank < kCL - 1 ? cluster.map_shared_rank(Pa, rank + 1) : nullptr;
#pragma unroll
        for (int k = 0; k < NPT; ++k) {
            const int q = q0 + k;
            if (q >= 0 && q < m) {
                Pa[q] = phi[k];
                if (Pr && q >= m - kHB) Pr[q - m] = phi[k];        // the right neighbour's left halo: its slab indices -kHB .. -1
                if (Pl && q < kHR) Pl[m + q] = phi[k];             // the left neighbour's right halo: its slab indices m .. m + kHR - 1
            }
            if (q == -1) cs.left_adj = phi[k];
        }
    }
    if (flags & kCRestrict) {                       // Source_{l+1} = 4 (injected residual) - d_{l+1} (first difference); Phi_{l+1} := 0 implicitly
        __syncthreads();
        const CLevel& cc = cs.lv[l + 1];
        const bool cdist = (l + 1 < cs.n_dist);
        const int mc = m >> 1;
        const double dc = cc.d;
        double* Sc = cdist ? c_dyn + cc.offS + kHB : cluster.map_shared_rank(c_dyn + cc.offS, 0);
        double* Scl = (cdist && rank > 0) ? cluster.map_shared_rank(Sc, rank - 1) : nullptr;
        double* Scr = (cdist && rank < kCL - 1) ? cluster.map_shared_rank(Sc, rank + 1) : nullptr;
        for (int q = t; q < mc; q += kCT) {
            const int ic = rank * mc + q;
            const double lft = (q == 0) ? cs.left_adj : Pa[2 * q - 1], mid = Pa[2 * q], rgt = Pa[2 * q + 1];
            double v = 4. * (Sa[2 * q] + lft - 2. * mid + rgt) - dc * (rgt - lft);
            if (ic == 0) v = 0.;
            if (cdist) {
                Sc[q] = v;
                if (Scr && q >= mc - kHB) Scr[q - mc] = v;
                if (Scl && q < kHR) Scl[mc + q] = v;
            } else Sc[slot_local(cc.n, ic)] = v;
        }
    }
    cluster.barrier_arrive();                       // slab, halo pushes (and the restricted source) stored
    if (dbg) {
        const long long tq5 = clock64();
        cs.dbg[2] += tq1 - tq0; cs.dbg[4] += tq2 - tq1; cs.dbg[5] += tq3 - tq2; cs.dbg[3] += tq4 - tq3; cs.dbg[6] += tq5 - tq4; cs.dbg[8 + l] += tq5 - tq0;
    }
}

// ---------------------------------------------------------------------------------------------------------
// CTA 0: the 1024-node level by the whole CTA (owner-major, 4 nodes per thread)
// ---------------------------------------------------------------------------------------------------------
__device__ __noinline__ void visit_block_local(int l, int flags, int sweeps)
{
    constexpr int NPT = 4, NC = 2;
    const unsigned full = 0xffffffffu;
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    const CLevel& c = cs.lv[l]; const CLevel& cc = cs.lv[l + 1];
    const double a = c.a, b = c.b;
    double* P = c_dyn + c.offP; double* S = c_dyn + c.offS;
    double* Pc = c_dyn + cc.offP; double* Sc = c_dyn + cc.offS;
    double phi[NPT], hs[NPT];
#pragma unroll
    for (int k = 0; k < NPT; ++k) { phi[k] = (flags & kCLoad) ? P[k * kCT + t] : 0.; hs[k] = 0.5 * S[k * kCT + t]; }
    if (flags & kCProlong) {
        // coarse level (512 nodes, warp layout: node i at (i % 16) 32 + i / 16); this thread's coarse nodes: 2t, 2t+1, and 2t+2 of the right neighbour
        double corr[NC + 1];
#pragma unroll
        for (int q = 0; q <= NC; ++q) { const int ic = t * NC + q; corr[q] = ic >= cc.n ? 0. : Pc[slot_local(cc.n, ic)]; }
#pragma unroll
        for (int q = 0; q < NC; ++q) { phi[2 * q] += corr[q]; phi[2 * q + 1] += 0.5 * (corr[q] + corr[q + 1]); }
    }
    const ScanK sk = scan_constants(c, lane);
    double apow[NPT];
#pragma unroll
    for (int k = 0; k < NPT; ++k) apow[k] = c.apow[k];
    if (t == 0) cs.updates += (unsigned long long)sweeps * (unsigned long long)(c.n - 1);
    if (lane == 0) cs.uinit[w] = phi[0];
    __syncthreads();
    double nb_cross = (w + 1 < kCT / 32) ? cs.uinit[w + 1] : 0.;     // last warp: the boundary (0: a correction level)
    double cin = 0.;
    for (int sw = 0; sw < sweeps; ++sw) {
        const int pb = sw & 1;
        double nb = __shfl_down_sync(full, phi[0], 1);
        if (lane == 31) nb = nb_cross;
        double x = 0.;
#pragma unroll
        for (int k = 0; k < NPT; ++k) {
            const double cc_ = fma(b, (k + 1 < NPT) ? phi[k + 1] : nb, hs[k]);
            x = (t == 0 && k == 0) ? phi[0] : fma(a, x, cc_);
            phi[k] = x;
        }
        const double Pw = warp_scan(x, sk);
        if (lane == 31) cs.wtot[pb][w] = Pw;
        if (lane == 0) cs.ufirst[pb][w] = phi[0];
        __syncthreads();
        auto carry_into = [&](int ww) -> double {       // new value of the node before warp ww's first node
            double carry = 0., bp = 1.;
            for (int k = 1; k <= ww && bp >= kTinyC; ++k) { carry = fma(bp, cs.wtot[pb][ww - k], carry); bp *= sk.B; }
            return carry;
        };
        double Pex = __shfl_up_sync(full, Pw, 1);
        if (lane == 0) Pex = 0.;
        cin = fma(sk.Alane, carry_into(w), Pex);
        if (t == 0) cin = 0.;
#pragma unroll
        for (int k = 0; k < NPT; ++k) phi[k] = fma(apow[k], cin, phi[k]);
        if (lane == 31 && w + 1 < kCT / 32) nb_cross = fma(a, carry_into(w + 1), cs.ufirst[pb][w + 1]);
    }
#pragma unroll
    for (int k = 0; k < NPT; ++k) P[k * kCT + t] = phi[k];
    if (flags & kCRestrict) {
#pragma unroll
        for (int q = 0; q < NC; ++q) {
            const int k = 2 * q;
            const double lft = (q == 0) ? cin : phi[k - 1], mid = phi[k], rgt = phi[k + 1];
            double v = 4. * (2. * hs[k] + lft - 2. * mid + rgt) - cc.d * (rgt - lft);
            if (t == 0 && q == 0) v = 0.;
            Sc[slot_local(cc.n, t * NC + q)] = v;
        }
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------------------------------------
// CTA 0, warp 0: the levels of 64 .. 512 nodes (owner-major on 32 lanes), no block barrier
// ---------------------------------------------------------------------------------------------------------
template <int NPT>
__device__ __noinline__ void visit_warp_local(int l, int flags, int sweeps)
{
    constexpr int NC = NPT / 2;
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const CLevel& c = cs.lv[l]; const CLevel& cc = cs.lv[l + 1];
    const double a = c.a, b = c.b;
    double* P = c_dyn + c.offP; double* S = c_dyn + c.offS;
    double* Pc = c_dyn + cc.offP; double* Sc = c_dyn + cc.offS;
    double phi[NPT], hs[NPT];
#pragma unroll
    for (int k = 0; k < NPT; ++k) { phi[k] = (flags & kCLoad) ? P[k * 32 + lane] : 0.; hs[k] = 0.5 * S[k * 32 + lane]; }
    if (flags & kCProlong) {
        double corr[NC + 1];
#pragma unroll
        for (int q = 0; q < NC; ++q) corr[q] = Pc[(NC == 1) ? lane : q * 32 + lane];
        corr[NC] = (lane == 31) ? 0. : Pc[(NC == 1) ? lane + 1 : lane + 1];      // first coarse node of the right neighbour (slot 0*32 + lane+1), boundary: 0
#pragma unroll
        for (int q = 0; q < NC; ++q) { phi[2 * q] += corr[q]; phi[2 * q + 1] += 0.5 * (corr[q] + corr[q + 1]); }
    }
    const ScanK sk = scan_constants(c, lane);
    if (lane == 0) cs.updates += (unsigned long long)sweeps * (unsigned long long)(c.n - 1);
    double cin = 0.;
    for (int sw = 0; sw < sweeps; ++sw) {
        double nb = __shfl_down_sync(full, phi[0], 1);
        if (lane == 31) nb = 0.;                                // right boundary of a correction level
        double x = 0.;
#pragma unroll
        for (int k = 0; k < NPT; ++k) {
            const double cc_ = fma(b, (k + 1 < NPT) ? phi[k + 1] : nb, hs[k]);
            x = (lane == 0 && k == 0) ? phi[0] : fma(a, x, cc_);
            phi[k] = x;
        }
        const double Pw = warp_scan(x, sk);
        cin = __shfl_up_sync(full, Pw, 1);
        if (lane == 0) cin = 0.;
#pragma unroll
        for (int k = 0; k < NPT; ++k) phi[k] = fma(c.apow[k], cin, phi[k]);
    }
#pragma unroll
    for (int k = 0; k < NPT; ++k) P[k * 32 + lane] = phi[k];
    if (flags & kCRestrict) {
#pragma unroll
        for (int q = 0; q < NC; ++q) {
            const int k = 2 * q;
            const double lft = (q == 0) ? cin : phi[k - 1], mid = phi[k], rgt = phi[k + 1];
            double v = 4. * (2. * hs[k] + lft - 2. * mid + rgt) - cc.d * (rgt - lft);
            if (lane == 0 && q == 0) v = 0.;
            Sc[(NC == 1) ? lane : q * 32 + lane] = v;
        }
    }
    __syncwarp();
}

// Phi_md = G Source_md: the whole sub-cycle below the 32-node level (built by coarse_op_kernel, poisson.cu)
__device__ __forceinline__ void dense_apply(const double* G)
{
    const CLevel& c = cs.lv[cs.md];
    const double* S = c_dyn + c.offS; double* P = c_dyn + c.offP;
    const int i = threadIdx.x;
    double a0 = 0., a1 = 0., a2 = 0., a3 = 0.;
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
        a0 = fma(G[(j + 0) * 32 + i], S[j + 0], a0);
        a1 = fma(G[(j + 1) * 32 + i], S[j + 1], a1);
        a2 = fma(G[(j + 2) * 32 + i], S[j + 2], a2);
        a3 = fma(G[(j + 3) * 32 + i], S[j + 3], a3);
    }
    P[i] = (a0 + a1) + (a2 + a3);
    __syncwarp();
}

__device__ __forceinline__ void dist_visit(int l, int flags, int sweeps)
{
    switch (cs.lv[l].m) {
        case 2048: visit_dist<9>(l, flags, sweeps); break;
        case 1024: visit_dist<5>(l, flags, sweeps); break;
        case 512: visit_dist<3>(l, flags, sweeps); break;
        default: visit_dist<2>(l, flags, sweeps); break;      // 256
    }
}
__device__ __forceinline__ void warp_visit(int l, int flags, int sweeps)
{
    switch (cs.lv[l].n) {
        case 512: visit_warp_local<16>(l, flags, sweeps); break;
        case 256: visit_warp_local<8>(l, flags, sweeps); break;
        case 128: visit_warp_local<4>(l, flags, sweeps); break;
        default: visit_warp_local<2>(l, flags, sweeps); break;  // 64
    }
}

// the levels of CTA 0 between the last distributed down-visit and the first distributed up-visit
__device__ __noinline__ void local_subcycle(const double* G, bool tri)
{
    const int lb = cs.lb, md = cs.md;
    if (tri) {
        // exact solve of the 1024-node level by warp 0 (poisson_tri.cuh; G = its table) in place of the sub-cycle below
        if (threadIdx.x < 32) {
            const CLevel& c = cs.lv[lb];
            const double* S = c_dyn + c.offS;
            double* P = c_dyn + c.offP;
            tri_solve_warp(G, c.a, c.b, [&](int i) { return 0.5 * S[slot_local(kTriN, i)]; }, [&](int i, double v) { P[slot_local(kTriN, i)] = v; });
        }
        __syncthreads();
    } else {
        visit_block_local(lb, kCRestrict, 3);
        if (threadIdx.x < 32) {
            for (int l = lb + 1; l < md; ++l) warp_visit(l, kCRestrict, 3);
            dense_apply(G);
            for (int l = md - 1; l > lb; --l) warp_visit(l, kCLoad | kCProlong, 3);
        }
        __syncthreads();
        visit_block_local(lb, kCLoad | kCProlong, 3);
    }
    // every CTA gets its share (+ halos) of Phi_lb for the prolongation into the last distributed level: remote stores into its copy
    {
        cg::cluster_group cluster = cg::this_cluster();
        const CLevel& c = cs.lv[lb];
        const double* P = c_dyn + c.offP;
        const int mc = c.n / kCL;                   // coarse share of a CTA
        const int span = kHB / 2 + mc + kHR;        // coarse slab indices -kHB/2 .. mc + kHR - 1
        for (int e = threadIdx.x; e < kCL * span; e += kCT) {
            const int r = e / span, qc = e % span - kHB / 2;
            const int ic = r * mc + qc;
            if (ic < 0 || ic >= c.n) continue;
            double* dst = cluster.map_shared_rank(c_dyn + cs.offCb + kHB, r);
            dst[qc] = P[slot_local(c.n, ic)];
        }
    }
}

}  // namespace

__global__ void __cluster_dims__(kCL, 1, 1) __launch_bounds__(kCT, 2) poisson_cluster_kernel(GridDev g, ClusterPoissonArgs a)
{
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int k = blockIdx.x / kCL;
    if (a.skip && *reinterpret_cast<const int*>(reinterpret_cast<const char*>(a.skip) + (size_t)k * a.skip_stride_bytes)) return;   // uniform over the cluster
    if (a.step) {
        const int sc = *reinterpret_cast<const int*>(reinterpret_cast<const char*>(a.step) + (size_t)k * a.skip_stride_bytes);
        if (sc < a.step_min || sc >= a.step_max) return;
    }
    const int L = g.L, N = g.N, t = threadIdx.x;
    if (t < L && t < 16) {
        // level table (thread l fills level l): distributed levels (>= 2048 nodes), then CTA 0's local ones; every thread walks the
        // same placement
        int off = 0;
        CLevel c;
        for (int l = 0; l <= t; ++l) {
            c.n = 1 << (L - l);
            c.m = 0; c.offP = c.offS = 0; c.npt = 1;
            if (c.n >= kCL * kCT) {
                c.m = c.n / kCL;
                c.offP = off; off += kHB + c.m + kHR; c.offS = off; off += kHB + c.m + kHR;
                c.npt = c.m == 2048 ? 9 : (c.m == 1024 ? 5 : (c.m == 512 ? 3 : 2));
            }
            else if (c.n >= 32) { c.offP = off; off += c.n + 4; c.offS = off; off += c.n + 4; c.npt = c.n >= 1024 ? 4 : max(1, c.n >> 5); }
        }
        c.d = g.delta * (double)(1 << t);
        c.a = 0.5 * (1. + 0.5 * c.d);
        c.b = 0.5 * (1. - 0.5 * c.d);
        level_constants(c);
        cs.lv[t] = c;
    }
    if (t == 0) {
        int nd = 0;
        for (int l = 0; l < L; ++l) if ((1 << (L - l)) >= kCL * kCT) ++nd;
        cs.L = L; cs.n_dist = nd; cs.lb = L - 10; cs.md = L - 5;
        cs.offCb = a.smem_doubles - kTriTableDoubles - (kHB + 128 + kHR);
        cs.right_bc = (a.Zbc && !a.rho_prev) ? (double)a.Zbc[k] : 0.;      // increment form: dU vanishes on both boundaries
        cs.updates = 0;
        cs.dbg_on = a.dbg != nullptr && k == 0;
        for (int q = 0; q < 32; ++q) cs.dbg[q] = 0;
    }
    const long long t_begin = clock64();
    __syncthreads();
    double* G = c_dyn + a.smem_doubles - kTriTableDoubles;      // the dense operator (32 x 32) or the table of the exact solve
    const bool tri = a.coarse_tri != nullptr;
    if (rank == 0) {
        if (tri) { for (int i = t; i < kTriTableDoubles; i += kCT) G[i] = a.coarse_tri[i]; }
        else { for (int i = t; i < 32 * 32; i += kCT) G[i] = a.coarse_op[i]; }
    }
    // import: Source_0 = r 4 pi K rho (PoissonSolver.h:55-74) and Phi_0 (the previous U, or zero in increment form) - every CTA its slab
    // AND its halos, straight from global memory
    {
        const CLevel c0 = cs.lv[0];
        const int s = rank * c0.m;
        const double* u = a.U + (size_t)k * a.ldU;
        const double* rho = a.rho + (size_t)k * a.rho_stride;
        double* P = c_dyn + c0.offP + kHB; double* S = c_dyn + c0.offS + kHB;
        double* rp = a.rho_prev ? a.rho_prev + (size_t)k * a.rho_stride : nullptr;
        for (int q = t - kHB; q < c0.m + kHR; q += kCT) {
            const int i = s + q;
            if (i < 0 || i > N - 1) continue;
            const double r = rho[i];
            const double base = rp ? rp[i] : 0.;
            P[q] = rp ? 0. : u[i];
            S[q] = (i >= 1 && i < N - 1) ? g.psrc[i] * (r - base) : 0.;
        }
    }
    cluster.sync();         // everybody has read rho_prev of its halos before anybody overwrites it
    if (a.rho_prev) {
        const CLevel c0 = cs.lv[0];
        const int s = rank * c0.m;
        const double* rho = a.rho + (size_t)k * a.rho_stride;
        double* rp = a.rho_prev + (size_t)k * a.rho_stride;
        for (int q = t; q < c0.m; q += kCT) rp[s + q] = rho[s + q];
        if (rank == kCL - 1 && t == 0) rp[N - 1] = rho[N - 1];
    }
    __syncthreads();
    cluster.barrier_arrive();

    const int nd = cs.n_dist;
    const int nv = a.n_vcycles;
    for (int cyc = 0; cyc < nv; ++cyc) {
        // down-leg (the level-0 down-visit of every cycle but the first was fused into the previous top)
        for (int l = (cyc == 0 ? 0 : 1); l < nd; ++l) dist_visit(l, (l == 0 ? kCLoad : 0) | kCRestrict, 3);
        // CTA 0: the local levels; the others wait at the barrier
        const long long tl0 = clock64();
        cluster.barrier_wait();
        const long long tl1 = clock64();
        if (rank == 0) local_subcycle(G, tri);
        cluster.barrier_arrive();
        if (t == 0 && cs.dbg_on) { cs.dbg[1] += clock64() - tl1; cs.dbg[7] += tl1 - tl0; }
        // up-leg
        for (int l = nd - 1; l >= 1; --l) dist_visit(l, kCLoad | kCProlong, 3);
        if (cyc == nv - 1) dist_visit(0, kCLoad | kCProlong, 3);
        else dist_visit(0, kCLoad | kCProlong | kCRestrict, 6);
    }
    cluster.barrier_wait();                 // also: no CTA leaves while a neighbour may still read its shared memory
    {
        const CLevel c0 = cs.lv[0];
        const int s = rank * c0.m;
        double* u = a.U + (size_t)k * a.ldU;
        const double* P = c_dyn + c0.offP + kHB;
        if (a.rho_prev) { for (int q = t; q < c0.m; q += kCT) u[s + q] += P[q]; }
        else {
            for (int q = t; q < c0.m; q += kCT) u[s + q] = P[q];
            if (rank == kCL - 1 && t == 0) u[N - 1] = cs.right_bc;
        }
    }
    if (rank == 0 && t == 0 && a.work) atomicAdd(a.work, cs.updates);
    if (t == 0 && cs.dbg_on) { cs.dbg[0] = clock64() - t_begin; for (int q = 0; q < 32; ++q) a.dbg[rank * 32 + q] = cs.dbg[q]; }
}

// shared memory per CTA (doubles): slabs of the distributed levels + CTA 0's local levels + the dense operator
static int cluster_smem_doubles(int L)
{
    int off = 0;
    for (int l = 0; l < L; ++l) {
        const int n = 1 << (L - l);
        if (n >= kCL * kCT) off += 2 * (kHB + n / kCL + kHR);
        else if (n >= 32) off += 2 * (n + 4);
    }
    return off + (kHB + 128 + kHR) + kTriTableDoubles;
}

bool poisson_cluster_supported(int L, double delta)
{
    if (L < 11 || L > 14 || !(delta > 0.)) return false;
    // the window scheme and the one-warp carry need a^64 < 1e-19 on every distributed level
    const int last_dist = L - 11;
    const double a = 0.5 * (1. + 0.5 * delta * (double)(1 << last_dist));
    return std::pow(a, 64.) < kTinyC;
}

int poisson_cluster_init_device()
{
    DFT_CHECK(cudaFuncSetAttribute(poisson_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, cluster_smem_doubles(14) * (int)sizeof(double)));
    return 0;
}

void launch_poisson_cluster(const GridDev& g, const ClusterPoissonArgs& a_in, cudaStream_t st)
{
    ClusterPoissonArgs a = a_in;
    a.smem_doubles = cluster_smem_doubles(g.L);
    poisson_cluster_kernel<<<a.n_dens * kCL, kCT, (size_t)a.smem_doubles * sizeof(double), st>>>(g, a);
}

}  // namespace dft
