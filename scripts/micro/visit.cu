// micro-benchmark: cycles per level visit of the Poisson kernels, hot in the instruction cache (one CTA)
#include "../../dftatom_b200/csrc/poisson.cu"
#include <cstdio>
using namespace dft;
__global__ void __launch_bounds__(kPT) k_visit(PoissonLevels lv, double delta, double* phi, double* src, int dyn_doubles, long long* out, int reps)
{
    hierarchy_setup(lv, delta, phi, src, dyn_doubles, nullptr, nullptr);
    Ctl ctl{ false, false };
    for (int l = 0; l < lv.L; ++l) {
        const Ref p = Ref::P(l), s = Ref::S(l);
        for (int i = threadIdx.x; i < lv.size[l]; i += blockDim.x) { p.st(i, 1e-3 * i); s.st(i, 1e-6 * (i % 7)); }
    }
    __syncthreads();
    for (int l = 0; l < lv.L; ++l) {
        block_begin(ctl);
        __syncthreads();
        long long t0 = clock64();
        for (int r = 0; r < reps; ++r) smooth(ctl, l, 3);
        block_begin(ctl);
        long long t1 = clock64();
        if (l + 1 < lv.L) { for (int r = 0; r < reps; ++r) restrict_to(ctl, l + 1); }
        block_begin(ctl);
        long long t2 = clock64();
        if (l + 1 < lv.L) { for (int r = 0; r < reps; ++r) prolong_from(ctl, l + 1); }
        block_begin(ctl);
        long long t3 = clock64();
        if (threadIdx.x == 0) { out[3 * l] = (t1 - t0) / reps; out[3 * l + 1] = (t2 - t1) / reps; out[3 * l + 2] = (t3 - t2) / reps; }
    }
}
int main()
{
    const int L = 14; const double delta = 5e-4;
    PoissonLevels lv = make_levels(L);
    double *phi, *src; long long* out;
    cudaMalloc(&phi, 8 * lv.total); cudaMalloc(&src, 8 * lv.total); cudaMalloc(&out, 8 * 3 * 24);
    cudaMemset(phi, 0, 8 * lv.total); cudaMemset(src, 0, 8 * lv.total);
    const int dd = dyn_doubles_for(lv);
    cudaFuncSetAttribute(k_visit, cudaFuncAttributeMaxDynamicSharedMemorySize, dd * 8);
    for (int it = 0; it < 2; ++it) k_visit<<<1, kPT, dd * 8>>>(lv, delta, phi, src, dd, out, 50);
    long long h[72]; cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    for (int l = 0; l < L; ++l) printf("level %2d n=%6d: visit(3 sweeps) %6lld  restrict_to(l+1) %6lld  prolong_from(l+1) %6lld cycles\n", l, lv.size[l] - 1, h[3 * l], h[3 * l + 1], h[3 * l + 2]);
    return 0;
}
