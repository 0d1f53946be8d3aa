// micro-benchmark: FP64 FMA dependent-issue latency and single-warp throughput on sm_100a
#include <cstdio>
#include <cuda_runtime.h>
template <int CHAINS>
__global__ void k(double* out, long long* cyc, int iters)
{
    double a[CHAINS];
    for (int c = 0; c < CHAINS; ++c) a[c] = threadIdx.x * 1e-9 + c;
    const double m = 0.999999, b = 1e-7;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) a[c] = fma(a[c], m, b);
    }
    long long t1 = clock64();
    double s = 0; for (int c = 0; c < CHAINS; ++c) s += a[c];
    out[threadIdx.x] = s;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
template <int CHAINS> void run(const char* name)
{
    double* out; long long* cyc; cudaMalloc(&out, 8 * 1024); cudaMalloc(&cyc, 8);
    const int iters = 4096;
    for (int warps = 1; warps <= 8; warps *= 2) {
        k<CHAINS><<<1, 32 * warps>>>(out, cyc, iters);
        k<CHAINS><<<1, 32 * warps>>>(out, cyc, iters);
        long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("%s chains=%d warps=%d: %.2f cycles per iteration (= %.2f per DFMA)\n", name, CHAINS, warps, (double)h / iters, (double)h / iters / CHAINS);
    }
}
int main() { run<1>("dfma"); run<2>("dfma"); run<4>("dfma"); run<8>("dfma"); return 0; }
