// micro-benchmark: dependent-issue latencies on sm_100a that bound the Poisson sweeps
//   dfma chain, shfl(64-bit)+dfma chain, lds(64-bit) pointer chase, bar.sync round with 16 warps, global (L2) load chase
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_lat(double* out, long long* cyc, int iters, const int* chase_g)
{
    __shared__ int chase[1024];
    __shared__ double sd[1024];
    const int lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) { chase[i] = (i * 33 + 7) & 1023; sd[i] = i * 1e-3; }
    __syncthreads();
    double a = threadIdx.x * 1e-9 + 1.;
    const double m = 0.999999, b = 1e-7;
    long long t0, t1;
    // 1. dfma
    t0 = clock64();
    for (int i = 0; i < iters; ++i) a = fma(a, m, b);
    t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    // 2. shfl + dfma
    t0 = clock64();
    for (int i = 0; i < iters; ++i) a = fma(m, __shfl_up_sync(0xffffffffu, a, 1), a);
    t1 = clock64();
    if (threadIdx.x == 0) cyc[1] = t1 - t0;
    // 3. lds chase (32-bit index -> 64-bit value add)
    int p = lane;
    t0 = clock64();
    for (int i = 0; i < iters; ++i) p = chase[p];
    t1 = clock64();
    if (threadIdx.x == 0) cyc[2] = t1 - t0;
    a += p;
    // 4. lds f64 dependent: address from value
    t0 = clock64();
    for (int i = 0; i < iters; ++i) { a = sd[((int)a) & 1023] + 1.; }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[3] = t1 - t0;
    // 5. bar.sync
    t0 = clock64();
    for (int i = 0; i < iters; ++i) { __syncthreads(); }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[4] = t1 - t0;
    // 6. sts + bar + lds (cross-warp exchange)
    t0 = clock64();
    for (int i = 0; i < iters; ++i) { if (lane == 31) sd[threadIdx.x >> 5] = a; __syncthreads(); a += sd[((threadIdx.x >> 5) + 15) & 15]; __syncthreads(); }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[5] = t1 - t0;
    // 7. global (L2) chase
    p = lane;
    t0 = clock64();
    for (int i = 0; i < 256; ++i) p = chase_g[p];
    t1 = clock64();
    if (threadIdx.x == 0) cyc[6] = (t1 - t0) * (iters / 256);
    a += p;
    out[threadIdx.x] = a;
}
int main()
{
    double* out; long long* cyc; int* cg; cudaMalloc(&out, 8 * 1024); cudaMalloc(&cyc, 64); cudaMalloc(&cg, 4 << 20);
    int* h = new int[1 << 20]; for (int i = 0; i < (1 << 20); ++i) h[i] = (int)(((long long)i * 40503 + 12345) & ((1 << 20) - 1));
    cudaMemcpy(cg, h, 4 << 20, cudaMemcpyHostToDevice);
    const int iters = 2048;
    const char* names[7] = { "dfma", "shfl64+dfma", "lds32 chase", "cvt+lds64+dadd", "bar.sync", "sts+bar+lds+dadd+bar", "ldg(L2) chase" };
    for (int warps = 1; warps <= 16; warps *= 4) {
        k_lat<<<1, 32 * warps>>>(out, cyc, iters, cg);
        k_lat<<<1, 32 * warps>>>(out, cyc, iters, cg);
        long long c[8]; cudaMemcpy(c, cyc, 56, cudaMemcpyDeviceToHost);
        printf("warps=%d:", warps);
        for (int i = 0; i < 7; ++i) printf("  %s %.1f", names[i], (double)c[i] / iters);
        printf("\n");
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
