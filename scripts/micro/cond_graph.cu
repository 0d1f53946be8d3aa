#include <cuda_runtime.h>
#include <cstdio>
__global__ void body(int* c) { if (threadIdx.x == 0) atomicAdd(c, 1); }
__global__ void cond(cudaGraphConditionalHandle h, const int* c, int lim) { cudaGraphSetConditional(h, *c < lim ? 1u : 0u); }
int main() {
    int* d; cudaMalloc(&d, 4); cudaMemset(d, 0, 4);
    cudaStream_t st; cudaStreamCreate(&st);
    cudaGraph_t g; cudaGraphCreate(&g, 0);
    cudaGraphConditionalHandle h; cudaGraphConditionalHandleCreate(&h, g, 1, cudaGraphCondAssignDefault);
    cudaGraphNodeParams p = {}; p.type = cudaGraphNodeTypeConditional; p.conditional.handle = h; p.conditional.type = cudaGraphCondTypeWhile; p.conditional.size = 1;
    cudaGraphNode_t node; cudaError_t e = cudaGraphAddNode(&node, g, nullptr, 0, &p);
    printf("addnode %s\n", cudaGetErrorString(e));
    cudaGraph_t bodyg = p.conditional.phGraph_out[0];
    e = cudaStreamBeginCaptureToGraph(st, bodyg, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal);
    printf("begin %s\n", cudaGetErrorString(e));
    body<<<4, 32, 0, st>>>(d);
    cond<<<1, 1, 0, st>>>(h, d, 40);
    e = cudaStreamEndCapture(st, nullptr); printf("end %s\n", cudaGetErrorString(e));
    cudaGraphExec_t ex; e = cudaGraphInstantiate(&ex, g, 0); printf("inst %s\n", cudaGetErrorString(e));
    e = cudaGraphLaunch(ex, st); cudaStreamSynchronize(st);
    int hv = 0; cudaMemcpy(&hv, d, 4, cudaMemcpyDeviceToHost);
    printf("launch %s count %d (expect 40)\n", cudaGetErrorString(e), hv);
    return 0;
}
