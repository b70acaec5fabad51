#include <cuda_runtime.h>
#include <stdio.h>
__global__ void decide(cudaGraphConditionalHandle h, int* counter, int max_iter) {
    int c = ++(*counter);
    cudaGraphSetConditional(h, c < max_iter ? 1u : 0u);
}
__global__ void body(int* x) { atomicAdd(x, 1); }
int main() {
    cudaGraph_t g; cudaGraphCreate(&g, 0);
    cudaGraphConditionalHandle h;
    cudaGraphConditionalHandleCreate(&h, g, 1, cudaGraphCondAssignDefault);
    cudaGraphNodeParams p = {};
    p.type = cudaGraphNodeTypeConditional;
    p.conditional.handle = h; p.conditional.type = cudaGraphCondTypeWhile; p.conditional.size = 1;
    cudaGraphNode_t node;
    cudaError_t e = cudaGraphAddNode(&node, g, nullptr, 0, &p);
    printf("add node: %s\n", cudaGetErrorString(e));
    cudaGraph_t bodyg = p.conditional.phGraph_out[0];
    cudaStream_t s; cudaStreamCreate(&s);
    int *x, *c; cudaMalloc(&x, 4); cudaMalloc(&c, 4); cudaMemset(x, 0, 4); cudaMemset(c, 0, 4);
    e = cudaStreamBeginCaptureToGraph(s, bodyg, nullptr, nullptr, 0, cudaStreamCaptureModeRelaxed);
    printf("begin capture: %s\n", cudaGetErrorString(e));
    body<<<1, 1, 0, s>>>(x);
    decide<<<1, 1, 0, s>>>(h, c, 7);
    cudaGraph_t out; e = cudaStreamEndCapture(s, &out);
    printf("end capture: %s\n", cudaGetErrorString(e));
    cudaGraphExec_t ex; e = cudaGraphInstantiate(&ex, g, 0);
    printf("instantiate: %s\n", cudaGetErrorString(e));
    cudaGraphLaunch(ex, s); cudaStreamSynchronize(s);
    int hx; cudaMemcpy(&hx, x, 4, cudaMemcpyDeviceToHost);
    printf("x = %d (expect 7)\n", hx);
    return 0;
}
