// Feasibility: does programmatic dependent launch shorten the node-to-node latency of tiny dependent kernels inside a
// captured graph (plain graph and WHILE-conditional body)?   nvcc -arch=sm_100a -o pdl_test pdl_test.cu
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); return 1; } } while (0)

__global__ void k_step(double *a, const double *b, int n, int pdl)
{
    if (pdl) asm volatile("griddepcontrol.wait;" ::: "memory");
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] = b[i] * 0.5 + a[i == 0 ? n - 1 : i - 1] * 0.25 + 1.0;
}
__global__ void k_cond(cudaGraphConditionalHandle h, int *counter, int limit)
{
    if (threadIdx.x == 0) { int c = ++(*counter); cudaGraphSetConditional(h, c < limit ? 1u : 0u); }
}

static cudaError_t launch(double *a, double *b, int n, int blocks, cudaStream_t s, int pdl)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(blocks); cfg.blockDim = dim3(128); cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, k_step, a, (const double *)b, n, pdl);
}

int main()
{
    const int n = 4096, blocks = 32, chain = 200;
    double *a, *b; int *counter;
    CK(cudaMalloc(&a, n * 8)); CK(cudaMalloc(&b, n * 8)); CK(cudaMalloc(&counter, 4));
    CK(cudaMemset(a, 0, n * 8)); CK(cudaMemset(b, 0, n * 8));
    cudaStream_t s; CK(cudaStreamCreate(&s));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int pdl = 0; pdl < 2; ++pdl) {
        // (1) plain captured graph
        cudaGraph_t g; cudaGraphExec_t ge;
        CK(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
        for (int i = 0; i < chain; ++i) CK(launch(i & 1 ? a : b, i & 1 ? b : a, n, blocks, s, pdl));
        CK(cudaStreamEndCapture(s, &g));
        CK(cudaGraphInstantiate(&ge, g, 0));
        CK(cudaGraphLaunch(ge, s)); CK(cudaStreamSynchronize(s));
        CK(cudaEventRecord(e0, s));
        for (int r = 0; r < 10; ++r) CK(cudaGraphLaunch(ge, s));
        CK(cudaEventRecord(e1, s)); CK(cudaStreamSynchronize(s));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("pdl=%d plain graph: %.3f us per node\n", pdl, ms * 1e3 / (10 * chain));
        cudaGraphExecDestroy(ge); cudaGraphDestroy(g);
        // (2) WHILE body
        cudaGraph_t gw; CK(cudaGraphCreate(&gw, 0));
        cudaGraphConditionalHandle h; CK(cudaGraphConditionalHandleCreate(&h, gw, 1, cudaGraphCondAssignDefault));
        cudaGraphNodeParams cp = {cudaGraphNodeTypeConditional};
        cp.conditional.handle = h; cp.conditional.type = cudaGraphCondTypeWhile; cp.conditional.size = 1;
        cudaGraphNode_t wn; CK(cudaGraphAddNode(&wn, gw, nullptr, 0, &cp));
        cudaGraph_t body = cp.conditional.phGraph_out[0];
        CK(cudaStreamBeginCaptureToGraph(s, body, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
        for (int i = 0; i < chain; ++i) CK(launch(i & 1 ? a : b, i & 1 ? b : a, n, blocks, s, pdl));
        k_cond<<<1, 32, 0, s>>>(h, counter, 20);
        CK(cudaStreamEndCapture(s, nullptr));
        cudaError_t ie = cudaGraphInstantiate(&ge, gw, 0);
        if (ie != cudaSuccess) { printf("pdl=%d WHILE body: instantiate failed: %s\n", pdl, cudaGetErrorString(ie)); cudaGetLastError(); continue; }
        CK(cudaMemsetAsync(counter, 0, 4, s)); CK(cudaGraphLaunch(ge, s)); CK(cudaStreamSynchronize(s));
        CK(cudaMemsetAsync(counter, 0, 4, s));
        CK(cudaEventRecord(e0, s));
        CK(cudaGraphLaunch(ge, s));
        CK(cudaEventRecord(e1, s)); CK(cudaStreamSynchronize(s));
        CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("pdl=%d WHILE body (20 iterations): %.3f us per node\n", pdl, ms * 1e3 / (20 * (chain + 1)));
        cudaGraphExecDestroy(ge); cudaGraphDestroy(gw);
    }
    return 0;
}
