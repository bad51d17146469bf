// What does a dependent kernel boundary inside a CUDA graph cost on the B200, and what does a grid-wide barrier inside one
// persistent (cooperative) kernel cost? Decides whether a single-launch "megakernel" V-cycle could beat the 13 dependent
// launches of the 1025^2 cycle.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o launch_vs_gridsync launch_vs_gridsync.cu
#include <cooperative_groups.h>
#include <cstdio>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;

__global__ void tiny(double *p, int blocks_work)
{
    if (blockIdx.x < blocks_work && threadIdx.x == 0) p[blockIdx.x] += 1.0;
}
__global__ void persistent(double *p, int phases)
{
    cg::grid_group g = cg::this_grid();
    for (int k = 0; k < phases; ++k) {
        if (threadIdx.x == 0) p[blockIdx.x] += 1.0;
        g.sync();
    }
}
// hand-rolled barrier: one atomic counter, generation flag
__global__ void persistent_atomic(double *p, int phases, unsigned int *counter, volatile unsigned int *gen)
{
    for (int k = 0; k < phases; ++k) {
        if (threadIdx.x == 0) p[blockIdx.x] += 1.0;
        __syncthreads();
        if (threadIdx.x == 0) {
            const unsigned int my = *gen;
            __threadfence();
            if (atomicAdd(counter, 1u) == gridDim.x - 1) { *counter = 0; __threadfence(); *gen = my + 1; }
            else while (*gen == my) { }
            __threadfence();
        }
        __syncthreads();
    }
}
int main()
{
    double *p; cudaMalloc(&p, 1 << 20); cudaMemset(p, 0, 1 << 20);
    unsigned int *ctr; cudaMalloc(&ctr, 8); cudaMemset(ctr, 0, 8);
    cudaStream_t st; cudaStreamCreate(&st);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int K = 13, reps = 200;
    for (int blocks : {16, 148, 592, 1184}) {
        // graph of K dependent tiny kernels
        cudaGraph_t g; cudaGraphExec_t ge;
        cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
        for (int k = 0; k < K; ++k) tiny<<<blocks, 256, 0, st>>>(p, blocks);
        cudaStreamEndCapture(st, &g);
        cudaGraphInstantiate(&ge, g, 0);
        for (int r = 0; r < 20; ++r) cudaGraphLaunch(ge, st);
        cudaStreamSynchronize(st);
        cudaEventRecord(e0, st);
        for (int r = 0; r < reps; ++r) cudaGraphLaunch(ge, st);
        cudaEventRecord(e1, st);
        cudaStreamSynchronize(st);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double graph_us = ms * 1e3 / reps;
        double coop_us = -1, atom_us = -1;
        if (blocks <= 1184) {
            int phases = K;
            void *args[] = {&p, &phases};
            for (int r = 0; r < 5; ++r) cudaLaunchCooperativeKernel((void *)persistent, dim3(blocks), dim3(256), args, 0, st);
            cudaStreamSynchronize(st);
            cudaEventRecord(e0, st);
            for (int r = 0; r < reps; ++r) cudaLaunchCooperativeKernel((void *)persistent, dim3(blocks), dim3(256), args, 0, st);
            cudaEventRecord(e1, st);
            cudaStreamSynchronize(st);
            if (cudaGetLastError() == cudaSuccess) { cudaEventElapsedTime(&ms, e0, e1); coop_us = ms * 1e3 / reps; }
            unsigned int *gen = ctr + 1;
            void *args2[] = {&p, &phases, &ctr, &gen};
            for (int r = 0; r < 5; ++r) cudaLaunchCooperativeKernel((void *)persistent_atomic, dim3(blocks), dim3(256), args2, 0, st);
            cudaStreamSynchronize(st);
            cudaEventRecord(e0, st);
            for (int r = 0; r < reps; ++r) cudaLaunchCooperativeKernel((void *)persistent_atomic, dim3(blocks), dim3(256), args2, 0, st);
            cudaEventRecord(e1, st);
            cudaStreamSynchronize(st);
            if (cudaGetLastError() == cudaSuccess) { cudaEventElapsedTime(&ms, e0, e1); atom_us = ms * 1e3 / reps; }
        }
        printf("%5d blocks: graph of %d dependent kernels %.2f us (%.2f per kernel) | cooperative kernel with %d grid.sync %.2f us (%.2f per phase) | "
               "hand-rolled atomic barrier %.2f us (%.2f per phase)\n", blocks, K, graph_us, graph_us / K, K, coop_us, coop_us / K, atom_us, atom_us / K);
        cudaGraphExecDestroy(ge); cudaGraphDestroy(g);
    }
    return 0;
}
