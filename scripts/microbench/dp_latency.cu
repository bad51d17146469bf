// Dependent-chain latencies on the B200 (cycles per op): DADD, DMUL, DFMA, shared-memory load-use, warp shuffle of a double.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o dp_latency dp_latency.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double *out, long long *cyc, double a, double b)
{
    __shared__ double sm[64];
    const int lane = threadIdx.x;
    sm[lane] = a + lane; sm[lane + 32] = b;
    __syncwarp();
    double x = a + lane;
    long long t0, t1;
    const int N = 512;
    t0 = clock64();
#pragma unroll 64
    for (int i = 0; i < N; ++i) x = x + b;
    t1 = clock64();
    if (lane == 0) cyc[0] = (t1 - t0);
    double y = x;
    t0 = clock64();
#pragma unroll 64
    for (int i = 0; i < N; ++i) y = y * b;
    t1 = clock64();
    if (lane == 0) cyc[1] = (t1 - t0);
    double z = y;
    t0 = clock64();
#pragma unroll 64
    for (int i = 0; i < N; ++i) z = __fma_rn(z, b, a);
    t1 = clock64();
    if (lane == 0) cyc[2] = (t1 - t0);
    // shared-memory pointer chase (load-use latency incl. address arithmetic)
    int idx = lane;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) idx = ((volatile int *)sm)[idx & 63] & 63;
    t1 = clock64();
    if (lane == 0) cyc[3] = (t1 - t0);
    double w = z + idx;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) w = __shfl_xor_sync(0xffffffffu, w, 1);
    t1 = clock64();
    if (lane == 0) cyc[4] = (t1 - t0);
    // store -> syncwarp -> load round trip
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) {
        ((volatile double *)sm)[lane] = w;
        __syncwarp();
        w = ((volatile double *)sm)[lane ^ 1];
        __syncwarp();
    }
    t1 = clock64();
    if (lane == 0) cyc[5] = (t1 - t0);
    // two independent DADD chains (ILP 2)
    double p = w, q = w + 1.0;
    t0 = clock64();
#pragma unroll 64
    for (int i = 0; i < N; ++i) { p = p + b; q = q + a; }
    t1 = clock64();
    if (lane == 0) cyc[6] = (t1 - t0);
    // float add chain for comparison
    float f = (float)p;
    t0 = clock64();
#pragma unroll 64
    for (int i = 0; i < N; ++i) f = f + (float)b;
    t1 = clock64();
    if (lane == 0) cyc[7] = (t1 - t0);
    out[lane] = p + q + f;
}
int main()
{
    double *out; long long *cyc;
    cudaMalloc(&out, 32 * 8); cudaMallocManaged(&cyc, 8 * 8);
    for (int rep = 0; rep < 2; ++rep) { k<<<1, 32>>>(out, cyc, 1.000001, 0.999999); cudaDeviceSynchronize(); }
    const char *names[] = {"DADD", "DMUL", "DFMA", "LDS chase", "SHFL f64", "STS+sync+LDS+sync", "2xDADD (per pair)", "FADD"};
    for (int i = 0; i < 8; ++i) printf("%-20s %.2f cycles/op\n", names[i], cyc[i] / 512.0);
    return 0;
}
