// Where do the cycles of one in-register Jacobi sweep go? Dependent chains on one warp of a B200.
#include <cstdio>
#include <cuda_runtime.h>
#define FULL 0xffffffffu
__global__ void k(double *out, long long *cyc, double C, double h, double f, double w, int nx)
{
    const int lane = threadIdx.x;
    const int N = 256;
    const bool interior = (lane % 5) >= 1 && (lane % 5) <= 3 && lane >= 5 && lane < 20;
    const int up = lane + nx, dn = lane - nx;
    double x = 1.0 + lane * 0.001;
    long long t0, t1;
    // T1: 8 dependent DP ops
    t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < N; ++i) { const double r = ((x + f + f + f - C * x) * h - f); x = x + w * r; }
    t1 = clock64();
    if (lane == 0) cyc[0] = t1 - t0;
    // T2: + 4 f64 shuffles
    t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < N; ++i) {
        const double e = __shfl_down_sync(FULL, x, 1), ww = __shfl_up_sync(FULL, x, 1), nn = __shfl_sync(FULL, x, up), ss = __shfl_sync(FULL, x, dn);
        const double r = ((e + ww + nn + ss - C * x) * h - f);
        x = x + w * r;
    }
    t1 = clock64();
    if (lane == 0) cyc[1] = t1 - t0;
    // T3: + select
    t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < N; ++i) {
        const double e = __shfl_down_sync(FULL, x, 1), ww = __shfl_up_sync(FULL, x, 1), nn = __shfl_sync(FULL, x, up), ss = __shfl_sync(FULL, x, dn);
        const double r = ((e + ww + nn + ss - C * x) * h - f);
        const double xn = x + w * r;
        x = interior ? xn : x;
    }
    t1 = clock64();
    if (lane == 0) cyc[2] = t1 - t0;
    // T4: 1 shuffle + 8 ops
    t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < N; ++i) {
        const double e = __shfl_down_sync(FULL, x, 1);
        const double r = ((e + f + f + f - C * x) * h - f);
        x = x + w * r;
    }
    t1 = clock64();
    if (lane == 0) cyc[3] = t1 - t0;
    // T5: shared-memory exchange instead of shuffles (value stays in a register)
    __shared__ double sm[2][64];
    sm[0][lane] = 0; sm[0][lane + 32] = 0; sm[1][lane] = 0; sm[1][lane + 32] = 0;
    __syncwarp();
    t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < N; ++i) {
        double *b = sm[i & 1] + 8;
        b[lane] = x;
        __syncwarp();
        const double r = ((b[lane + 1] + b[lane - 1] + b[lane + nx] + b[lane - nx] - C * x) * h - f);
        const double xn = x + w * r;
        x = interior ? xn : x;
    }
    t1 = clock64();
    if (lane == 0) cyc[4] = t1 - t0;
    // T6: chain of DADD -> DMUL alternating
    t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < N; ++i) { x = x + f; x = x * h; x = x + f; x = x * h; x = x + f; x = x * h; x = x + f; x = x * h; }
    t1 = clock64();
    if (lane == 0) cyc[5] = t1 - t0;
    out[lane] = x;
}
int main()
{
    double *out; long long *cyc;
    cudaMalloc(&out, 32 * 8); cudaMallocManaged(&cyc, 8 * 8);
    for (int rep = 0; rep < 2; ++rep) { k<<<1, 32>>>(out, cyc, 4.0, 1.0000001, 1e-3, 0.2, 5); cudaDeviceSynchronize(); }
    const char *names[] = {"T1 8 DP ops", "T2 +4 shfl64", "T3 +select", "T4 1 shfl64 + 8 ops", "T5 smem exchange + select", "T6 8 alternating DADD/DMUL"};
    for (int i = 0; i < 6; ++i) printf("%-32s %.1f cycles/iter\n", names[i], cyc[i] / 256.0);
    return 0;
}
