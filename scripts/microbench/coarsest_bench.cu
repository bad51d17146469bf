// Cycles per sweep of the in-warp coarsest Jacobi solve on a 5x5 grid (the default coarse_solve_size), one warp:
//   (a) warp_coarsest_jacobi<1>  -- the shared-memory version (still used for coarsest grids with more than 32 points)
//   (b) warp_coarsest_reg<false> -- the register-resident version with speculative 8-sweep batches (production for <= 32 points)
//   (c) the bare dependent chain of RegSweep::sweep without any exit test (lower bound)
// Also prints the clock stamps at the start / end of each batch's sweeps (B2S_COARSEST_STAMPS) and checks that (a) and
// (b) agree bit for bit in the fields and in the sweep count (the returned sums differ by reduction order only).
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -std=c++17 -I finalprojectrepo.jl_b200/csrc \
//              -o scripts/microbench/coarsest_bench scripts/microbench/coarsest_bench.cu
#include <cstdio>
#define B2S_COARSEST_STAMPS 1
#include "multigrid2d_kernels.cuh"
using namespace b2s;

__device__ __noinline__ double pure_sweeps(double *u, const double *rhs, int nx, int ny, double h, double c, long long *cyc)
{
    const int lane = threadIdx.x & 31, n = nx * ny;
    const bool valid = lane < n;
    const int j = lane / nx, i = lane - j * nx;
    RegSweep<false> sw;
    sw.interior = valid && i >= 1 && i <= nx - 2 && j >= 1 && j <= ny - 2;
    sw.red = ((i + j) & 1) == 0;
    sw.up = lane + nx; sw.dn = lane - nx;
    sw.f = valid ? rhs[lane] : 0.0;
    sw.exact = true;
    const Coef k = make_coef(h, c, 4.0 / 5.0);
    sw.C = k.C; sw.s2 = k._h2; sw.kw = k.w;
    double v = valid ? u[lane] : 0.0;
    double a[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tot = 0.0;
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < 32; ++it) {
#pragma unroll
        for (int b = 0; b < 8; ++b) v = sw.sweep(v, a[b]);
        tot += a[0] + a[1] + a[2] + a[3] + a[4] + a[5] + a[6] + a[7];
    }
    const long long t1 = clock64();
    if (lane == 0) cyc[0] = t1 - t0;
    if (valid) u[lane] = v;
    return tot;
}

__global__ void __launch_bounds__(1024) kern(double *out, long long *cyc, int *sw, double sstar, int iters)
{
    __shared__ double u[32], f[32], t[32];
    const int lane = threadIdx.x;
    const Coef k = make_coef(0.25, 0.0, 0.8);
    for (int rep = 0; rep < 4; ++rep) {  // the last repetition is reported (warm instruction cache)
        u[lane] = 0.0; t[lane] = 0.0; f[lane] = 1.0 + 0.01 * lane;
        __syncwarp();
        long long t0 = clock64();
        double r = warp_coarsest_jacobi<1>(u, f, t, 5, 5, k, sstar, iters, sw);
        long long t1 = clock64();
        if (lane == 0) { cyc[0] = t1 - t0; out[0] = r; }
        if (lane < 25) out[32 + lane] = u[lane];
        __syncwarp();
        u[lane] = 0.0; t[lane] = 0.0;
        __syncwarp();
        t0 = clock64();
        r = warp_coarsest_reg<false>(u, f, 5, 5, 0.25, 0.0, sstar, iters, sw + 1);
        t1 = clock64();
        if (lane == 0) {
            cyc[1] = t1 - t0; out[1] = r;
            for (int q = 0; q < 16; ++q) cyc[8 + q] = g_stamps[q] ? g_stamps[q] - t0 : 0;
        }
        if (lane < 25) out[64 + lane] = u[lane];
        __syncwarp();
        out[2] = pure_sweeps(u, f, 5, 5, 0.25, 0.0, cyc + 2);
        __syncwarp();
    }
}

int main()
{
    double *out; long long *cyc; int *sw;
    cudaMallocManaged(&out, 128 * 8); cudaMallocManaged(&cyc, 32 * 8); cudaMallocManaged(&sw, 64);
    const double sstars[3] = {0.0, 1e-9, 0.0};
    const int its[3] = {64, 100, 8};
    for (int mode = 0; mode < 3; ++mode) {
        kern<<<1, 32>>>(out, cyc, sw, sstars[mode], its[mode]);
        const cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        int same = sw[0] == sw[1];  // the returned sums differ in the last bits only (reduction order)
        for (int i = 0; i < 25; ++i) same &= (out[32 + i] == out[64 + i]);
        printf("cap %3d, S* %g: shared-memory version %lld cycles / %d sweeps = %.1f per sweep | register version %lld cycles / %d sweeps"
               " = %.1f per sweep | bare chain %.1f per sweep | fields and sweep counts identical: %d, sums differ by %.1e (relative)\n",
               its[mode], sstars[mode], cyc[0], sw[0], (double)cyc[0] / sw[0], cyc[1], sw[1], (double)cyc[1] / sw[1], cyc[2] / 256.0, same,
               out[0] != 0.0 ? (out[1] - out[0]) / out[0] : 0.0);
        printf("  batch stamps (cycles since entry):");
        for (int q = 0; q < 16; ++q) printf(" %lld", cyc[8 + q]);
        printf("\n");
    }
    return 0;
}
