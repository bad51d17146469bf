// Cycles per sweep of the in-warp coarsest Jacobi solve (5x5 grid, 64 sweeps, exit test never true).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -std=c++17 -I finalprojectrepo.jl_b200/csrc -o coarsest_bench coarsest_bench.cu
#include <cstdio>
#define B2S_COARSEST_STAMPS 1
#include "multigrid2d_kernels.cuh"
using namespace b2s;

constexpr int kCoarsestBatch = 8;
// register-resident variant: one point per lane, neighbours by shuffle
__device__ __noinline__ double reg_coarsest_jacobi(double *u, const double *rhs, int nx, int ny, const Coef &k, double sstar, int iters,
                                                   int *sweeps_out)
{
    const int lane = threadIdx.x & 31, n = nx * ny;
    const bool valid = lane < n;
    const int j = lane / nx, i = lane - j * nx;
    const bool interior = valid && i >= 1 && i <= nx - 2 && j >= 1 && j <= ny - 2;
    double v = valid ? u[lane] : 0.0;
    const double f = valid ? rhs[lane] : 0.0;
    double tot = 0.0;
    int s = 0;
    for (;;) {
        const double snap = v;
        const int nb = min(kCoarsestBatch, iters - s);
        double acc[kCoarsestBatch];
#pragma unroll
        for (int b = 0; b < kCoarsestBatch; ++b) {
            acc[b] = 0.0;
            if (b < nb) {
                const double e = __shfl_down_sync(0xffffffffu, v, 1), w = __shfl_up_sync(0xffffffffu, v, 1);
                const double nn = __shfl_sync(0xffffffffu, v, lane + nx), ss = __shfl_sync(0xffffffffu, v, lane - nx);
                if (interior) {
                    const double r = ((e + w + nn + ss - k.C * v) * k._h2 - f);
                    acc[b] = r * r;
                    v = v + k.w * r;
                }
            }
        }
#pragma unroll
        for (int b = 0; b < kCoarsestBatch; ++b) acc[b] = warp_sum(acc[b]);
        int hit = -1;
#pragma unroll
        for (int b = kCoarsestBatch - 1; b >= 0; --b)
            if (b < nb && acc[b] < sstar) hit = b;
        if (hit < 0 && s + nb < iters) { s += nb; continue; }
        if (hit < 0) hit = nb - 1;
#pragma unroll
        for (int b = 0; b < kCoarsestBatch; ++b)
            if (b == hit) tot = acc[b];
        if (hit != nb - 1) {
            v = snap;
            for (int b = 0; b <= hit; ++b) {
                const double e = __shfl_down_sync(0xffffffffu, v, 1), w = __shfl_up_sync(0xffffffffu, v, 1);
                const double nn = __shfl_sync(0xffffffffu, v, lane + nx), ss = __shfl_sync(0xffffffffu, v, lane - nx);
                if (interior) {
                    const double r = ((e + w + nn + ss - k.C * v) * k._h2 - f);
                    v = v + k.w * r;
                }
            }
        }
        s += hit + 1;
        break;
    }
    if (valid) u[lane] = v;
    __syncwarp();
    if (sweeps_out != nullptr && lane == 0) *sweeps_out = s;
    return tot;
}


// variant C: branch-free sweeps, coefficients by value, full batches of 8 with a transpose-reduction (9 f64 shuffles
// per batch instead of 40), tail batches by the generic path
__device__ __forceinline__ double sweep_reg(double v, double f, bool interior, int up, int dn, double C, double _h2, double kw,
                                            double &acc)
{
    const double e = __shfl_down_sync(0xffffffffu, v, 1), w = __shfl_up_sync(0xffffffffu, v, 1);
    const double nn = __shfl_sync(0xffffffffu, v, up), ss = __shfl_sync(0xffffffffu, v, dn);
    const double r = ((e + w + nn + ss - C * v) * _h2 - f);
    const double vn = v + kw * r;
    acc = interior ? r * r : 0.0;
    return interior ? vn : v;
}
__device__ __noinline__ double reg2_coarsest_jacobi(double *u, const double *rhs, int nx, int ny, double C, double _h2, double kw,
                                                    double sstar, int iters, int *sweeps_out)
{
    const int lane = threadIdx.x & 31, n = nx * ny;
    const bool valid = lane < n;
    const int j = lane / nx, i = lane - j * nx;
    const bool interior = valid && i >= 1 && i <= nx - 2 && j >= 1 && j <= ny - 2;
    const int up = lane + nx, dn = lane - nx;
    double v = valid ? u[lane] : 0.0;
    const double f = valid ? rhs[lane] : 0.0;
    double tot = 0.0;
    int s = 0;
    for (;;) {
        const double snap = v;
        const int nb = min(8, iters - s);
        double a0 = 0, a1 = 0, a2 = 0, a3 = 0, a4 = 0, a5 = 0, a6 = 0, a7 = 0;
        if (nb == 8) {
            v = sweep_reg(v, f, interior, up, dn, C, _h2, kw, a0);
            v = sweep_reg(v, f, interior, up, dn, C, _h2, kw, a1);
            v = sweep_reg(v, f, interior, up, dn, C, _h2, kw, a2);
            v = sweep_reg(v, f, interior, up, dn, C, _h2, kw, a3);
            v = sweep_reg(v, f, interior, up, dn, C, _h2, kw, a4);
            v = sweep_reg(v, f, interior, up, dn, C, _h2, kw, a5);
            v = sweep_reg(v, f, interior, up, dn, C, _h2, kw, a6);
            v = sweep_reg(v, f, interior, up, dn, C, _h2, kw, a7);
        } else {
            double a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            for (int b = 0; b < nb; ++b) v = sweep_reg(v, f, interior, up, dn, C, _h2, kw, a[b]);
            a0 = a[0]; a1 = a[1]; a2 = a[2]; a3 = a[3]; a4 = a[4]; a5 = a[5]; a6 = a[6]; a7 = a[7];
        }
        // transpose-reduction: after the three halving steps lane l holds the partial sum of sweep ((l>>2)&1)*4 + ((l>>3)&1)*2 + ((l>>4)&1)
        {
            const bool hi16 = lane & 16, hi8 = lane & 8, hi4 = lane & 4;
            double s0 = hi16 ? a0 : a1, k0 = hi16 ? a1 : a0;  // send s*, keep k*
            double s1 = hi16 ? a2 : a3, k1 = hi16 ? a3 : a2;
            double s2 = hi16 ? a4 : a5, k2 = hi16 ? a5 : a4;
            double s3 = hi16 ? a6 : a7, k3 = hi16 ? a7 : a6;
            k0 += __shfl_xor_sync(0xffffffffu, s0, 16); k1 += __shfl_xor_sync(0xffffffffu, s1, 16);
            k2 += __shfl_xor_sync(0xffffffffu, s2, 16); k3 += __shfl_xor_sync(0xffffffffu, s3, 16);
            // k0: sweep 0|1, k1: 2|3, k2: 4|5, k3: 6|7 (second index on lanes with bit 16)
            double t0 = hi8 ? k0 : k1, m0 = hi8 ? k1 : k0;
            double t1 = hi8 ? k2 : k3, m1 = hi8 ? k3 : k2;
            m0 += __shfl_xor_sync(0xffffffffu, t0, 8); m1 += __shfl_xor_sync(0xffffffffu, t1, 8);
            double t2 = hi4 ? m0 : m1, m2 = hi4 ? m1 : m0;
            m2 += __shfl_xor_sync(0xffffffffu, t2, 4);
            m2 += __shfl_xor_sync(0xffffffffu, m2, 2);
            m2 += __shfl_xor_sync(0xffffffffu, m2, 1);
            // lane l now holds the total of sweep b(l) = 4*bit2 + 2*bit3 + bit4
            const int myb = ((lane >> 2) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 4) & 1);
            const unsigned ok = __ballot_sync(0xffffffffu, myb < nb && m2 < sstar);
            // first sweep b in 0..nb-1 whose total satisfies the test
            int hit = -1;
#pragma unroll
            for (int b = 7; b >= 0; --b) {
                const int src = ((b >> 2) & 1) * 4 + ((b >> 1) & 1) * 8 + (b & 1) * 16;  // a lane that holds sweep b
                if ((ok >> src) & 1u) hit = b;
            }
            if (hit < 0 && s + nb < iters) { s += nb; continue; }
            if (hit < 0) hit = nb - 1;
            const int src = ((hit >> 2) & 1) * 4 + ((hit >> 1) & 1) * 8 + (hit & 1) * 16;
            tot = __shfl_sync(0xffffffffu, m2, src);
            if (hit != nb - 1) {
                v = snap;
                double dummy;
                for (int b = 0; b <= hit; ++b) v = sweep_reg(v, f, interior, up, dn, C, _h2, kw, dummy);
            }
            s += hit + 1;
        }
        break;
    }
    if (valid) u[lane] = v;
    __syncwarp();
    if (sweeps_out != nullptr && lane == 0) *sweeps_out = s;
    return tot;
}

// pure sweep rate of RegSweep (no reductions): 32 batches of 8 unrolled sweeps
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
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < 32; ++it) {
#pragma unroll
        for (int b = 0; b < 8; ++b) v = sw.sweep(v, a[b]);
        tot += a[0] + a[1] + a[2] + a[3] + a[4] + a[5] + a[6] + a[7];
    }
    long long t1 = clock64();
    if (lane == 0) cyc[0] = t1 - t0;
    // same without keeping the per-sweep sums
    t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < 32; ++it) {
#pragma unroll
        for (int b = 0; b < 8; ++b) { double d; v = sw.sweep(v, d); }
    }
    t1 = clock64();
    if (lane == 0) cyc[1] = t1 - t0;
    if (valid) u[lane] = v;
    return tot;
}

__global__ void __launch_bounds__(1024) kern(double *out, long long *cyc, int *sw, double sstar, int iters)
{
    __shared__ double u[32], f[32], t[32];
    const int lane = threadIdx.x;
    const Coef k = make_coef(0.25, 0.0, 0.8);
    for (int rep = 0; rep < 6; ++rep) {
        u[lane] = 0.0; t[lane] = 0.0; f[lane] = 1.0 + 0.01 * lane;
        __syncwarp();
        long long t0 = clock64();
        double r = warp_coarsest_jacobi<1>(u, f, t, 5, 5, k, sstar, iters, sw);
        long long t1 = clock64();
        if (lane == 0) { cyc[0] = t1 - t0; cyc[8 + rep] = t1 - t0; out[0] = r; }
        if (lane < 25) out[32 + lane] = u[lane];
        __syncwarp();
        u[lane] = 0.0; t[lane] = 0.0;
        __syncwarp();
        t0 = clock64();
        r = reg_coarsest_jacobi(u, f, 5, 5, k, sstar, iters, sw + 1);
        t1 = clock64();
        if (lane == 0) { cyc[1] = t1 - t0; cyc[16 + rep] = t1 - t0; out[1] = r; }
        if (lane < 25) out[64 + lane] = u[lane];
        __syncwarp();
        u[lane] = 0.0; t[lane] = 0.0;
        __syncwarp();
        t0 = clock64();
        r = warp_coarsest_reg<false>(u, f, 5, 5, 0.25, 0.0, sstar, iters, sw + 2);
        t1 = clock64();
        if (lane == 0) { cyc[2] = t1 - t0; cyc[24 + rep] = t1 - t0; out[2] = r; for (int q = 0; q < 20; ++q) cyc[40 + q] = g_stamps[q] - t0; }
        if (lane < 25) out[96 + lane] = u[lane];
        __syncwarp();
        out[3] = pure_sweeps(u, f, 5, 5, 0.25, 0.0, cyc + 4);
        __syncwarp();
    }
}
int main()
{
    double *out; long long *cyc; int *sw;
    cudaMallocManaged(&out, 128 * 8); cudaMallocManaged(&cyc, 64 * 8); cudaMallocManaged(&sw, 64);
    for (int mode = 0; mode < 1; ++mode) {
        const double sstar = mode != 1 ? 0.0 : 1e-9;
        const int iters = mode == 0 ? 64 : mode == 1 ? 100 : 8 * (mode - 1);
        kern<<<1, 32>>>(out, cyc, sw, sstar, iters);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        for (int v = 0; v < 3; ++v) { printf("  version %d cycles per rep:", v); for (int r = 0; r < 6; ++r) printf(" %lld", cyc[8 + 8 * v + r]); printf("\n"); }
        printf("  pure sweeps: %.1f cycles/sweep with sums, %.1f without\n", cyc[4] / 256.0, cyc[5] / 256.0);
        printf("  stamps:"); for (int q = 0; q < 20; ++q) printf(" %lld", cyc[40 + q]); printf("\n");
        int same = 1;
        for (int i = 0; i < 25; ++i) same &= (out[32 + i] == out[64 + i]) && (out[32 + i] == out[96 + i]);
        printf("production warp_coarsest_reg: %lld cycles, %d sweeps (%.1f/sweep), ss %.17g\n", cyc[2], sw[2], (double)cyc[2] / sw[2], out[2]);
        printf("mode %d: smem version %lld cycles, %d sweeps (%.1f/sweep), ss %.17g | register version %lld cycles, %d sweeps (%.1f/sweep), ss %.17g | fields identical: %d\n",
               mode, cyc[0], sw[0], (double)cyc[0] / sw[0], out[0], cyc[1], sw[1], (double)cyc[1] / sw[1], out[1], same);
    }
    return 0;
}
