// ns2d.cu -- the Navier-Stokes step around the multigrid solves (streamfunction-vorticity Boussinesq, explicit or
// semi-implicit): the caller of hot path 2, fused into two kernels.
//
// Reference: navier_stokes_2D, scripts-part2/part2.jl:140-262; kernels :90-137; compute_dt :76-87.
#include "common.cuh"

#include <math.h>

#include <algorithm>
#include <vector>

using namespace b2s;

cudaStream_t b2s_mg_stream_internal(b2s_mg *h);
void b2s_mg_count_launches_internal(b2s_mg *h, long long n);

namespace {

constexpr int kMGBX = 128;

__device__ __forceinline__ void atomic_max_nonneg(double *addr, double v)
{
    // non-negative doubles order like their bit patterns
    atomicMax((unsigned long long *)addr, (unsigned long long)__double_as_longlong(v));
}

// compute_velocity! (part2.jl:90-96) + the three maxima of compute_dt (:76-87): max v, max |vx|, max |vy|.
__global__ void __launch_bounds__(kMGBX) ns_velocity_kernel(const double *__restrict__ S, double hx, double hy,
                                                            double *__restrict__ vx, double *__restrict__ vy, int nx, int ny,
                                                            int rows, double *maxima)
{
    const int i = blockIdx.x * kMGBX + threadIdx.x;
    const int j0 = max(1, blockIdx.y * rows), j1 = min(blockIdx.y * rows + rows, ny - 1);
    double mv = 0.0, mx = 0.0, my = 0.0;
    if (i >= 1 && i <= nx - 2) {
        for (int j = j0; j < j1; ++j) {
            const size_t p = (size_t)i + (size_t)nx * j;
            const double a = (S[p + nx] - S[p - nx]) / (2 * hy);
            const double b = -(S[p + 1] - S[p - 1]) / (2 * hx);
            vx[p] = a;
            vy[p] = b;
            mv = fmax(mv, sqrt(a * a + b * b));
            mx = fmax(mx, fabs(a));
            my = fmax(my, fabs(b));
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mv = fmax(mv, __shfl_xor_sync(0xffffffffu, mv, o));
        mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        my = fmax(my, __shfl_xor_sync(0xffffffffu, my, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomic_max_nonneg(maxima + 0, mv);
        atomic_max_nonneg(maxima + 1, mx);
        atomic_max_nonneg(maxima + 2, my);
    }
}

struct TermsArgs {
    const double *T, *W, *vx, *vy;
    double *Ra_dTdx, *dT2, *dW2;  // kept for parity checks
    double *outT, *outW;          // beta > 0: T_rhs, W_rhs ; beta == 0: T_new, W_new
    int nx, ny, rows;
    double hx, hy, k, Pr, Ra, beta, dt, cT, cW;
    int diffusion;                // beta !~ 1
    int implicit;                 // beta > 0
};

// compute_Ra_dTdx!, compute_diffusion2d! x2, compute_advection2d_x!/_y! x2 (part2.jl:99-137) and the right-hand sides
// / explicit update (:219-230) in one pass. Frame points: every stencil term is 0 there, like the reference's arrays.
__global__ void __launch_bounds__(kMGBX) ns_terms_kernel(const TermsArgs a)
{
    const int nx = a.nx, ny = a.ny;
    const int i = blockIdx.x * kMGBX + threadIdx.x;
    if (i >= nx) return;
    const int j0 = blockIdx.y * a.rows, j1 = min(j0 + a.rows, ny);
    const double hx = a.hx, hy = a.hy;
    for (int j = j0; j < j1; ++j) {
        const size_t p = (size_t)i + (size_t)nx * j;
        const double *T = a.T, *W = a.W;
        double ra = 0.0, dT2 = 0.0, dW2 = 0.0, dTx = 0.0, dTy = 0.0, dWx = 0.0, dWy = 0.0;
        if (i >= 1 && i <= nx - 2 && j >= 1 && j <= ny - 2) {
            ra = a.Ra * (T[p + 1] - T[p - 1]) / (2 * hx);
            if (a.diffusion) {
                dT2 = a.k * ((T[p + 1] - 2 * T[p] + T[p - 1]) / (hx * hx) + (T[p + nx] - 2 * T[p] + T[p - nx]) / (hy * hy));
                dW2 = a.Pr * ((W[p + 1] - 2 * W[p] + W[p - 1]) / (hx * hx) + (W[p + nx] - 2 * W[p] + W[p - nx]) / (hy * hy));
            }
            const double vx = a.vx[p], vy = a.vy[p];
            dTx = vx > 0 ? vx * (T[p] - T[p - 1]) / hx : vx * (T[p + 1] - T[p]) / hx;
            dTy = vy > 0 ? vy * (T[p] - T[p - nx]) / hy : vy * (T[p + nx] - T[p]) / hy;
            dWx = vx > 0 ? vx * (W[p] - W[p - 1]) / hx : vx * (W[p + 1] - W[p]) / hx;
            dWy = vy > 0 ? vy * (W[p] - W[p - nx]) / hy : vy * (W[p + nx] - W[p]) / hy;
            a.Ra_dTdx[p] = ra;
            if (a.diffusion) { a.dT2[p] = dT2; a.dW2[p] = dW2; }
        }
        if (a.implicit) {
            a.outT[p] = -a.cT * (T[p] + a.dt * (((1.0 - a.beta) * dT2 - dTx) - dTy));
            a.outW[p] = -a.cW * (W[p] + a.dt * ((((1.0 - a.beta) * dW2 - dWx) - dWy) - a.Pr * ra));
        } else {
            a.outT[p] = T[p] + a.dt * ((dT2 - dTx) - dTy);
            a.outW[p] = W[p] + a.dt * (((dW2 - dWx) - dWy) - a.Pr * ra);
        }
    }
}

}  // namespace

struct b2s_ns2d {
    b2s_ns2d_params p;
    b2s_mg *mg = nullptr;
    int device = 0;
    int solver = B2S_NS_SOLVER_VCYCLE;
    double *T = nullptr, *W = nullptr, *S = nullptr, *vx = nullptr, *vy = nullptr, *Ra_dTdx = nullptr, *dT2 = nullptr,
           *dW2 = nullptr, *bufA = nullptr, *bufB = nullptr, *maxima = nullptr, *maxima_pin = nullptr;
};

extern "C" {

int b2s_ns2d_destroy(b2s_ns2d *h)
{
    if (!h) return B2S_OK;
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(h->device);
    if (h->mg) b2s_mg_destroy(h->mg);
    double *arrs[] = {h->T, h->W, h->S, h->vx, h->vy, h->Ra_dTdx, h->dT2, h->dW2, h->bufA, h->bufB, h->maxima};
    for (double *a : arrs)
        if (a) cudaFree(a);
    if (h->maxima_pin) cudaFreeHost(h->maxima_pin);
    delete h;
    if (prev >= 0) cudaSetDevice(prev);
    return B2S_OK;
}

int b2s_ns2d_create(b2s_ns2d **out, const b2s_ns2d_params *p, const b2s_mg_config *mgc)
{
    B2S_REQUIRE(out && p, B2S_ERR_BAD_ARG, "NULL argument");
    *out = nullptr;
    b2s_mg_config c;
    if (mgc) c = *mgc;
    else { c = b2s_mg_config(); c.coarse_solve_size = 5; c.use_graph = 1; c.smem_levels = 1; c.fuse_sweeps = 1; }  // MGOpt() multigrid.jl:21
    c.nx = p->nx; c.ny = p->ny;
    b2s_ns2d *h = new b2s_ns2d();
    h->p = *p;
    h->device = c.device;
    int rc = b2s_mg_create(&h->mg, &c);
    if (rc != B2S_OK) { delete h; return rc; }
    DeviceGuard guard;
    guard.set(c.device);
    const size_t bytes = (size_t)p->nx * p->ny * sizeof(double);
    double **arrs[] = {&h->T, &h->W, &h->S, &h->vx, &h->vy, &h->Ra_dTdx, &h->dT2, &h->dW2, &h->bufA, &h->bufB};
    for (double **a : arrs) {
        if (cudaMalloc(a, bytes) != cudaSuccess || cudaMemset(*a, 0, bytes) != cudaSuccess) {
            set_error("allocation of %zu bytes failed: %s", bytes, cudaGetErrorString(cudaGetLastError()));
            b2s_ns2d_destroy(h);
            return B2S_ERR_CUDA;
        }
    }
    if (cudaMalloc(&h->maxima, 4 * sizeof(double)) != cudaSuccess || cudaMallocHost(&h->maxima_pin, 4 * sizeof(double)) != cudaSuccess) {
        set_error("allocation failed");
        b2s_ns2d_destroy(h);
        return B2S_ERR_CUDA;
    }
    *out = h;
    return B2S_OK;
}

int b2s_ns2d_set_solver(b2s_ns2d *h, int solver)
{
    B2S_REQUIRE(h, B2S_ERR_BAD_ARG, "NULL handle");
    B2S_REQUIRE(solver == B2S_NS_SOLVER_VCYCLE || solver == B2S_NS_SOLVER_MG_PCG, B2S_ERR_BAD_ARG, "unknown solver %d", solver);
    h->solver = solver;
    return B2S_OK;
}

static double *ns_field(b2s_ns2d *h, int which) { return which == 0 ? h->T : (which == 1 ? h->W : (which == 2 ? h->S : nullptr)); }

int b2s_ns2d_set_field(b2s_ns2d *h, int which, const double *host)
{
    B2S_REQUIRE(h && host && ns_field(h, which), B2S_ERR_BAD_ARG, "bad argument");
    DeviceGuard guard;
    guard.set(h->device);
    B2S_CUDA(cudaStreamSynchronize(b2s_mg_stream_internal(h->mg)));
    B2S_CUDA(cudaMemcpy(ns_field(h, which), host, (size_t)h->p.nx * h->p.ny * sizeof(double), cudaMemcpyHostToDevice));
    return B2S_OK;
}

int b2s_ns2d_get_field(b2s_ns2d *h, int which, double *host)
{
    B2S_REQUIRE(h && host && ns_field(h, which), B2S_ERR_BAD_ARG, "bad argument");
    DeviceGuard guard;
    guard.set(h->device);
    B2S_CUDA(cudaStreamSynchronize(b2s_mg_stream_internal(h->mg)));
    B2S_CUDA(cudaMemcpy(host, ns_field(h, which), (size_t)h->p.nx * h->p.ny * sizeof(double), cudaMemcpyDeviceToHost));
    return B2S_OK;
}

int b2s_ns2d_get_aux(b2s_ns2d *h, int which, double *host)
{
    B2S_REQUIRE(h && host && which >= 0 && which <= 4, B2S_ERR_BAD_ARG, "bad argument");
    double *src[] = {h->vx, h->vy, h->Ra_dTdx, h->dT2, h->dW2};
    DeviceGuard guard;
    guard.set(h->device);
    B2S_CUDA(cudaStreamSynchronize(b2s_mg_stream_internal(h->mg)));
    B2S_CUDA(cudaMemcpy(host, src[which], (size_t)h->p.nx * h->p.ny * sizeof(double), cudaMemcpyDeviceToHost));
    return B2S_OK;
}

int b2s_ns2d_init_cosine(b2s_ns2d *h, int which)
{
    B2S_REQUIRE(h && ns_field(h, which), B2S_ERR_BAD_ARG, "bad argument");
    // init_array!(M, cosine, h, width)  part2.jl:58-63, evaluated on the host like the reference
    const int nx = h->p.nx, ny = h->p.ny;
    const double hh = 1.0 / (ny - 1.0), width = (nx - 1.0) / (ny - 1.0);
    std::vector<double> M((size_t)nx * ny);
    for (int i = 0; i < nx; ++i) {
        const double v = 0.5 * (1.0 + cos((3.0 * M_PI * (double)i * hh) / width));
        for (int j = 0; j < ny; ++j) M[(size_t)i + (size_t)nx * j] = v;
    }
    return b2s_ns2d_set_field(h, which, M.data());
}

int b2s_ns2d_step(b2s_ns2d *h, b2s_ns2d_stepinfo *info)
{
    B2S_REQUIRE(h, B2S_ERR_BAD_ARG, "NULL handle");
    DeviceGuard guard;
    guard.set(h->device);
    const b2s_ns2d_params &P = h->p;
    const int nx = P.nx, ny = P.ny;
    const double hh = 1.0 / (ny - 1.0), hx = hh, hy = hh;
    const double dt_dif = (P.a_dif * (hh * hh)) / fmax(P.k, P.Pr);
    cudaStream_t st = b2s_mg_stream_internal(h->mg);
    b2s_ns2d_stepinfo inf = {};
    // D S = W, Dirichlet 0                                                               part2.jl:187
    const bool pcg = h->solver == B2S_NS_SOLVER_MG_PCG;  // S and W: Dirichlet solves; T (BCs inside the cycle) always cycles
    if (pcg) B2S_CHECK(b2s_mg_pcg_solve2(h->mg, h->S, h->W, hh, 0.0, P.tol, P.niters, B2S_PCG_TOL_RHS, &inf.r_S, &inf.cycles_S));
    else B2S_CHECK(b2s_mg_solve(h->mg, h->S, h->W, hh, 0.0, P.tol, P.niters, 0, &inf.r_S, &inf.cycles_S, nullptr));
    const int bx = (nx + kMGBX - 1) / kMGBX;
    const int want = std::max(1, (148 * 16) / bx);
    const int rows = std::max(4, (ny + want - 1) / want);
    dim3 grid(bx, (ny + rows - 1) / rows, 1);
    B2S_CUDA(cudaMemsetAsync(h->maxima, 0, 4 * sizeof(double), st));
    ns_velocity_kernel<<<grid, kMGBX, 0, st>>>(h->S, hx, hy, h->vx, h->vy, nx, ny, rows, h->maxima);
    B2S_CUDA(cudaGetLastError());
    B2S_CUDA(cudaMemcpyAsync(h->maxima_pin, h->maxima, 3 * sizeof(double), cudaMemcpyDeviceToHost, st));
    // apply_boundary_conditions!(T)                                                      part2.jl:199
    B2S_CHECK(b2s_apply_bc2d(h->T, nx, ny, 0, st));
    B2S_CUDA(cudaStreamSynchronize(st));
    double dt;
    if (h->maxima_pin[0] == 0) dt = dt_dif;  // compute_dt, part2.jl:76-87
    else {
        const double dt_adv = P.a_adv * fmin(hh / h->maxima_pin[1], hh / h->maxima_pin[2]);
        dt = (P.beta >= 0.5 ? dt_adv : fmin(dt_dif, dt_adv));
    }
    inf.dt = dt;
    const bool beta_is_one = fabs(P.beta - 1.0) <= 1.4901161193847656e-08 * fmax(fabs(P.beta), 1.0);  // isapprox
    TermsArgs a = {};
    a.T = h->T; a.W = h->W; a.vx = h->vx; a.vy = h->vy; a.Ra_dTdx = h->Ra_dTdx; a.dT2 = h->dT2; a.dW2 = h->dW2;
    a.outT = h->bufA; a.outW = h->bufB; a.nx = nx; a.ny = ny; a.rows = rows;
    a.hx = hx; a.hy = hy; a.k = P.k; a.Pr = P.Pr; a.Ra = P.Ra; a.beta = P.beta; a.dt = dt;
    a.diffusion = !beta_is_one; a.implicit = P.beta > 0.0;
    double cT = 0.0, cW = 0.0;
    if (a.implicit) { cT = 1.0 / (P.beta * dt); cW = cT / P.Pr; }
    a.cT = cT; a.cW = cW;
    ns_terms_kernel<<<grid, kMGBX, 0, st>>>(a);
    B2S_CUDA(cudaGetLastError());
    b2s_mg_count_launches_internal(h->mg, 3);
    if (a.implicit) {
        B2S_CHECK(b2s_mg_solve(h->mg, h->T, h->bufA, hh, cT, P.tol, P.niters, 1, &inf.r_T, &inf.cycles_T, nullptr));  // :221
        if (pcg) B2S_CHECK(b2s_mg_pcg_solve2(h->mg, h->W, h->bufB, hh, cW, P.tol, P.niters, B2S_PCG_TOL_RHS, &inf.r_W, &inf.cycles_W));
        else B2S_CHECK(b2s_mg_solve(h->mg, h->W, h->bufB, hh, cW, P.tol, P.niters, 0, &inf.r_W, &inf.cycles_W, nullptr));  // :226
    } else {
        std::swap(h->T, h->bufA);  // T .= T + dt*(...), W .= W + dt*(...)   :229-230
        std::swap(h->W, h->bufB);
        B2S_CUDA(cudaStreamSynchronize(st));
    }
    if (info) *info = inf;
    return B2S_OK;
}

}  // extern "C"
