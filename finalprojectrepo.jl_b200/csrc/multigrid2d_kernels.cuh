// multigrid2d_kernels.cuh -- device code of hot path 2: matrix-free geometric multigrid for (lap - c) u = f on a
// 2-D node grid, its smoothers, transfer operators, the collapsed shared-memory coarse hierarchy and CG.
//
// Reference semantics (file:line relative to the reference repository):
//   scripts-part2/multigrid.jl:173-188   residual_2DPoisson!       res = ((uE+uW+uN+uS - C u)*_h2 - f), interior
//   scripts-part2/multigrid.jl:245-258   iteration_2DPoisson!      r_rms (pre-update) ; u += alpha*(h^2/C)*res
//   scripts-part2/multigrid.jl:330-358   restrict! / wrapper       zero, injection at even (0-based) points, Neumann
//   scripts-part2/multigrid.jl:403-472   prolongate* / wrapper     zero, bilinear scatter, Neumann  (here: a gather
//                                                                  that reproduces the CPU arrival order, no atomics)
//   scripts-part2/multigrid.jl:91-170    Vcycle_2DPoisson!
//   scripts-part2/krylov.jl:7-13,55-91   matvec, cg!
//   scripts-part2/part2_utils.jl:21-39   boundary conditions
// All arithmetic in the order written there; the translation unit is compiled with -fmad=false.
#pragma once
#include "common.cuh"

#include <cooperative_groups.h>

namespace b2s {

constexpr int kMaxLevels = 24;

// Per-level constants of a call, computed once on the host (plain IEEE arithmetic, no contraction: the same values the
// kernels used to derive per thread with two FP64 divisions) and uploaded together with the call arguments.
struct LevelCoef {
    double h;        // h * 2^level, doubled level by level (multigrid.jl:133)
    double C;        // 4 + c h^2
    double _h2;      // 1 / h^2
    double wJ;       // damped-Jacobi weight (4/5) (h^2 / C)   multigrid.jl:245-258
    double h2;       // h^2
    double wGS;      // Gauss-Seidel weight 1.0 (h^2 / C)      multigrid.jl:286
    double inv_h2;   // exact reciprocal of h^2 when it is a power of two (DivH2), else 0
    int exact, pad;
};

// Per-call arguments live in device memory so that a captured CUDA graph of the V-cycle is independent of them.
struct MGCall {
    double *u;          // finest-level unknown (caller's array)
    const double *rhs;  // finest-level right-hand side (caller's array)
    double h, c, tol;
    int apply_bcs;      // Neumann handling inside restrict/prolongate (multigrid.jl:355-357,468-470)
    int bc_before;      // apply_boundary_conditions!(u) before the cycle (multigrid.jl:60-62)
    double *sumsq;      // [0]: sum res^2 of the last sweep on the finest level; [1]: sum f^2
    int *coarse_sweeps; // sweeps / iterations of the coarsest solve
    // device-resident MGsolve loop (multigrid.jl:58-76): cycles enqueued after `done` is raised return at once
    int done;           // every kernel of a cycle reads this together with the scalars above (same cache line)
    int ncycles;        // V-cycles completed
    int niters;         // cap
    int check;          // 1: evaluate r_rms < tol*f_rms after every cycle (MGsolve); 0: fixed number of cycles
    double n_points;    // nx*ny of the finest level
    double *hist;       // r_rms / f_rms per cycle (capacity kMaxHist)
    int rb_combine;     // 1: sumsq[0] = sumsq[2] + sumsq[3] (red + black) before the test
    int pad;
    const LevelCoef *lev;  // [kMaxLevels], device memory right behind this struct
};
constexpr int kMaxHist = 4096;

// End of a V-cycle: r_rms, convergence test and bookkeeping on the device (multigrid.jl:64-75). One thread. The fused
// upward kernels of the finest level run it in the block that completes the residual norm (no separate launch).
__device__ __forceinline__ void cycle_end(MGCall *cp);
__global__ void mg_cycle_end_kernel(MGCall *cp)
{
    if (threadIdx.x != 0 || blockIdx.x != 0 || cp->done) return;
    cycle_end(cp);
}
__device__ __forceinline__ void cycle_end(MGCall *cp)
{
    double *ss = cp->sumsq;
    if (cp->rb_combine) ss[0] = ss[2] + ss[3];
    const double r_rms = sqrt(ss[0] / cp->n_points);
    const double f_rms = sqrt(ss[1] / cp->n_points);
    const int k = cp->ncycles;
    if (k < kMaxHist) cp->hist[k] = r_rms / f_rms;
    cp->ncycles = k + 1;
    const double tolf = cp->tol * f_rms;
    if ((cp->check && (r_rms < tolf || r_rms != r_rms)) || k + 1 >= cp->niters) cp->done = 1;
}

__device__ __forceinline__ double level_h(const MGCall *cp, int level)
{
    double h = cp->h;
    for (int l = 0; l < level; ++l) h = h * 2;  // Vcycle(..., h*2, ...) multigrid.jl:133
    return h;
}

struct Coef {
    double C, _h2, w;
};
__device__ __forceinline__ Coef make_coef(double h, double c, double alpha)
{
    Coef k;
    k.C = 4.0 + c * (h * h);
    k._h2 = 1 / (h * h);
    k.w = alpha * ((h * h) / (4.0 + c * (h * h)));
    return k;
}

// The Gauss-Seidel residual is written with "/ h^2" in the reference (multigrid.jl:279-283). When h^2 is a power of two
// (every grid with h = 1/2^k: all the configured shapes) x / h^2 and x * (1/h^2) are the same correctly rounded value, so
// the ~20-instruction FP64 division is replaced by one multiplication without changing a bit; any other h divides.
// The per-level constants sit right behind the call block in device memory (set_call uploads both with one copy): the
// address does not depend on a loaded pointer, so this load and the one of cp->done are issued together.
__device__ __forceinline__ const LevelCoef *level_consts(const MGCall *cp, int level)
{
    return reinterpret_cast<const LevelCoef *>(cp + 1) + level;
}
__device__ __forceinline__ Coef level_coef(const MGCall *cp, int level)
{
    const LevelCoef *L = level_consts(cp, level);
    Coef k;
    k.C = L->C; k._h2 = L->_h2; k.w = L->wJ;
    return k;
}

struct DivH2 {
    double h2, inv;
    bool exact;
};
__device__ __forceinline__ DivH2 make_div_h2(double h2)
{
    DivH2 d;
    d.h2 = h2;
    const long long b = __double_as_longlong(h2);
    const int e = (int)((b >> 52) & 0x7ff);
    d.exact = b > 0 && (b & 0x000fffffffffffffLL) == 0 && e >= 2 && e <= 2044;  // +2^k, and 1/h2 is normal too
    d.inv = d.exact ? 1.0 / h2 : 0.0;
    return d;
}
__device__ __forceinline__ double div_h2(double x, const DivH2 &d) { return d.exact ? x * d.inv : x / d.h2; }

__device__ __forceinline__ double point_residual(const double *__restrict__ u, const double *__restrict__ f, int nx, size_t p,
                                                 const Coef &k)
{
    return ((u[p + 1] + u[p - 1] + u[p + nx] + u[p - nx] - k.C * u[p]) * k._h2 - f[p]);
}

// ---------------------------------------------------------------------------------------------------------------
// Fine-level (global memory) kernels. Thread block = 128 threads along x; each thread marches `rows` rows in y.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kMGBX = 128;

// Device-resident state of a coarsest-level Jacobi solve that runs in global memory (level too large for shared memory)
struct CoarseLoop {
    double sstar;   // exit threshold on sum res^2 (exit_threshold)
    int sweeps, done, iters, pad;
};

struct SweepArgs {
    const MGCall *cp;
    int level;          // 0: arrays come from *cp
    const double *u;    // used when level > 0 or cp == nullptr
    const double *rhs;
    double *out;        // Jacobi: u_new ; residual: res
    int nx, ny, rows;
    double h, c, alpha; // used when cp == nullptr (L0 calls)
    int mode;           // 0: residual only (interior of out) ; 1: Jacobi sweep out = u + w*res (frame copied)
    int want_norm;
    double *partials;
    unsigned int *ticket;
    double *sumsq_out;
    int swap_io;        // level 0 ping-pong: 1 -> read tmp (u arg), write cp->u
    CoarseLoop *loop;   // nullable: global-memory coarsest solve -- skip when done, run the exit test in the last block
};

__global__ void __launch_bounds__(kMGBX) mg_sweep_kernel(const SweepArgs a)
{
    __shared__ double red[32];
    if (a.loop != nullptr && a.loop->done) return;
    if (a.cp != nullptr && a.cp->done) return;
    const double *u = a.u;
    const double *rhs = a.rhs;
    double *out = a.out;
    double h = a.h, c = a.c;
    if (a.cp != nullptr) {
        c = a.cp->c;
        h = level_h(a.cp, a.level);
        if (a.level == 0) {
            rhs = a.cp->rhs;
            if (a.swap_io) out = a.cp->u; else u = a.cp->u;
        }
    }
    const Coef k = make_coef(h, c, a.alpha);
    const int nx = a.nx, ny = a.ny;
    const int i = blockIdx.x * kMGBX + threadIdx.x;
    const int j0 = blockIdx.y * a.rows;
    const int j1 = min(j0 + a.rows, ny);
    double acc = 0.0;
    if (i < nx) {
        const bool xin = i >= 1 && i <= nx - 2;
        for (int j = j0; j < j1; ++j) {
            const size_t p = (size_t)i + (size_t)nx * j;
            const bool interior = xin && j >= 1 && j <= ny - 2;
            if (interior) {
                const double r = point_residual(u, rhs, nx, p, k);
                if (a.want_norm) acc += r * r;
                out[p] = a.mode == 0 ? r : u[p] + k.w * r;
            } else if (a.mode == 1) {
                out[p] = u[p];  // u += w*res with res == 0 on the frame
            }
        }
    }
    if (a.want_norm) {
        const int nblocks = gridDim.x * gridDim.y;
        const int bl = blockIdx.x + gridDim.x * blockIdx.y;
        const double bsum = block_sum(acc, red);
        double total;
        if (grid_sum_last_block(bsum, a.partials, a.ticket, nblocks, bl, red, &total)) {
            *a.sumsq_out = total;
            if (a.loop != nullptr) {  // res_rms < tol_rhs -> break   (multigrid.jl:150-156)
                const int sw = a.loop->sweeps + 1;
                a.loop->sweeps = sw;
                if (total < a.loop->sstar || sw >= a.loop->iters) a.loop->done = 1;
            }
        }
    }
}

// Red-black Gauss-Seidel half sweep in place (variant B). colour 0 = (i+j) even. Accumulates pre-update res^2 into
// sumsq_out[colour] when want_norm.
struct RbgsArgs {
    const MGCall *cp;
    int level;
    double *u;
    const double *rhs;
    int nx, ny, rows, colour;
    double h, c;
    int want_norm;
    double *partials;
    unsigned int *ticket;
    double *sumsq_out;
    const CoarseLoop *loop;  // nullable: global-memory coarsest solve -- skip when its exit test has fired
};

__global__ void __launch_bounds__(kMGBX) mg_rbgs_kernel(const RbgsArgs a)
{
    __shared__ double red[32];
    if (a.loop != nullptr && a.loop->done) return;
    if (a.cp != nullptr && a.cp->done) return;
    double *u = a.u;
    const double *rhs = a.rhs;
    double h = a.h, c = a.c;
    if (a.cp != nullptr) {
        c = a.cp->c;
        h = level_h(a.cp, a.level);
        if (a.level == 0) { u = a.cp->u; rhs = a.cp->rhs; }
    }
    const double C = 4.0 + c * (h * h), h2 = h * h, w = 1.0 * (h2 / C);
    const DivH2 dh = make_div_h2(h2);
    const int nx = a.nx, ny = a.ny;
    // each thread owns every second point of a row
    const int t = blockIdx.x * kMGBX + threadIdx.x;
    const int j0 = max(1, blockIdx.y * a.rows), j1 = min(blockIdx.y * a.rows + a.rows, ny - 1);
    double acc = 0.0;
    for (int j = j0; j < j1; ++j) {
        const int i = 1 + ((1 + j + a.colour) & 1) + 2 * t;
        if (i <= nx - 2) {
            const size_t p = (size_t)i + (size_t)nx * j;
            const double r = div_h2(u[p + 1] + u[p - 1] + u[p + nx] + u[p - nx] - C * u[p], dh) - rhs[p];
            u[p] = u[p] + w * r;
            acc += r * r;
        }
    }
    if (a.want_norm) {
        const int nblocks = gridDim.x * gridDim.y;
        const int bl = blockIdx.x + gridDim.x * blockIdx.y;
        const double bsum = block_sum(acc, red);
        double total;
        if (grid_sum_last_block(bsum, a.partials, a.ticket, nblocks, bl, red, &total)) a.sumsq_out[a.colour] = total;
    }
}

// Coarse right-hand side from the fine level: zero frame, injection or full weighting of the fine residual (FROM_RES:
// computed on the fly from u and rhs -- fused residual + restriction) or of a given fine field, Neumann copy.
struct RestrictArgs {
    const MGCall *cp;
    int level;             // fine level
    const double *u;       // fine unknown (FROM_RES) or the fine field to restrict
    const double *rhs;     // fine rhs (FROM_RES)
    double *coarse;
    double *zero_out;      // nullable: coarse-level unknown, reset to 0 in the same pass (corr_c .= 0, multigrid.jl:132)
    int nx, ny, nxc, nyc;
    double h, c;
    int from_res, full_weighting, apply_bcs;
};

template <bool FROM_RES>
__device__ __forceinline__ double fine_value(const double *__restrict__ u, const double *__restrict__ f, int nx, int i, int j,
                                             const Coef &k)
{
    const size_t p = (size_t)i + (size_t)nx * j;
    if (FROM_RES) return point_residual(u, f, nx, p, k);
    return u[p];
}

template <bool FROM_RES>
__device__ __forceinline__ double coarse_value(const double *__restrict__ u, const double *__restrict__ f, int nx, int I, int J,
                                               int nxc, int nyc, bool fw, const Coef &k)
{
    if (I < 1 || I > nxc - 2 || J < 1 || J > nyc - 2) return 0.0;
    const int i = 2 * I, j = 2 * J;
    if (!fw) return fine_value<FROM_RES>(u, f, nx, i, j, k);
    const double corners = (fine_value<FROM_RES>(u, f, nx, i - 1, j - 1, k) + fine_value<FROM_RES>(u, f, nx, i + 1, j - 1, k)) +
                           (fine_value<FROM_RES>(u, f, nx, i - 1, j + 1, k) + fine_value<FROM_RES>(u, f, nx, i + 1, j + 1, k));
    const double edges = (fine_value<FROM_RES>(u, f, nx, i - 1, j, k) + fine_value<FROM_RES>(u, f, nx, i + 1, j, k)) +
                         (fine_value<FROM_RES>(u, f, nx, i, j - 1, k) + fine_value<FROM_RES>(u, f, nx, i, j + 1, k));
    return ((corners + 2.0 * edges) + 4.0 * fine_value<FROM_RES>(u, f, nx, i, j, k)) * 0.0625;
}

__global__ void __launch_bounds__(256) mg_restrict_kernel(const RestrictArgs a)
{
    if (a.cp != nullptr && a.cp->done) return;
    const double *u = a.u;
    const double *rhs = a.rhs;
    double h = a.h, c = a.c;
    int apply_bcs = a.apply_bcs;
    if (a.cp != nullptr) {
        c = a.cp->c;
        h = level_h(a.cp, a.level);
        apply_bcs = a.cp->apply_bcs;
        if (a.level == 0) { u = a.cp->u; rhs = a.cp->rhs; }
    }
    const Coef k = make_coef(h, c, 1.0);
    const int I = blockIdx.x * 64 + threadIdx.x, J = blockIdx.y * 4 + threadIdx.y;
    if (I >= a.nxc || J >= a.nyc) return;
    int Is = I;
    if (apply_bcs) {  // coarse[0,:] = coarse[1,:]; coarse[nxc-1,:] = coarse[nxc-2,:]   part2_utils.jl:34-39
        if (I == 0) Is = 1;
        else if (I == a.nxc - 1) Is = a.nxc - 2;
    }
    const bool fw = a.full_weighting != 0;
    const double v = a.from_res ? coarse_value<true>(u, rhs, a.nx, Is, J, a.nxc, a.nyc, fw, k)
                                : coarse_value<false>(u, rhs, a.nx, Is, J, a.nxc, a.nyc, fw, k);
    a.coarse[(size_t)I + (size_t)a.nxc * J] = v;
    if (a.zero_out != nullptr) a.zero_out[(size_t)I + (size_t)a.nxc * J] = 0.0;
}

// Bilinear prolongation as a gather. ec's boundary ring is treated as zero (the reference scatters from interior
// coarse points only); the summation order per fine point is the arrival order of the reference's CPU loop.
__device__ __forceinline__ double coarse_at(const double *__restrict__ ec, int nxc, int nyc, int I, int J)
{
    return (I >= 1 && I <= nxc - 2 && J >= 1 && J <= nyc - 2) ? ec[(size_t)I + (size_t)nxc * J] : 0.0;
}
__device__ __forceinline__ double prolong_value(const double *__restrict__ ec, int nxc, int nyc, int i, int j)
{
    const int I = i >> 1, J = j >> 1;
    const bool io = i & 1, jo = j & 1;
    if (!io && !jo) return coarse_at(ec, nxc, nyc, I, J);
    if (io && !jo) return 0.5 * coarse_at(ec, nxc, nyc, I, J) + 0.5 * coarse_at(ec, nxc, nyc, I + 1, J);
    if (!io && jo) return 0.5 * coarse_at(ec, nxc, nyc, I, J) + 0.5 * coarse_at(ec, nxc, nyc, I, J + 1);
    return ((0.25 * coarse_at(ec, nxc, nyc, I, J) + 0.25 * coarse_at(ec, nxc, nyc, I + 1, J)) +
            0.25 * coarse_at(ec, nxc, nyc, I, J + 1)) + 0.25 * coarse_at(ec, nxc, nyc, I + 1, J + 1);
}
__device__ __forceinline__ double prolong_value_bc(const double *__restrict__ ec, int nxc, int nyc, int nx, int i, int j,
                                                   int apply_bcs)
{
    if (apply_bcs) {  // fine[0,:] = fine[1,:]; fine[nx-1,:] = fine[nx-2,:]
        if (i == 0) i = 1;
        else if (i == nx - 1) i = nx - 2;
    }
    return prolong_value(ec, nxc, nyc, i, j);
}

struct ProlongArgs {
    const MGCall *cp;
    int level;            // fine level
    const double *coarse;
    double *fine;         // mode 0: fine = P(coarse) ; mode 1: fine = fine - P(coarse)
    int nx, ny, nxc, nyc, rows;
    int mode, apply_bcs;
};

__global__ void __launch_bounds__(kMGBX) mg_prolong_kernel(const ProlongArgs a)
{
    if (a.cp != nullptr && a.cp->done) return;
    double *fine = a.fine;
    int apply_bcs = a.apply_bcs;
    if (a.cp != nullptr) {
        apply_bcs = a.cp->apply_bcs;
        if (a.level == 0) fine = a.cp->u;
    }
    const int i = blockIdx.x * kMGBX + threadIdx.x;
    if (i >= a.nx) return;
    const int j0 = blockIdx.y * a.rows, j1 = min(j0 + a.rows, a.ny);
    for (int j = j0; j < j1; ++j) {
        const size_t p = (size_t)i + (size_t)a.nx * j;
        const double e = prolong_value_bc(a.coarse, a.nxc, a.nyc, a.nx, i, j, apply_bcs);
        fine[p] = a.mode == 0 ? e : fine[p] - e;  // u_f .= u_f - corr_f   multigrid.jl:139
    }
}

// Fused prolongation + correction + first post-smoothing Jacobi sweep:
//   out = J(u - P(ec)),  and the corrected frame is carried over.  Same arithmetic as the three separate steps.
struct ProlongSmoothArgs {
    const MGCall *cp;
    int level;
    const double *coarse;
    const double *u;    // fine unknown before correction
    const double *rhs;
    double *out;
    int nx, ny, nxc, nyc, rows;
};

__global__ void __launch_bounds__(kMGBX) mg_prolong_smooth_kernel(const ProlongSmoothArgs a)
{
    const double *u = a.u;
    const double *rhs = a.rhs;
    const MGCall *cp = a.cp;
    if (cp->done) return;
    if (a.level == 0) { u = cp->u; rhs = cp->rhs; }
    const int apply_bcs = cp->apply_bcs;
    const Coef k = level_coef(cp, a.level);
    const int nx = a.nx, ny = a.ny, nxc = a.nxc, nyc = a.nyc;
    const int i = blockIdx.x * kMGBX + threadIdx.x;
    if (i >= nx) return;
    const int j0 = blockIdx.y * a.rows, j1 = min(j0 + a.rows, ny);
    auto corrected = [&](int ii, int jj) {
        return u[(size_t)ii + (size_t)nx * jj] - prolong_value_bc(a.coarse, nxc, nyc, nx, ii, jj, apply_bcs);
    };
    const bool xin = i >= 1 && i <= nx - 2;
    for (int j = j0; j < j1; ++j) {
        const size_t p = (size_t)i + (size_t)nx * j;
        const double uc = corrected(i, j);
        if (xin && j >= 1 && j <= ny - 2) {
            const double r = ((corrected(i + 1, j) + corrected(i - 1, j) + corrected(i, j + 1) + corrected(i, j - 1) - k.C * uc) *
                                  k._h2 - rhs[p]);
            a.out[p] = uc + k.w * r;
        } else {
            a.out[p] = uc;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Temporally blocked fine-level kernels (variant A: damped Jacobi + injection). A block stages its tile of u and rhs
// (plus a halo as wide as the number of fused steps) in shared memory and performs the WHOLE downward or upward part of
// one level there, so a level costs two passes over HBM/L2 instead of five kernels:
//   mg_down_kernel:  u_s = J(J(u)) ;  rc = inject(residual(u_s)) (+ Neumann) ;  ec = 0        [reads u, rhs; writes u_s, rc, ec]
//   mg_up_kernel:    u   = J(J(u_s - P(ec))) ; sum res^2 of the last sweep                    [reads u_s, rhs, ec; writes u]
// Every point value is produced by exactly the arithmetic of the unfused kernels (halo points are recomputed
// redundantly by neighbouring blocks), so results are bit-identical to them.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kTileThreads = 256;
// B2S_TILE_FG = 1 (experiment, OFF by default): the right-hand side is NOT staged in shared memory; every use reads it from
// global memory through the read-only path. Two arrays per tile instead of three: the 64x16 tiles of a 1025^2 level need
// 25-28 KB instead of 41 KB, so 7-8 blocks instead of 5 are resident per SM and the 1105 blocks of that level run in
// (almost) one wave instead of 1.5. Measured on B200 (profiles/r02_tile_rhs_global_ab.jsonl): bit-identical and 2.7 %
// SLOWER per 1025^2 V-cycle (0.0796 vs 0.0775 ms) -- the L2 latency of the rhs loads sits in every point's dependency
// chain and outweighs the extra resident blocks.
#ifndef B2S_TILE_FG
#define B2S_TILE_FG 0
#endif
#if B2S_TILE_FG
#define B2S_TILE_RHS(Fsm, s, rhsg, i, j, nx) __ldg((rhsg) + ((size_t)(i) + (size_t)(nx) * (size_t)(j)))
#else
#define B2S_TILE_RHS(Fsm, s, rhsg, i, j, nx) (Fsm)[s]
#endif
template <int TW, int TH>
struct TileCfg {
    static constexpr int kTW = TW, kTH = TH;
    static constexpr int kTP = TW + 8;  // shared-memory row pitch (tile + 2*3 halo, padded)
    static constexpr int kTRows = TH + 6;
    static constexpr int kCW = TW / 2 + 3, kCH = TH / 2 + 3;  // coarse window staged by the upward kernel
    static constexpr int kArrays = B2S_TILE_FG ? 2 : 3;  // u ping, u pong (, rhs)
    static constexpr size_t kSmemBytes = ((size_t)kArrays * kTP * kTRows + (size_t)kCW * kCH) * sizeof(double);
};

struct TileArgs {
    const MGCall *cp;
    int level;
    const double *u_in;   // down: u_l (level 0: cp->u) ; up: smoothed u_s (tmp_l)
    const double *rhs;    // level 0: cp->rhs
    double *u_out;        // down: tmp_l ; up: u_l (level 0: cp->u)
    double *rc;           // down: coarse rhs ; up: unused
    double *ec;           // down: coarse unknown to reset ; up: coarse correction (read)
    int nx, ny, nxc, nyc;
    int want_norm;
    double *partials;
    unsigned int *ticket;
    double *sumsq_out;
    int fused_end;        // up, finest level: the block that completes the norm also runs cycle_end()
};

// 8-byte asynchronous global->shared copy (LDGSTS); pred == false zero-fills the destination without reading.
__device__ __forceinline__ void cp_async8(double *smem_dst, const double *gmem_src, bool pred)
{
    const int src_bytes = pred ? 8 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(src_bytes)
                 : "memory");
}
__device__ __forceinline__ void cp_async_wait_all()
{
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

// one Jacobi sweep inside shared memory over a window of global coordinates. CHECKED = false: the whole window is
// known to lie in the interior of the domain (no per-point tests).
// B2S_TILE_YB = 2 (default; measured -1.3 % per 1025^2 V-cycle, profiles/r02_tile_yblock_ab.jsonl): a thread owns two
// vertically adjacent points per step -- the
// column values of rows r-1 .. r+2 are loaded once (4 loads) and shared by both stencils, the index arithmetic is paid
// once per pair, every shared-memory access stays unit-stride across the lanes (10 loads + 2 stores per 2 points
// instead of 12 + 2). Window heights are even for every tile shape. Each point is still the unfused kernels' arithmetic.
#ifndef B2S_TILE_YB
#define B2S_TILE_YB 2
#endif
template <int kTP, bool CHECKED>
__device__ __forceinline__ void tile_sweep(const double *__restrict__ src, const double *__restrict__ F, double *__restrict__ dst,
                                           int gx0, int gy0, int wx0, int wy0, int W, int H, int nx, int ny, const Coef &k,
                                           const double *__restrict__ rhs_g)
{
    (void)F; (void)rhs_g;
    // (gx0, gy0): global coordinates of shared-memory element (0,0); window origin (wx0, wy0), size W x H
    const int s0 = (wy0 - gy0) * kTP + (wx0 - gx0);
#if B2S_TILE_YB == 2
    for (int idx = threadIdx.x; idx < W * (H >> 1); idx += kTileThreads) {
        const int rp = idx / W, c = idx - rp * W;
        const int r = 2 * rp;
        const int s = s0 + r * kTP + c;
        const double a0 = src[s - kTP], a1 = src[s], a2 = src[s + kTP], a3 = src[s + 2 * kTP];
        double v0 = a1, v1 = a2;
        bool w0 = true, w1 = true, u0 = true, u1 = true;
        const int i = wx0 + c, j = wy0 + r;
        if (CHECKED) {
            const bool iin = i >= 0 && i < nx, iint = i >= 1 && i <= nx - 2;
            w0 = iin && j >= 0 && j < ny;
            w1 = iin && j + 1 >= 0 && j + 1 < ny;
            u0 = iint && j >= 1 && j <= ny - 2;
            u1 = iint && j + 1 >= 1 && j + 1 <= ny - 2;
        }
        if (u0) {
            const double res = ((src[s + 1] + src[s - 1] + a2 + a0 - k.C * a1) * k._h2 - B2S_TILE_RHS(F, s, rhs_g, i, j, nx));
            v0 = a1 + k.w * res;
        }
        if (u1) {
            const double res = ((src[s + kTP + 1] + src[s + kTP - 1] + a3 + a1 - k.C * a2) * k._h2 -
                                B2S_TILE_RHS(F, s + kTP, rhs_g, i, j + 1, nx));
            v1 = a2 + k.w * res;
        }
        if (w0) dst[s] = v0;
        if (w1) dst[s + kTP] = v1;
    }
#else
    for (int idx = threadIdx.x; idx < W * H; idx += kTileThreads) {
        const int r = idx / W, c = idx - r * W;
        const int s = s0 + r * kTP + c;
        double v = src[s];
        bool upd = true;
        const int i = wx0 + c, j = wy0 + r;
        if (CHECKED) {
            if (i < 0 || j < 0 || i >= nx || j >= ny) continue;
            upd = i >= 1 && i <= nx - 2 && j >= 1 && j <= ny - 2;
        }
        if (upd) {
            const double res = ((src[s + 1] + src[s - 1] + src[s + kTP] + src[s - kTP] - k.C * v) * k._h2 -
                                B2S_TILE_RHS(F, s, rhs_g, i, j, nx));
            v = v + k.w * res;
        }
        dst[s] = v;
    }
#endif
}

template <int TW, int TH>
__global__ void __launch_bounds__(kTileThreads) mg_down_kernel(const TileArgs a)
{
    using Cf = TileCfg<TW, TH>;
    constexpr int kTW = Cf::kTW, kTH = Cf::kTH, kTP = Cf::kTP, kTRows = Cf::kTRows, kCW = Cf::kCW, kCH = Cf::kCH;
    (void)kCW; (void)kCH; (void)kTRows;
    extern __shared__ __align__(16) double tsm[];
    double *A = tsm, *B = tsm + kTP * kTRows, *F = tsm + (Cf::kArrays - 1) * kTP * kTRows;  // F: only when staged
    const MGCall *cp = a.cp;
    // Levels below the finest get their array pointers from the launch arguments: their staging loads are issued before
    // the call block (done flag, constants) has arrived, which takes one L2 round trip off the critical path of these
    // latency-bound kernels; the loads of a cycle that turns out to be a no-op are harmless reads.
    const double *u = a.u_in, *rhs = a.rhs;
    if (a.level == 0) {
        if (cp->done) return;
        u = cp->u; rhs = cp->rhs;
    }
    const int nx = a.nx, ny = a.ny;
    const int X0 = blockIdx.x * kTW, Y0 = blockIdx.y * kTH;
    const int gx0 = X0 - 3, gy0 = Y0 - 3;
    // stage u on tile+3 and rhs on tile+2 (asynchronous copies: all loads of the block are in flight at once).
    // One warp per staged row, lanes along x: no per-element division, the row pointer and the row tests are hoisted.
    for (int r = threadIdx.x >> 5; r < kTRows; r += kTileThreads / 32) {
        const int j = gy0 + r;
        const bool jin = j >= 0 && j < ny;
        const size_t rowoff = jin ? (size_t)nx * j : 0;
        const bool frow = jin && r >= 1 && r < kTRows - 1;
        double *Ar = A + r * kTP, *Fr = F + r * kTP;
#pragma unroll
        for (int c = threadIdx.x & 31; c < kTW + 6; c += 32) {
            const int i = gx0 + c;
            const bool in = jin && i >= 0 && i < nx;
            const size_t p = in ? rowoff + i : 0;
            cp_async8(Ar + c, u + p, in);
#if !B2S_TILE_FG
            cp_async8(Fr + c, rhs + p, in && frow && c >= 1 && c < kTW + 5);
#endif
        }
    }
    const int done = a.level == 0 ? 0 : cp->done;
    const int apply_bcs = cp->apply_bcs;
    const Coef k = level_coef(cp, a.level);
    cp_async_wait_all();
    __syncthreads();
    if (done) return;
    // block-uniform: does the widest stencil window stay strictly inside the domain?
    const bool inner = X0 - 2 >= 1 && Y0 - 2 >= 1 && X0 + kTW + 1 <= nx - 2 && Y0 + kTH + 1 <= ny - 2;
    if (inner) {
        tile_sweep<kTP, false>(A, F, B, gx0, gy0, X0 - 2, Y0 - 2, kTW + 4, kTH + 4, nx, ny, k, rhs);
        __syncthreads();
        tile_sweep<kTP, false>(B, F, A, gx0, gy0, X0 - 1, Y0 - 1, kTW + 2, kTH + 2, nx, ny, k, rhs);
    } else {
        tile_sweep<kTP, true>(A, F, B, gx0, gy0, X0 - 2, Y0 - 2, kTW + 4, kTH + 4, nx, ny, k, rhs);
        __syncthreads();
        tile_sweep<kTP, true>(B, F, A, gx0, gy0, X0 - 1, Y0 - 1, kTW + 2, kTH + 2, nx, ny, k, rhs);
    }
    __syncthreads();
    // smoothed u out
    for (int idx = threadIdx.x; idx < kTW * kTH; idx += kTileThreads) {
        const int r = idx / kTW, c = idx - r * kTW;
        const int i = X0 + c, j = Y0 + r;
        if (i < nx && j < ny) a.u_out[(size_t)i + (size_t)nx * j] = A[(r + 3) * kTP + c + 3];
    }
    // coarse rhs = injected residual of the smoothed u (+ Neumann copies), coarse unknown = 0
    const int nxc = a.nxc, nyc = a.nyc;
    for (int idx = threadIdx.x; idx < (kTW / 2) * (kTH / 2); idx += kTileThreads) {
        const int r = idx / (kTW / 2), c = idx - r * (kTW / 2);
        const int I = X0 / 2 + c, J = Y0 / 2 + r;
        if (I >= nxc || J >= nyc) continue;
        const size_t pc = (size_t)I + (size_t)nxc * J;
        a.ec[pc] = 0.0;
        const bool interior = I >= 1 && I <= nxc - 2 && J >= 1 && J <= nyc - 2;
        if (interior) {
            const int s = (2 * r + 3) * kTP + 2 * c + 3;
            const double v = ((A[s + 1] + A[s - 1] + A[s + kTP] + A[s - kTP] - k.C * A[s]) * k._h2 -
                              B2S_TILE_RHS(F, s, rhs, 2 * I, 2 * J, nx));
            a.rc[pc] = v;
            if (apply_bcs) {  // coarse[0,:] = coarse[1,:] ; coarse[nxc-1,:] = coarse[nxc-2,:]
                if (I == 1) a.rc[(size_t)0 + (size_t)nxc * J] = v;
                if (I == nxc - 2) a.rc[(size_t)(nxc - 1) + (size_t)nxc * J] = v;
            }
        } else if (!(apply_bcs && (I == 0 || I == nxc - 1) && J >= 1 && J <= nyc - 2)) {
            a.rc[pc] = 0.0;
        }
    }
}

// bilinear prolongation from the staged coarse window (boundary-ring entries were staged as 0)
template <int kCW>
__device__ __forceinline__ double prolong_from_window(const double *__restrict__ Cw, int cx0, int cy0, int nx, int i, int j,
                                                      int apply_bcs)
{
    if (apply_bcs) {
        if (i == 0) i = 1;
        else if (i == nx - 1) i = nx - 2;
    }
    const int I = (i >> 1) - cx0, J = (j >> 1) - cy0;
    const double *q = Cw + J * kCW + I;
    const bool io = i & 1, jo = j & 1;
    if (!io && !jo) return q[0];
    if (io && !jo) return 0.5 * q[0] + 0.5 * q[1];
    if (!io && jo) return 0.5 * q[0] + 0.5 * q[kCW];
    return ((0.25 * q[0] + 0.25 * q[1]) + 0.25 * q[kCW]) + 0.25 * q[kCW + 1];
}

template <int TW, int TH>
__global__ void __launch_bounds__(kTileThreads) mg_up_kernel(const TileArgs a)
{
    using Cf = TileCfg<TW, TH>;
    constexpr int kTW = Cf::kTW, kTH = Cf::kTH, kTP = Cf::kTP, kTRows = Cf::kTRows, kCW = Cf::kCW, kCH = Cf::kCH;
    (void)kCW; (void)kCH; (void)kTRows;
    extern __shared__ __align__(16) double tsm[];
    __shared__ double red[32];
    double *A = tsm, *B = tsm + kTP * kTRows, *F = tsm + (Cf::kArrays - 1) * kTP * kTRows, *Cw = tsm + Cf::kArrays * kTP * kTRows;
    const MGCall *cp = a.cp;
    const double *rhs = a.rhs;
    double *out = a.u_out;
    if (a.level == 0) {  // finest level: the arrays are the caller's (see mg_down_kernel for the ordering below)
        if (cp->done) return;
        rhs = cp->rhs; out = cp->u;
    }
    const int nx = a.nx, ny = a.ny, nxc = a.nxc, nyc = a.nyc;
    const int X0 = blockIdx.x * kTW, Y0 = blockIdx.y * kTH;
    const int gx0 = X0 - 3, gy0 = Y0 - 3;
    const int cx0 = X0 / 2 - 1, cy0 = Y0 / 2 - 1;
    // stage the smoothed u on tile+2, rhs on tile+1 and the coarse correction window (one warp per row, lanes along x)
    for (int r = threadIdx.x >> 5; r < kTH + 4; r += kTileThreads / 32) {
        const int j = Y0 - 2 + r;
        const bool jin = j >= 0 && j < ny;
        const size_t rowoff = jin ? (size_t)nx * j : 0;
        const bool frow = jin && r >= 1 && r < kTH + 3;
        double *Ar = A + (r + 1) * kTP + 1, *Fr = F + (r + 1) * kTP + 1;
#pragma unroll
        for (int c = threadIdx.x & 31; c < kTW + 4; c += 32) {
            const int i = X0 - 2 + c;
            const bool in = jin && i >= 0 && i < nx;
            const size_t p = in ? rowoff + i : 0;
            cp_async8(Ar + c, a.u_in + p, in);
#if !B2S_TILE_FG
            cp_async8(Fr + c, rhs + p, in && frow && c >= 1 && c < kTW + 3);
#endif
        }
    }
    for (int r = threadIdx.x >> 5; r < kCH; r += kTileThreads / 32) {
        const int J = cy0 + r;
        const bool jin = J >= 1 && J <= nyc - 2;  // the boundary ring counts as 0
        const size_t rowoff = jin ? (size_t)nxc * J : 0;
#pragma unroll
        for (int c = threadIdx.x & 31; c < kCW; c += 32) {
            const int I = cx0 + c;
            const bool in = jin && I >= 1 && I <= nxc - 2;
            cp_async8(Cw + r * kCW + c, a.ec + (in ? rowoff + I : 0), in);
        }
    }
    const int done = a.level == 0 ? 0 : cp->done;
    const int apply_bcs = cp->apply_bcs;
    const Coef k = level_coef(cp, a.level);
    cp_async_wait_all();
    __syncthreads();
    if (done) return;
    // u_f .= u_f - corr_f on tile+2
    for (int idx = threadIdx.x; idx < (kTW + 4) * (kTH + 4); idx += kTileThreads) {
        const int r = idx / (kTW + 4), c = idx - r * (kTW + 4);
        const int i = X0 - 2 + c, j = Y0 - 2 + r;
        if (i < 0 || j < 0 || i >= nx || j >= ny) continue;
        const int s = (r + 1) * kTP + c + 1;
        A[s] = A[s] - prolong_from_window<kCW>(Cw, cx0, cy0, nx, i, j, apply_bcs);
    }
    __syncthreads();
    const bool inner = X0 - 1 >= 1 && Y0 - 1 >= 1 && X0 + kTW <= nx - 2 && Y0 + kTH <= ny - 2;
    if (inner) tile_sweep<kTP, false>(A, F, B, gx0, gy0, X0 - 1, Y0 - 1, kTW + 2, kTH + 2, nx, ny, k, rhs);
    else tile_sweep<kTP, true>(A, F, B, gx0, gy0, X0 - 1, Y0 - 1, kTW + 2, kTH + 2, nx, ny, k, rhs);
    __syncthreads();
    double acc = 0.0;
    for (int idx = threadIdx.x; idx < kTW * kTH; idx += kTileThreads) {
        const int r = idx / kTW, c = idx - r * kTW;
        const int i = X0 + c, j = Y0 + r;
        if (i >= nx || j >= ny) continue;
        const int s = (r + 3) * kTP + c + 3;
        double v = B[s];
        if (i >= 1 && i <= nx - 2 && j >= 1 && j <= ny - 2) {
            const double res = ((B[s + 1] + B[s - 1] + B[s + kTP] + B[s - kTP] - k.C * v) * k._h2 - B2S_TILE_RHS(F, s, rhs, i, j, nx));
            acc += res * res;
            v = v + k.w * res;
        }
        out[(size_t)i + (size_t)nx * j] = v;
    }
    if (a.want_norm) {
        const int nblocks = gridDim.x * gridDim.y;
        const int bl = blockIdx.x + gridDim.x * blockIdx.y;
        const double bsum = block_sum(acc, red);
        double total;
        if (grid_sum_last_block(bsum, a.partials, a.ticket, nblocks, bl, red, &total)) {
            *a.sumsq_out = total;
            if (a.fused_end) cycle_end(const_cast<MGCall *>(a.cp));
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Streaming (y-marching) versions of the two fused level kernels. A block owns a strip of kSW columns (+3 halo columns
// per side, one thread per column) and marches over a chunk of rows; rows of u and rhs arrive through a cp.async ring
// kSD rows ahead of use, every pipeline stage (sweep 1, sweep 2, residual/output) trails the previous one by one row,
// a thread keeps the y neighbours of its own column in registers and reads only the x neighbours from shared memory.
// One __syncthreads per row, no redundant rows except 3 (2) warm-up rows per chunk, ~3x fewer instructions per point
// than the tile kernels; same arithmetic per point, hence bit-identical results.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kSNT = 128;         // threads per block = strip columns including the halo
constexpr int kSW = kSNT - 6;     // output columns per strip (even)
#ifndef B2S_STREAM_D
#define B2S_STREAM_D 6
#define B2S_STREAM_RING 12
#define B2S_STREAM_MINBLOCKS 6
#endif
constexpr int kSD = B2S_STREAM_D;        // prefetch depth in rows
constexpr int kSRing = B2S_STREAM_RING;  // ring rows for the streamed inputs (> kSD + 2, a multiple of 4; = unroll factor)
constexpr int kSCRing = kSRing / 2;  // coarse-row ring of the upward kernel
constexpr int kSP = kSNT + 2;     // ring row pitch: one pad element on each side
constexpr int kSCW = kSW / 2 + 5; // coarse window width of the upward kernel
constexpr int kSCP = kSCW + 1;

constexpr int kSP2 = kSNT + 4;    // pitch of the bulk-copied input rings (parity shift + pads, rows stay 16-byte aligned)

__device__ __forceinline__ void mbar_arrive_plain(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// One elected thread streams row r of two fields (the strip's columns X0-3 .. X0+kSNT-4, clipped to the grid) into ring
// rows with 1-D bulk copies (UBLKCP): no per-thread address arithmetic and no LSU traffic for the loads. Rows of a
// (2^k+1)-wide grid start at alternating 8-byte parities, bulk copies need 16-byte alignment on both sides: the copy
// starts at the even element at or before the first needed one and lands in the ring row shifted by the row parity
// (column x of row r sits at index x-(X0-4)+(r&1)); the consumers know that parity at compile time.
__device__ __forceinline__ void stream_issue_row(double *ringA, double *ringB, const double *gA, const double *gB, int r, int slot,
                                                 uint64_t *bar, int X0, int nx, int ny, int r_last)
{
    if (r < 0 || r >= ny || r > r_last) { mbar_arrive_plain(bar); return; }
    const int x_lo = max(X0 - 3, 0), x_hi = min(X0 + kSNT - 4, nx - 1);
    long long g0 = (long long)x_lo + (long long)nx * r;
    const int shift = (int)(g0 & 1);
    g0 -= shift;
    int n_el = ((x_hi - x_lo + 1) + shift + 1) & ~1;
    const long long total = (long long)nx * ny;
    const int dst = (x_lo - shift) - (X0 - 4) + (r & 1);
    double *dA = ringA + (size_t)slot * kSP2 + dst, *dB = ringB + (size_t)slot * kSP2 + dst;
    if (g0 + n_el > total) {  // the last pair would read one element past the array: fetch the last element by hand
        n_el -= 2;
        dA[n_el] = gA[total - 1];
        dB[n_el] = gB[total - 1];
    }
    if (n_el <= 0) { mbar_arrive_plain(bar); return; }
    const uint32_t bytes = (uint32_t)n_el * 8u;
    mbar_arrive_expect_tx(bar, 2u * bytes);
    bulk_copy_g2s(dA, gA + g0, bytes, bar);
    bulk_copy_g2s(dB, gB + g0, bytes, bar);
}

// Ring slots are relative to the first streamed row of the block, and the row loop is unrolled by the ring size, so
// every shared-memory index below is a compile-time constant and the register queues rotate by renaming.
__global__ void __launch_bounds__(kSNT, B2S_STREAM_MINBLOCKS) mg_down_stream_kernel(const TileArgs a, int ch)
{
    __shared__ __align__(128) double U0[kSRing][kSP2], Fr[kSRing][kSP2];
    __shared__ double S1[4][kSP], S2[4][kSP];
    __shared__ uint64_t bars[kSRing];
    const MGCall *cp = a.cp;
    if (cp->done) return;
    const double *u = a.u_in, *rhs = a.rhs;
    if (a.level == 0) { u = cp->u; rhs = cp->rhs; }
    const int apply_bcs = cp->apply_bcs;
    const Coef k = level_coef(cp, a.level);
    const int nx = a.nx, ny = a.ny, nxc = a.nxc, nyc = a.nyc;
    const int t = threadIdx.x, c = t + 1;
    const int X0 = blockIdx.x * kSW, Y0 = blockIdx.y * ch, Y1 = min(Y0 + ch, ny);
    const int x = X0 - 3 + t;
    const bool dx = x >= 0 && x < nx, ix = x >= 1 && x <= nx - 2;
    const bool outcol = t >= 3 && t <= kSNT - 4 && dx;
    const int s_begin = Y0 - 3, s_end = Y1 + 2;
    if (t == 0) {
        for (int r = 0; r < 4; ++r) { S1[r][0] = 0.0; S1[r][kSP - 1] = 0.0; S2[r][0] = 0.0; S2[r][kSP - 1] = 0.0; }
        for (int r = 0; r < kSRing; ++r) mbar_init(&bars[r], 1);
        fence_mbar_init();
        fence_proxy_async();
#pragma unroll
        for (int j = 0; j < kSD; ++j)  // prologue: rows s_begin .. s_begin+kSD-1 -> slots 0 .. kSD-1
            stream_issue_row(&U0[0][0], &Fr[0][0], u, rhs, s_begin + j, j, &bars[j], X0, nx, ny, s_end);
    }
    __syncthreads();
    // per-column facts of the coarse point (x/2, .) this thread writes on even rows
    const bool xeven = (x & 1) == 0;
    const int Ic = x >> 1;
    const bool cint = Ic >= 1 && Ic <= nxc - 2;
    const bool cmir_lo = apply_bcs && Ic == 1, cmir_hi = apply_bcs && Ic == nxc - 2;
    const bool cedge_bc = apply_bcs && (Ic == 0 || Ic == nxc - 1);
    double u_a = 0.0, u_b = 0.0, u_c = 0.0;  // U0 rows s-2, s-1, s of this column
    double p_a = 0.0, p_b = 0.0;             // S1 rows s-3, s-2
    double q_a = 0.0, q_b = 0.0;             // S2 rows s-4, s-3
    double f_a = 0.0, f_b = 0.0, f_c = 0.0, f_d = 0.0;  // rhs rows s-3 .. s
    uint32_t phase = 0;
    for (int s0 = s_begin; s0 <= s_end; s0 += kSRing, phase ^= 1u) {
#pragma unroll
        for (int j = 0; j < kSRing; ++j) {
            const int s = s0 + j;
            // row s sits in slot j, shifted by its parity: s = Y0 - 3 + (multiple of kSRing) + j with Y0, kSRing even
            constexpr int kDummy = 0; (void)kDummy;
            const int sg = (j + 1) & 1;       // parity shift of row s
            const int sg_a = j & 1;           // ... of row s-1
            mbar_wait(&bars[j], phase);
            __syncthreads();                   // publishes the S1/S2 rows of the previous step; all reads of slot j+kSD done
            if (t == 0)
                stream_issue_row(&U0[0][0], &Fr[0][0], u, rhs, s + kSD, (j + kSD) % kSRing, &bars[(j + kSD) % kSRing], X0, nx, ny,
                                 s_end);
            u_a = u_b; u_b = u_c; u_c = U0[j][c + sg];
            f_a = f_b; f_b = f_c; f_c = f_d; f_d = Fr[j][c + sg];
            // stage A: first sweep at row s-1
            const int ya = s - 1;
            double p_c = u_b;
            if (ix && (unsigned)(ya - 1) < (unsigned)(ny - 2)) {
                const double *row = U0[(j + kSRing - 1) % kSRing] + sg_a;
                const double res = ((row[c + 1] + row[c - 1] + u_c + u_a - k.C * u_b) * k._h2 - f_c);
                p_c = u_b + k.w * res;
            }
            S1[(j + 3) & 3][c] = p_c;
            // stage B: second sweep at row s-2 (x neighbours of S1[s-2] were published one step ago)
            const int yb = s - 2;
            double q_c = p_b;
            if (ix && (unsigned)(yb - 1) < (unsigned)(ny - 2)) {
                const double *row = S1[(j + 2) & 3];
                const double res = ((row[c + 1] + row[c - 1] + p_c + p_a - k.C * p_b) * k._h2 - f_b);
                q_c = p_b + k.w * res;
            }
            S2[(j + 2) & 3][c] = q_c;
            // stage C: output row s-3: smoothed u, injected residual -> coarse rhs, coarse unknown = 0
            const int yc = s - 3;
            if (outcol && (unsigned)(yc - Y0) < (unsigned)(Y1 - Y0)) {
                a.u_out[(size_t)x + (size_t)nx * yc] = q_b;
                // yc = Y0 - 6 + (s0 - s_begin) + j with Y0, kSRing even: the row is even iff j is even (compile time)
                if ((j & 1) == 0 && xeven) {
                    const int J = yc >> 1;
                    const size_t pc = (size_t)Ic + (size_t)nxc * J;
                    a.ec[pc] = 0.0;
                    const bool jint = (unsigned)(J - 1) < (unsigned)(nyc - 2);
                    if (cint && jint) {
                        const double *row = S2[(j + 1) & 3];
                        const double v = ((row[c + 1] + row[c - 1] + q_c + q_a - k.C * q_b) * k._h2 - f_a);
                        a.rc[pc] = v;
                        if (cmir_lo) a.rc[(size_t)0 + (size_t)nxc * J] = v;
                        if (cmir_hi) a.rc[(size_t)(nxc - 1) + (size_t)nxc * J] = v;
                    } else if (!(cedge_bc && jint)) {
                        a.rc[pc] = 0.0;
                    }
                }
            }
            p_a = p_b; p_b = p_c;
            q_a = q_b; q_b = q_c;
        }
    }
}

__global__ void __launch_bounds__(kSNT, B2S_STREAM_MINBLOCKS) mg_up_stream_kernel(const TileArgs a, int ch)
{
    __shared__ __align__(128) double Us[kSRing][kSP2], Fr[kSRing][kSP2];
    __shared__ double C0[4][kSP], T1[4][kSP], Ec[kSCRing][kSCP];
    __shared__ uint64_t bars[kSRing];
    __shared__ double red[32];
    const MGCall *cp = a.cp;
    if (cp->done) return;
    const double *rhs = a.rhs;
    double *out = a.u_out;
    if (a.level == 0) { rhs = cp->rhs; out = cp->u; }
    const int apply_bcs = cp->apply_bcs;
    const Coef k = level_coef(cp, a.level);
    const int nx = a.nx, ny = a.ny, nxc = a.nxc, nyc = a.nyc;
    const int t = threadIdx.x, c = t + 1;
    const int X0 = blockIdx.x * kSW, Y0 = blockIdx.y * ch, Y1 = min(Y0 + ch, ny);  // ch is even
    const int x = X0 - 3 + t;
    const bool dx = x >= 0 && x < nx, ix = x >= 1 && x <= nx - 2;
    const bool outcol = t >= 3 && t <= kSNT - 4 && dx;
    const int cx0 = X0 / 2 - 2;
    if (t == 0) {
        for (int r = 0; r < 4; ++r) { C0[r][0] = 0.0; C0[r][kSP - 1] = 0.0; T1[r][0] = 0.0; T1[r][kSP - 1] = 0.0; }
    }
    // prolongation source column(s) of this thread (Neumann: fine[0,:] = fine[1,:], fine[nx-1,:] = fine[nx-2,:])
    int xs = x;
    if (apply_bcs) {
        if (x == 0) xs = 1;
        else if (x == nx - 1) xs = nx - 2;
    }
    const int Il = (xs >> 1) - cx0;  // local coarse column
    const bool xodd = xs & 1;
    const int s_begin = Y0 - 2, s_end = Y1 + 1;  // s_begin is even: row parity == slot parity
    const int K0 = s_begin >> 1;                  // coarse row held by slot 0 of the coarse ring
    const int Ic = cx0 + t;
    const bool cin = t < kSCW && Ic >= 1 && Ic <= nxc - 2;
    const double *gc = a.ec + (cin ? Ic : 0);
    // coarse row K = K0 + m goes to slot m % kSCRing (the boundary ring of the coarse grid counts as 0)
    auto issue_coarse = [&](int m, int slot) {
        if (t < kSCW) {
            const int K = K0 + m;
            const bool in = cin && K >= 1 && K <= nyc - 2;
            cp_async8(&Ec[slot][t], gc + (in ? (size_t)nxc * K : 0), in);
        }
    };
    if (t == 0) {
        for (int r = 0; r < kSRing; ++r) mbar_init(&bars[r], 1);
        fence_mbar_init();
        fence_proxy_async();
#pragma unroll
        for (int j = 0; j < kSD; ++j)
            stream_issue_row(&Us[0][0], &Fr[0][0], a.u_in, rhs, s_begin + j, j, &bars[j], X0, nx, ny, s_end);
    }
#pragma unroll
    for (int j = 0; j < kSD; ++j) {  // the (quarter-size) coarse rows keep using per-thread cp.async
        if (j == 0) issue_coarse(0, 0);
        if (j & 1) issue_coarse((j + 1) >> 1, ((j + 1) >> 1) % kSCRing);  // coarse row K is first needed by fine row 2K-1
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    __syncthreads();
    double c_a = 0.0, c_b = 0.0;  // corrected u rows s-2, s-1
    double t_a = 0.0, t_b = 0.0;  // first-sweep rows s-3, s-2
    double f_a = 0.0, f_b = 0.0, f_c = 0.0;  // rhs rows s-2 .. s
    double acc = 0.0;
    uint32_t phase = 0;
    for (int s0 = s_begin; s0 <= s_end; s0 += kSRing, phase ^= 1u) {
        const int m0 = (s0 - s_begin) >> 1;
#pragma unroll
        for (int j = 0; j < kSRing; ++j) {
            const int s = s0 + j;
            const int sg = j & 1;  // parity shift of row s (s_begin, kSRing even)
            if ((j + kSD) & 1) issue_coarse(m0 + ((j + kSD + 1) >> 1), ((j + kSD + 1) >> 1) % kSCRing);
            asm volatile("cp.async.commit_group;" ::: "memory");
            asm volatile("cp.async.wait_group %0;" ::"n"(kSD) : "memory");
            mbar_wait(&bars[j], phase);
            __syncthreads();
            if (t == 0)
                stream_issue_row(&Us[0][0], &Fr[0][0], a.u_in, rhs, s + kSD, (j + kSD) % kSRing, &bars[(j + kSD) % kSRing], X0, nx,
                                 ny, s_end);
            f_a = f_b; f_b = f_c; f_c = Fr[j][c + sg];
            // stage A: corrected u at row s:  u_s - P(ec)
            double e;
            {
                const double *r0 = Ec[(j >> 1) % kSCRing] + Il, *r1 = Ec[((j >> 1) + 1) % kSCRing] + Il;
                if (!(j & 1)) e = xodd ? 0.5 * r0[0] + 0.5 * r0[1] : r0[0];
                else e = xodd ? ((0.25 * r0[0] + 0.25 * r0[1]) + 0.25 * r1[0]) + 0.25 * r1[1] : 0.5 * r0[0] + 0.5 * r1[0];
            }
            const double c_c = Us[j][c + sg] - e;
            C0[j & 3][c] = c_c;
            // stage B: first post-sweep at row s-1
            const int yb = s - 1;
            double t_c = c_b;
            if (ix && (unsigned)(yb - 1) < (unsigned)(ny - 2)) {
                const double *row = C0[(j + 3) & 3];
                const double res = ((row[c + 1] + row[c - 1] + c_c + c_a - k.C * c_b) * k._h2 - f_b);
                t_c = c_b + k.w * res;
            }
            T1[(j + 3) & 3][c] = t_c;
            // stage C: second post-sweep at row s-2 -> u, sum of its pre-update res^2
            const int yc = s - 2;
            if (outcol && (unsigned)(yc - Y0) < (unsigned)(Y1 - Y0)) {
                double v = t_b;
                if (ix && (unsigned)(yc - 1) < (unsigned)(ny - 2)) {
                    const double *row = T1[(j + 2) & 3];
                    const double res = ((row[c + 1] + row[c - 1] + t_c + t_a - k.C * t_b) * k._h2 - f_a);
                    acc += res * res;
                    v = t_b + k.w * res;
                }
                out[(size_t)x + (size_t)nx * yc] = v;
            }
            c_a = c_b; c_b = c_c;
            t_a = t_b; t_b = t_c;
        }
    }
    if (a.want_norm) {
        const int nblocks = gridDim.x * gridDim.y;
        const int bl = blockIdx.x + gridDim.x * blockIdx.y;
        const double bsum = block_sum(acc, red);
        double total;
        if (grid_sum_last_block(bsum, a.partials, a.ticket, nblocks, bl, red, &total)) {
            *a.sumsq_out = total;
            if (a.fused_end) cycle_end(const_cast<MGCall *>(a.cp));
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Two-columns-per-thread streaming kernels: same pipeline as above, but every thread owns two adjacent columns, so one
// of the two x neighbours of each point is a register of the same thread, own-column values are read and stage results
// published as 16-byte pairs, the per-row index/predicate overhead is shared by two points, and a thread carries twice
// as many independent FP64 dependency chains (the chains, not issue slots or HBM, bound the one-column version:
// profiles/r01_ncu_source_mg_down_stream2.md). Strip = 256 columns (248 outputs + 4 halo columns per side).
// ---------------------------------------------------------------------------------------------------------------
constexpr int kS2NT = 128;
constexpr int kS2W = 2 * kS2NT - 8;   // output columns per strip
#ifndef B2S_S2_RING
#define B2S_S2_RING 8
#endif
#ifndef B2S_S2_D
#define B2S_S2_D (B2S_S2_RING - 2)
#endif
#ifndef B2S_S2_DC
#define B2S_S2_DC (B2S_S2_RING - 4)
#endif
constexpr int kS2D = B2S_S2_D, kS2Ring = B2S_S2_RING;  // bulk-copy prefetch depth (<= ring - 2) / ring rows (= unroll factor, multiple of 4)
constexpr int kS2DC = B2S_S2_DC;      // look-ahead of the coarse rows (cp.async groups), <= 2 * kS2CRing - 4
static_assert(kS2Ring % 4 == 0 && kS2D <= kS2Ring - 2 && kS2DC <= kS2Ring - 4 && kS2DC >= 1, "stream2 ring parameters");
constexpr int kS2CRing = kS2Ring / 2;
constexpr int kS2P = 2 * kS2NT + 4;   // ring pitch: 2 pad elements left, parity shift, right pad (even)
constexpr int kS2CW = kS2NT + 2;      // coarse window width
constexpr int kS2CP = kS2CW + 2;
constexpr size_t kS2SmemDown = ((size_t)2 * kS2Ring * kS2P + (size_t)2 * 4 * kS2P) * sizeof(double) + kS2Ring * sizeof(uint64_t);
constexpr size_t kS2SmemUp = kS2SmemDown + (size_t)kS2CRing * kS2CP * sizeof(double);

// bulk-copy one row of two fields for a 256-column strip starting at column X0-4 (see stream_issue_row)
__device__ __forceinline__ void stream2_issue_row(double *ringA, double *ringB, const double *gA, const double *gB, int r, int slot,
                                                  uint64_t *bar, int X0, int nx, int ny, int r_last)
{
    if (r < 0 || r >= ny || r > r_last) { mbar_arrive_plain(bar); return; }
    const int x_lo = max(X0 - 4, 0), x_hi = min(X0 + 2 * kS2NT - 5, nx - 1);
    long long g0 = (long long)x_lo + (long long)nx * r;
    const int shift = (int)(g0 & 1);
    g0 -= shift;
    int n_el = ((x_hi - x_lo + 1) + shift + 1) & ~1;
    const long long total = (long long)nx * ny;
    const int dst = (x_lo - shift) - (X0 - 4) + 2 + (r & 1);
    double *dA = ringA + (size_t)slot * kS2P + dst, *dB = ringB + (size_t)slot * kS2P + dst;
    if (g0 + n_el > total) {
        n_el -= 2;
        dA[n_el] = gA[total - 1];
        dB[n_el] = gB[total - 1];
    }
    if (n_el <= 0) { mbar_arrive_plain(bar); return; }
    const uint32_t bytes = (uint32_t)n_el * 8u;
    mbar_arrive_expect_tx(bar, 2u * bytes);
    bulk_copy_g2s(dA, gA + g0, bytes, bar);
    bulk_copy_g2s(dB, gB + g0, bytes, bar);
}

// own pair of a ring row (columns x0, x0+1 at indices ci+SG, ci+SG+1): one 16-byte load when the row's parity shift SG
// keeps it aligned, two 8-byte loads otherwise
template <int SG>
__device__ __forceinline__ double2 ld_pair(const double *row, int ci)
{
    if (SG == 0) return *reinterpret_cast<const double2 *>(row + ci);
    return make_double2(row[ci + 1], row[ci + 2]);
}

// x neighbours of a thread's column pair in the two-column streaming kernels: the left neighbour column is the SECOND
// column of thread t-1, the right one the FIRST column of thread t+1 -- both live in the neighbour lanes' registers, so
// they travel by warp shuffle; only the first / last lane of a warp reads the shared-memory ring row (where the edge
// lanes of the neighbouring warps published their pairs). The stride-2 8-byte ring loads this replaces were 2-way bank
// conflicts (ncu r02: 30-38 % of the streaming kernels' shared-memory wavefronts).
// MEASURED SLOWER and therefore OFF by default (profiles/r02_stream2_shuffle_ab.jsonl: 4097^2 0.372 vs 0.326 ms per V-cycle,
// bit-identical): the kernels are bound by the per-row dependency chain, not by shared-memory bandwidth, and a 64-bit
// shuffle (25 cycles, two SHFL) plus the predicated edge load is a longer link in that chain than the conflicted LDS.64.
#ifndef B2S_S2_SHFL
#define B2S_S2_SHFL 0
#endif
__device__ __forceinline__ void pair_neighbours(const double2 &v, const double *ring_row, int ci, int lane, double &xl, double &xr)
{
#if B2S_S2_SHFL
    xl = __shfl_up_sync(0xffffffffu, v.y, 1);
    xr = __shfl_down_sync(0xffffffffu, v.x, 1);
    if (lane == 0) xl = ring_row[ci - 1];
    if (lane == 31) xr = ring_row[ci + 2];
#else
    xl = ring_row[ci - 1];
    xr = ring_row[ci + 2];
#endif
}
// publish a pair in a ring row for the neighbouring warps' edge lanes (every lane when the shuffle path is disabled)
__device__ __forceinline__ void publish_pair(double *ring_row, int ci, int lane, const double2 &v)
{
#if B2S_S2_SHFL
    if (lane == 0 || lane == 31) *reinterpret_cast<double2 *>(ring_row + ci) = v;
#else
    *reinterpret_cast<double2 *>(ring_row + ci) = v;
#endif
}

__device__ __forceinline__ double jac_res(double xp, double xm, double yp, double ym, double c, double f, const Coef &k)
{
    return ((xp + xm + yp + ym - k.C * c) * k._h2 - f);
}

__global__ void __launch_bounds__(kS2NT + 32) mg_down_stream2_kernel(const TileArgs a, int ch)
{
    extern __shared__ __align__(128) unsigned char s2raw[];
    double *U0 = reinterpret_cast<double *>(s2raw);       // [kS2Ring][kS2P]
    double *Fr = U0 + kS2Ring * kS2P;                     // [kS2Ring][kS2P]
    double *S1 = Fr + kS2Ring * kS2P;                     // [4][kS2P]
    double *S2 = S1 + 4 * kS2P;                           // [4][kS2P]
    uint64_t *bars = reinterpret_cast<uint64_t *>(S2 + 4 * kS2P);
    const MGCall *cp = a.cp;
    if (cp->done) return;
    const double *u = a.u_in, *rhs = a.rhs;
    if (a.level == 0) { u = cp->u; rhs = cp->rhs; }
    const int apply_bcs = cp->apply_bcs;
    const Coef k = level_coef(cp, a.level);
    const int nx = a.nx, ny = a.ny, nxc = a.nxc, nyc = a.nyc;
    const int t = threadIdx.x, ci = 2 * t + 2;  // index of the thread's first column in a ring row (before the shift)
    const int X0 = blockIdx.x * kS2W, Y0 = blockIdx.y * ch, Y1 = min(Y0 + ch, ny);
    const int x0 = X0 - 4 + 2 * t, x1 = x0 + 1;
    const bool d0 = x0 >= 0 && x0 < nx, d1 = x1 >= 0 && x1 < nx;
    const bool i0 = x0 >= 1 && x0 <= nx - 2, i1 = x1 >= 1 && x1 <= nx - 2;
    const bool outt = t >= 2 && t <= kS2NT - 3;
    const bool o0 = outt && d0, o1 = outt && d1;
    const int s_begin = Y0 - 3, s_end = Y1 + 2;
    // Warp kS2NT/32 is the producer: its lane 0 issues the bulk copies of row s + kS2D while the consumer warps work on
    // row s, so the (scalar, ~100-instruction) issue path is off the consumers' critical path. It joins the per-row
    // block barrier, which is what tells it that the ring slot it refills has been read.
    if (threadIdx.x >= kS2NT) {
        if (threadIdx.x == kS2NT) {
            for (int r = 0; r < kS2Ring; ++r) mbar_init(&bars[r], 1);
            fence_mbar_init();
            fence_proxy_async();
#pragma unroll
            for (int j = 0; j < kS2D; ++j) stream2_issue_row(U0, Fr, u, rhs, s_begin + j, j, &bars[j], X0, nx, ny, s_end);
        }
        __syncthreads();
        for (int s0 = s_begin; s0 <= s_end; s0 += kS2Ring) {
#pragma unroll
            for (int j = 0; j < kS2Ring; ++j) {
                __syncthreads();
                if (threadIdx.x == kS2NT)
                    stream2_issue_row(U0, Fr, u, rhs, s0 + j + kS2D, (j + kS2D) % kS2Ring, &bars[(j + kS2D) % kS2Ring], X0, nx, ny,
                                      s_end);
            }
        }
        return;
    }
    if (t == 0) {
        for (int r = 0; r < 4; ++r) {
            S1[r * kS2P + 0] = S1[r * kS2P + 1] = S1[r * kS2P + kS2P - 2] = S1[r * kS2P + kS2P - 1] = 0.0;
            S2[r * kS2P + 0] = S2[r * kS2P + 1] = S2[r * kS2P + kS2P - 2] = S2[r * kS2P + kS2P - 1] = 0.0;
        }
    }
    __syncthreads();
    const int lane = t & 31;
    const int Ic = x0 >> 1;  // x0 is even: the coarse column this thread writes on even rows
    const bool cint = Ic >= 1 && Ic <= nxc - 2;
    const bool cmir_lo = apply_bcs && Ic == 1, cmir_hi = apply_bcs && Ic == nxc - 2;
    const bool cedge_bc = apply_bcs && (Ic == 0 || Ic == nxc - 1);
    double2 u_a = make_double2(0.0, 0.0), u_b = u_a, u_c = u_a;  // U0 rows s-2, s-1, s
    double2 p_a = u_a, p_b = u_a, q_a = u_a, q_b = u_a;          // S1 rows s-3, s-2 ; S2 rows s-4, s-3
    double2 f_a = u_a, f_b = u_a, f_c = u_a, f_d = u_a;          // rhs rows s-3 .. s
    uint32_t phase = 0;
    for (int s0 = s_begin; s0 <= s_end; s0 += kS2Ring, phase ^= 1u) {
#pragma unroll
        for (int j = 0; j < kS2Ring; ++j) {
            const int s = s0 + j;
            constexpr int kOdd = 1;
            const int sg = (j + kOdd) & 1;  // parity shift of row s (= Y0 - 3 + multiple of kS2Ring + j), of row s-1: j & 1
            mbar_wait(&bars[j], phase);
            __syncthreads();
            u_a = u_b; u_b = u_c;
            f_a = f_b; f_b = f_c; f_c = f_d;
            if (sg == 0) { u_c = ld_pair<0>(U0 + j * kS2P, ci); f_d = ld_pair<0>(Fr + j * kS2P, ci); }
            else { u_c = ld_pair<1>(U0 + j * kS2P, ci); f_d = ld_pair<1>(Fr + j * kS2P, ci); }
            // stage A: first sweep at row s-1
            const int ya = s - 1;
            double2 p_c = u_b;
            if ((unsigned)(ya - 1) < (unsigned)(ny - 2)) {
                const double *row = U0 + ((j + kS2Ring - 1) % kS2Ring) * kS2P + (j & 1);
                double xl, xr;
                pair_neighbours(u_b, row, ci, lane, xl, xr);
                if (i0) p_c.x = u_b.x + k.w * jac_res(u_b.y, xl, u_c.x, u_a.x, u_b.x, f_c.x, k);
                if (i1) p_c.y = u_b.y + k.w * jac_res(xr, u_b.x, u_c.y, u_a.y, u_b.y, f_c.y, k);
            }
            publish_pair(S1 + ((j + 3) & 3) * kS2P, ci, lane, p_c);
            // stage B: second sweep at row s-2
            const int yb = s - 2;
            double2 q_c = p_b;
            if ((unsigned)(yb - 1) < (unsigned)(ny - 2)) {
                const double *row = S1 + ((j + 2) & 3) * kS2P;
                double xl, xr;
                pair_neighbours(p_b, row, ci, lane, xl, xr);
                if (i0) q_c.x = p_b.x + k.w * jac_res(p_b.y, xl, p_c.x, p_a.x, p_b.x, f_b.x, k);
                if (i1) q_c.y = p_b.y + k.w * jac_res(xr, p_b.x, p_c.y, p_a.y, p_b.y, f_b.y, k);
            }
            publish_pair(S2 + ((j + 2) & 3) * kS2P, ci, lane, q_c);
            // stage C: output row s-3
            const int yc = s - 3;
#if B2S_S2_SHFL
            double xl_q = 0.0;  // left neighbour of q_b's first column (needed on even rows only: compile-time)
            if ((j & 1) == 0) {
                xl_q = __shfl_up_sync(0xffffffffu, q_b.y, 1);
                if (lane == 0) xl_q = S2[((j + 1) & 3) * kS2P + ci - 1];
            }
#endif
            if ((unsigned)(yc - Y0) < (unsigned)(Y1 - Y0)) {
                const size_t g = (size_t)x0 + (size_t)nx * yc;
                if (o0) a.u_out[g] = q_b.x;
                if (o1) a.u_out[g + 1] = q_b.y;
                if ((j & 1) == 0 && o0) {  // even row (compile time), x0 even: coarse point (Ic, yc/2)
                    const int J = yc >> 1;
                    const size_t pc = (size_t)Ic + (size_t)nxc * J;
                    a.ec[pc] = 0.0;
                    const bool jint = (unsigned)(J - 1) < (unsigned)(nyc - 2);
                    if (cint && jint) {
#if B2S_S2_SHFL
                        const double xl = xl_q;
#else
                        const double xl = S2[((j + 1) & 3) * kS2P + ci - 1];
#endif
                        const double v = jac_res(q_b.y, xl, q_c.x, q_a.x, q_b.x, f_a.x, k);
                        a.rc[pc] = v;
                        if (cmir_lo) a.rc[(size_t)0 + (size_t)nxc * J] = v;
                        if (cmir_hi) a.rc[(size_t)(nxc - 1) + (size_t)nxc * J] = v;
                    } else if (!(cedge_bc && jint)) {
                        a.rc[pc] = 0.0;
                    }
                }
            }
            p_a = p_b; p_b = p_c;
            q_a = q_b; q_b = q_c;
        }
    }
}

__global__ void __launch_bounds__(kS2NT + 32) mg_up_stream2_kernel(const TileArgs a, int ch)
{
    extern __shared__ __align__(128) unsigned char s2raw[];
    __shared__ double red[32];
    double *Us = reinterpret_cast<double *>(s2raw);
    double *Fr = Us + kS2Ring * kS2P;
    double *C0 = Fr + kS2Ring * kS2P;
    double *T1 = C0 + 4 * kS2P;
    uint64_t *bars = reinterpret_cast<uint64_t *>(T1 + 4 * kS2P);
    double *Ec = reinterpret_cast<double *>(bars + kS2Ring);  // [kS2CRing][kS2CP]
    const MGCall *cp = a.cp;
    if (cp->done) return;
    const double *rhs = a.rhs;
    double *out = a.u_out;
    if (a.level == 0) { rhs = cp->rhs; out = cp->u; }
    const int apply_bcs = cp->apply_bcs;
    const Coef k = level_coef(cp, a.level);
    const int nx = a.nx, ny = a.ny, nxc = a.nxc, nyc = a.nyc;
    const int t = threadIdx.x, ci = 2 * t + 2;
    const int X0 = blockIdx.x * kS2W, Y0 = blockIdx.y * ch, Y1 = min(Y0 + ch, ny);  // ch even
    const int x0 = X0 - 4 + 2 * t, x1 = x0 + 1;
    const bool d0 = x0 >= 0 && x0 < nx, d1 = x1 >= 0 && x1 < nx;
    const bool i0 = x0 >= 1 && x0 <= nx - 2, i1 = x1 >= 1 && x1 <= nx - 2;
    const bool outt = t >= 2 && t <= kS2NT - 3;
    const bool o0 = outt && d0, o1 = outt && d1;
    const int s_begin = Y0 - 2, s_end = Y1 + 1;  // even
    const int K0 = s_begin >> 1;
    const int cx0 = X0 / 2 - 3;                  // coarse window origin; local coarse column of x0: t + 1
    const int Il = t + 1;
    const bool bc_first = apply_bcs && x0 == 0;       // fine[0,:] = fine[1,:]
    const bool bc_last = apply_bcs && x0 == nx - 1;   // fine[nx-1,:] = fine[nx-2,:]
    const bool producer = threadIdx.x >= kS2NT;  // see mg_down_stream2_kernel
    const int lane = t & 31;
    if (threadIdx.x == kS2NT) {
        for (int r = 0; r < kS2Ring; ++r) mbar_init(&bars[r], 1);
        fence_mbar_init();
        fence_proxy_async();
#pragma unroll
        for (int j = 0; j < kS2D; ++j) stream2_issue_row(Us, Fr, a.u_in, rhs, s_begin + j, j, &bars[j], X0, nx, ny, s_end);
    }
    if (t == 0) {
        for (int r = 0; r < 4; ++r) {
            C0[r * kS2P + 0] = C0[r * kS2P + 1] = C0[r * kS2P + kS2P - 2] = C0[r * kS2P + kS2P - 1] = 0.0;
            T1[r * kS2P + 0] = T1[r * kS2P + 1] = T1[r * kS2P + kS2P - 2] = T1[r * kS2P + kS2P - 1] = 0.0;
        }
    }
    // coarse rows: coarse row K0 + m -> slot m % kS2CRing, window columns cx0 .. cx0 + kS2CW - 1 (boundary ring = 0)
    // (loaded by the 32 lanes of the producer warp with cp.async; its wait_group + the per-row barrier publish them)
    auto issue_coarse = [&](int m, int slot) {
        const int K = K0 + m;
        const bool krow = K >= 1 && K <= nyc - 2;
        for (int e = (int)threadIdx.x - kS2NT; e < kS2CW; e += 32) {
            const int I = cx0 + e;
            const bool in = krow && I >= 1 && I <= nxc - 2;
            cp_async8(Ec + slot * kS2CP + e, a.ec + (in ? (size_t)I + (size_t)nxc * K : 0), in);
        }
    };
    if (producer) {
#pragma unroll
        for (int j = 0; j < kS2DC; ++j) {
            if (j == 0) issue_coarse(0, 0);
            if (j & 1) issue_coarse((j + 1) >> 1, ((j + 1) >> 1) % kS2CRing);
            asm volatile("cp.async.commit_group;" ::: "memory");
        }
    }
    __syncthreads();
    double2 z2 = make_double2(0.0, 0.0);
    double2 c_a = z2, c_b = z2, t_a = z2, t_b = z2, f_a = z2, f_b = z2, f_c = z2;
    double acc = 0.0;
    uint32_t phase = 0;
    for (int s0 = s_begin; s0 <= s_end; s0 += kS2Ring, phase ^= 1u) {
        const int m0 = (s0 - s_begin) >> 1;
#pragma unroll
        for (int j = 0; j < kS2Ring; ++j) {
            const int s = s0 + j;
            if (producer) {
                if ((j + kS2DC) & 1) issue_coarse(m0 + ((j + kS2DC + 1) >> 1), ((j + kS2DC + 1) >> 1) % kS2CRing);
                asm volatile("cp.async.commit_group;" ::: "memory");
                asm volatile("cp.async.wait_group %0;" ::"n"(kS2DC) : "memory");  // coarse rows of fine row s have landed
                __syncthreads();
                if (threadIdx.x == kS2NT)
                    stream2_issue_row(Us, Fr, a.u_in, rhs, s + kS2D, (j + kS2D) % kS2Ring, &bars[(j + kS2D) % kS2Ring], X0, nx, ny,
                                      s_end);
                continue;
            }
            mbar_wait(&bars[j], phase);
            __syncthreads();
            f_a = f_b; f_b = f_c;
            double2 us;
            if ((j & 1) == 0) { us = ld_pair<0>(Us + j * kS2P, ci); f_c = ld_pair<0>(Fr + j * kS2P, ci); }
            else { us = ld_pair<1>(Us + j * kS2P, ci); f_c = ld_pair<1>(Fr + j * kS2P, ci); }
            // stage A: corrected u at row s = u_s - P(ec); x0 is even (coarse column Il), x1 odd (between Il and Il+1)
            double e0, e1;
            {
                const double *r0 = Ec + ((j >> 1) % kS2CRing) * kS2CP + Il, *r1 = Ec + (((j >> 1) + 1) % kS2CRing) * kS2CP + Il;
                if ((j & 1) == 0) {
                    e0 = r0[0];
                    e1 = 0.5 * r0[0] + 0.5 * r0[1];
                    if (bc_last) e0 = 0.5 * r0[-1] + 0.5 * r0[0];
                } else {
                    e0 = 0.5 * r0[0] + 0.5 * r1[0];
                    e1 = ((0.25 * r0[0] + 0.25 * r0[1]) + 0.25 * r1[0]) + 0.25 * r1[1];
                    if (bc_last) e0 = ((0.25 * r0[-1] + 0.25 * r0[0]) + 0.25 * r1[-1]) + 0.25 * r1[0];
                }
                if (bc_first) e0 = e1;
            }
            const double2 c_c = make_double2(us.x - e0, us.y - e1);
            publish_pair(C0 + (j & 3) * kS2P, ci, lane, c_c);
            // stage B: first post-sweep at row s-1
            const int yb = s - 1;
            double2 t_c = c_b;
            if ((unsigned)(yb - 1) < (unsigned)(ny - 2)) {
                const double *row = C0 + ((j + 3) & 3) * kS2P;
                double xl, xr;
                pair_neighbours(c_b, row, ci, lane, xl, xr);
                if (i0) t_c.x = c_b.x + k.w * jac_res(c_b.y, xl, c_c.x, c_a.x, c_b.x, f_b.x, k);
                if (i1) t_c.y = c_b.y + k.w * jac_res(xr, c_b.x, c_c.y, c_a.y, c_b.y, f_b.y, k);
            }
            publish_pair(T1 + ((j + 3) & 3) * kS2P, ci, lane, t_c);
            // stage C: second post-sweep at row s-2 -> u, sum of its pre-update res^2
            const int yc = s - 2;
            if ((unsigned)(yc - Y0) < (unsigned)(Y1 - Y0)) {
                double2 v = t_b;
                if ((unsigned)(yc - 1) < (unsigned)(ny - 2)) {
                    const double *row = T1 + ((j + 2) & 3) * kS2P;
                    double xl, xr;
                    pair_neighbours(t_b, row, ci, lane, xl, xr);
                    if (i0 && o0) {
                        const double res = jac_res(t_b.y, xl, t_c.x, t_a.x, t_b.x, f_a.x, k);
                        acc += res * res;
                        v.x = t_b.x + k.w * res;
                    }
                    if (i1 && o1) {
                        const double res = jac_res(xr, t_b.x, t_c.y, t_a.y, t_b.y, f_a.y, k);
                        acc += res * res;
                        v.y = t_b.y + k.w * res;
                    }
                }
                const size_t g = (size_t)x0 + (size_t)nx * yc;
                if (o0) out[g] = v.x;
                if (o1) out[g + 1] = v.y;
            }
            c_a = c_b; c_b = c_c;
            t_a = t_b; t_b = t_c;
        }
    }
    if (a.want_norm) {
        const int nblocks = gridDim.x * gridDim.y;
        const int bl = blockIdx.x + gridDim.x * blockIdx.y;
        const double bsum = block_sum(acc, red);
        double total;
        if (grid_sum_last_block(bsum, a.partials, a.ticket, nblocks, bl, red, &total)) {
            *a.sumsq_out = total;
            if (a.fused_end) cycle_end(const_cast<MGCall *>(a.cp));
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// One-warp-per-strip streaming kernels: the same two fused level passes with NO block barrier and NO shared-memory
// exchange between the pipeline stages. A block is a single warp that owns 64 columns (56 outputs + 4 halo columns
// per side, two columns per lane) and a chunk of rows. Input rows arrive through an 8-row bulk-copy ring (lane 0 refills
// the slot it has just consumed; completion on one mbarrier per slot); every stage keeps the y neighbours of its two
// columns in registers and gets the x neighbours of the previous stage from the adjacent lanes by warp shuffles. With
// ~9 KB of shared memory and one warp per block, 21-25 independent warps are resident per SM (the block-wide variants: 16
// consumer warps coupled four at a time by a barrier per row). Same arithmetic per point -> bit-identical results.
// Measured on B200 it is SLOWER than the block-wide two-column kernels (4097^2: 0.427 vs 0.358 ms per V-cycle): the
// 512-byte bulk copies and the 14 % halo overhead cost more than the barriers; kept as fuse_sweeps = 5 for comparison.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kWW = 56;      // output columns per strip
constexpr int kWRing = 8;    // ring rows = prefetch depth = unroll factor
constexpr int kWP = 68;      // ring pitch: 2 pad + 64 columns + parity shift + pad (even)

// lane 0: bulk-copy row r of two fields for the 64-column strip starting at column X0-4 (see stream_issue_row)
__device__ __forceinline__ void warp_issue_row(double *dstA, double *dstB, const double *gA, const double *gB, int r, uint64_t *bar,
                                               int X0, int nx, int ny, int r_last)
{
    if (r < 0 || r >= ny || r > r_last) { mbar_arrive_plain(bar); return; }
    const int x_lo = max(X0 - 4, 0), x_hi = min(X0 + 59, nx - 1);
    long long g0 = (long long)x_lo + (long long)nx * r;
    const int shift = (int)(g0 & 1);
    g0 -= shift;
    int n_el = ((x_hi - x_lo + 1) + shift + 1) & ~1;
    const long long total = (long long)nx * ny;
    const int dst = (x_lo - shift) - (X0 - 4) + 2 + (r & 1);
    if (g0 + n_el > total) {  // last element of the array: fetched by hand (the pair would overrun)
        n_el -= 2;
        dstA[dst + n_el] = gA[total - 1];
        dstB[dst + n_el] = gB[total - 1];
    }
    if (n_el <= 0) { mbar_arrive_plain(bar); return; }
    const uint32_t bytes = (uint32_t)n_el * 8u;
    mbar_arrive_expect_tx(bar, 2u * bytes);
    bulk_copy_g2s(dstA + dst, gA + g0, bytes, bar);
    bulk_copy_g2s(dstB + dst, gB + g0, bytes, bar);
}

__device__ __forceinline__ double shfl_up1(double v) { return __shfl_up_sync(0xffffffffu, v, 1); }
__device__ __forceinline__ double shfl_dn1(double v) { return __shfl_down_sync(0xffffffffu, v, 1); }

__global__ void __launch_bounds__(32) mg_down_warp_kernel(const TileArgs a, int ch)
{
    __shared__ __align__(128) double U0[kWRing][kWP], Fr[kWRing][kWP];
    __shared__ uint64_t bars[kWRing];
    const MGCall *cp = a.cp;
    if (cp->done) return;
    const double *u = a.u_in, *rhs = a.rhs;
    if (a.level == 0) { u = cp->u; rhs = cp->rhs; }
    const int apply_bcs = cp->apply_bcs;
    const Coef k = level_coef(cp, a.level);
    const int nx = a.nx, ny = a.ny, nxc = a.nxc, nyc = a.nyc;
    const int lane = threadIdx.x, ci = 2 * lane + 2;
    const int X0 = blockIdx.x * kWW, Y0 = blockIdx.y * ch, Y1 = min(Y0 + ch, ny);
    const int x0 = X0 - 4 + 2 * lane, x1 = x0 + 1;
    const bool d0 = x0 >= 0 && x0 < nx, d1 = x1 >= 0 && x1 < nx;
    const bool i0 = x0 >= 1 && x0 <= nx - 2, i1 = x1 >= 1 && x1 <= nx - 2;
    const bool outl = lane >= 2 && lane <= 29;
    const bool o0 = outl && d0, o1 = outl && d1;
    const int s_begin = Y0 - 3, s_end = Y1 + 2;
    if (lane == 0) {
        for (int r = 0; r < kWRing; ++r) mbar_init(&bars[r], 1);
        fence_mbar_init();
        fence_proxy_async();
#pragma unroll
        for (int j = 0; j < kWRing; ++j) warp_issue_row(U0[j], Fr[j], u, rhs, s_begin + j, &bars[j], X0, nx, ny, s_end);
    }
    __syncwarp();
    const int Ic = x0 >> 1;
    const bool cint = Ic >= 1 && Ic <= nxc - 2;
    const bool cmir_lo = apply_bcs && Ic == 1, cmir_hi = apply_bcs && Ic == nxc - 2;
    const bool cedge_bc = apply_bcs && (Ic == 0 || Ic == nxc - 1);
    double2 z2 = make_double2(0.0, 0.0);
    double2 u_a = z2, u_b = z2, u_c = z2, p_a = z2, p_b = z2, q_a = z2, q_b = z2, f_a = z2, f_b = z2, f_c = z2, f_d = z2;
    uint32_t phase = 0;
    for (int s0 = s_begin; s0 <= s_end; s0 += kWRing, phase ^= 1u) {
#pragma unroll
        for (int j = 0; j < kWRing; ++j) {
            const int s = s0 + j;
            mbar_wait(&bars[j], phase);
            u_a = u_b; u_b = u_c;
            f_a = f_b; f_b = f_c; f_c = f_d;
            if (((j + 1) & 1) == 0) { u_c = ld_pair<0>(U0[j], ci); f_d = ld_pair<0>(Fr[j], ci); }  // parity of row s: (j+1)&1
            else { u_c = ld_pair<1>(U0[j], ci); f_d = ld_pair<1>(Fr[j], ci); }
            // stage A: first sweep at row s-1 (x neighbours of the row live in the adjacent lanes)
            double xl = shfl_up1(u_b.y), xr = shfl_dn1(u_b.x);
            double2 p_c = u_b;
            if ((unsigned)(s - 2) < (unsigned)(ny - 2)) {
                if (i0) p_c.x = u_b.x + k.w * jac_res(u_b.y, xl, u_c.x, u_a.x, u_b.x, f_c.x, k);
                if (i1) p_c.y = u_b.y + k.w * jac_res(xr, u_b.x, u_c.y, u_a.y, u_b.y, f_c.y, k);
            }
            // the loaded values have been consumed: refill the slot with row s + kWRing
            __syncwarp();
            if (lane == 0) warp_issue_row(U0[j], Fr[j], u, rhs, s + kWRing, &bars[j], X0, nx, ny, s_end);
            // stage B: second sweep at row s-2
            xl = shfl_up1(p_b.y); xr = shfl_dn1(p_b.x);
            double2 q_c = p_b;
            if ((unsigned)(s - 3) < (unsigned)(ny - 2)) {
                if (i0) q_c.x = p_b.x + k.w * jac_res(p_b.y, xl, p_c.x, p_a.x, p_b.x, f_b.x, k);
                if (i1) q_c.y = p_b.y + k.w * jac_res(xr, p_b.x, p_c.y, p_a.y, p_b.y, f_b.y, k);
            }
            // stage C: output row s-3
            const int yc = s - 3;
            xl = shfl_up1(q_b.y);
            if ((unsigned)(yc - Y0) < (unsigned)(Y1 - Y0)) {
                const size_t g = (size_t)x0 + (size_t)nx * yc;
                if (o0) a.u_out[g] = q_b.x;
                if (o1) a.u_out[g + 1] = q_b.y;
                if ((j & 1) == 0 && o0) {  // even row (compile time), x0 even: coarse point (Ic, yc/2)
                    const int J = yc >> 1;
                    const size_t pc = (size_t)Ic + (size_t)nxc * J;
                    a.ec[pc] = 0.0;
                    const bool jint = (unsigned)(J - 1) < (unsigned)(nyc - 2);
                    if (cint && jint) {
                        const double v = jac_res(q_b.y, xl, q_c.x, q_a.x, q_b.x, f_a.x, k);
                        a.rc[pc] = v;
                        if (cmir_lo) a.rc[(size_t)0 + (size_t)nxc * J] = v;
                        if (cmir_hi) a.rc[(size_t)(nxc - 1) + (size_t)nxc * J] = v;
                    } else if (!(cedge_bc && jint)) {
                        a.rc[pc] = 0.0;
                    }
                }
            }
            p_a = p_b; p_b = p_c;
            q_a = q_b; q_b = q_c;
        }
    }
}

__global__ void __launch_bounds__(32) mg_up_warp_kernel(const TileArgs a, int ch)
{
    __shared__ __align__(128) double Us[kWRing][kWP], Fr[kWRing][kWP];
    __shared__ uint64_t bars[kWRing];
    __shared__ double red[32];
    const MGCall *cp = a.cp;
    if (cp->done) return;
    const double *rhs = a.rhs;
    double *out = a.u_out;
    if (a.level == 0) { rhs = cp->rhs; out = cp->u; }
    const int apply_bcs = cp->apply_bcs;
    const Coef k = level_coef(cp, a.level);
    const int nx = a.nx, ny = a.ny, nxc = a.nxc, nyc = a.nyc;
    const int lane = threadIdx.x, ci = 2 * lane + 2;
    const int X0 = blockIdx.x * kWW, Y0 = blockIdx.y * ch, Y1 = min(Y0 + ch, ny);  // ch even
    const int x0 = X0 - 4 + 2 * lane, x1 = x0 + 1;
    const bool d0 = x0 >= 0 && x0 < nx, d1 = x1 >= 0 && x1 < nx;
    const bool i0 = x0 >= 1 && x0 <= nx - 2, i1 = x1 >= 1 && x1 <= nx - 2;
    const bool outl = lane >= 2 && lane <= 29;
    const bool o0 = outl && d0, o1 = outl && d1;
    const int s_begin = Y0 - 2, s_end = Y1 + 1;  // even
    const bool bc_first = apply_bcs && x0 == 0, bc_last = apply_bcs && x0 == nx - 1;
    if (lane == 0) {
        for (int r = 0; r < kWRing; ++r) mbar_init(&bars[r], 1);
        fence_mbar_init();
        fence_proxy_async();
#pragma unroll
        for (int j = 0; j < kWRing; ++j) warp_issue_row(Us[j], Fr[j], a.u_in, rhs, s_begin + j, &bars[j], X0, nx, ny, s_end);
    }
    __syncwarp();
    // coarse correction: this lane's coarse column I = x0/2 (x0 is even), rows K; the boundary ring counts as 0. Each lane
    // loads its own column straight into registers, three coarse rows ahead; column I+1 / I-1 come from the adjacent lanes.
    const int Ic = x0 >> 1;
    const bool cin = Ic >= 1 && Ic <= nxc - 2;
    const double *gc = a.ec + (cin ? Ic : 0);
    auto load_coarse = [&](int K) -> double {
        return (cin && K >= 1 && K <= nyc - 2) ? gc[(size_t)nxc * K] : 0.0;
    };
    const int K0 = s_begin >> 1;
    double cA = load_coarse(K0), cB = load_coarse(K0 + 1), cC = load_coarse(K0 + 2), cD = load_coarse(K0 + 3);
    double2 z2 = make_double2(0.0, 0.0);
    double2 c_a = z2, c_b = z2, t_a = z2, t_b = z2, f_a = z2, f_b = z2, f_c = z2;
    double acc = 0.0;
    uint32_t phase = 0;
    for (int s0 = s_begin; s0 <= s_end; s0 += kWRing, phase ^= 1u) {
        const int Kb = (s0 >> 1);  // coarse row of fine row s0 (s0 even)
#pragma unroll
        for (int j = 0; j < kWRing; ++j) {
            const int s = s0 + j;
            mbar_wait(&bars[j], phase);
            f_a = f_b; f_b = f_c;
            double2 us;
            if ((j & 1) == 0) { us = ld_pair<0>(Us[j], ci); f_c = ld_pair<0>(Fr[j], ci); }
            else { us = ld_pair<1>(Us[j], ci); f_c = ld_pair<1>(Fr[j], ci); }
            // stage A: corrected u at row s. cA = coarse row s>>1, cB = the next one (this lane's column)
            const double a1 = shfl_dn1(cA), b1 = shfl_dn1(cB), am = shfl_up1(cA), bm = shfl_up1(cB);
            double e0, e1;
            if ((j & 1) == 0) {
                e0 = cA;
                e1 = 0.5 * cA + 0.5 * a1;
                if (bc_last) e0 = 0.5 * am + 0.5 * cA;
            } else {
                e0 = 0.5 * cA + 0.5 * cB;
                e1 = ((0.25 * cA + 0.25 * a1) + 0.25 * cB) + 0.25 * b1;
                if (bc_last) e0 = ((0.25 * am + 0.25 * cA) + 0.25 * bm) + 0.25 * cB;
            }
            if (bc_first) e0 = e1;
            const double2 c_c = make_double2(us.x - e0, us.y - e1);
            if (j & 1) {  // the next fine row starts the next coarse row
                cA = cB; cB = cC; cC = cD;
                cD = load_coarse(Kb + (j >> 1) + 4);
            }
            __syncwarp();
            if (lane == 0) warp_issue_row(Us[j], Fr[j], a.u_in, rhs, s + kWRing, &bars[j], X0, nx, ny, s_end);
            // stage B: first post-sweep at row s-1
            double xl = shfl_up1(c_b.y), xr = shfl_dn1(c_b.x);
            double2 t_c = c_b;
            if ((unsigned)(s - 2) < (unsigned)(ny - 2)) {
                if (i0) t_c.x = c_b.x + k.w * jac_res(c_b.y, xl, c_c.x, c_a.x, c_b.x, f_b.x, k);
                if (i1) t_c.y = c_b.y + k.w * jac_res(xr, c_b.x, c_c.y, c_a.y, c_b.y, f_b.y, k);
            }
            // stage C: second post-sweep at row s-2 -> u, sum of its pre-update res^2
            const int yc = s - 2;
            xl = shfl_up1(t_b.y); xr = shfl_dn1(t_b.x);
            if ((unsigned)(yc - Y0) < (unsigned)(Y1 - Y0)) {
                double2 v = t_b;
                if ((unsigned)(yc - 1) < (unsigned)(ny - 2)) {
                    if (i0 && o0) {
                        const double res = jac_res(t_b.y, xl, t_c.x, t_a.x, t_b.x, f_a.x, k);
                        acc += res * res;
                        v.x = t_b.x + k.w * res;
                    }
                    if (i1 && o1) {
                        const double res = jac_res(xr, t_b.x, t_c.y, t_a.y, t_b.y, f_a.y, k);
                        acc += res * res;
                        v.y = t_b.y + k.w * res;
                    }
                }
                const size_t g = (size_t)x0 + (size_t)nx * yc;
                if (o0) out[g] = v.x;
                if (o1) out[g + 1] = v.y;
            }
            c_a = c_b; c_b = c_c;
            t_a = t_b; t_b = t_c;
        }
    }
    if (a.want_norm) {
        const int nblocks = gridDim.x * gridDim.y;
        const int bl = blockIdx.x + gridDim.x * blockIdx.y;
        const double bsum = block_sum(acc, red);
        double total;
        if (grid_sum_last_block(bsum, a.partials, a.ticket, nblocks, bl, red, &total)) {
            *a.sumsq_out = total;
            if (a.fused_end) cycle_end(const_cast<MGCall *>(a.cp));
        }
    }
}

// apply_boundary_conditions!(T): Dirichlet T[:,0]=1, T[:,ny-1]=0, then Neumann T[0,:]=T[1,:], T[nx-1,:]=T[nx-2,:]
// (part2_utils.jl:21-39). kind: 0 both, 1 Dirichlet only, 2 Neumann only.
__global__ void mg_bc_kernel(const MGCall *cp, double *T, int nx, int ny, int kind)
{
    if (cp != nullptr) {
        if (cp->done || !cp->bc_before) return;
        T = cp->u;
    }
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const bool dir = kind == 0 || kind == 1, neu = kind == 0 || kind == 2;
    if (t < nx) {  // y edges; x-corners are finalised by the Neumann part below
        const int i = t;
        const bool corner = (i == 0 || i == nx - 1);
        if (dir && !(neu && corner)) {
            T[(size_t)i] = 1.0;
            T[(size_t)i + (size_t)nx * (ny - 1)] = 0.0;
        }
    } else if (t < nx + ny) {
        const int j = t - nx;
        if (neu) {
            double lo, hi;
            if (dir && j == 0) { lo = 1.0; hi = 1.0; }
            else if (dir && j == ny - 1) { lo = 0.0; hi = 0.0; }
            else { lo = T[(size_t)1 + (size_t)nx * j]; hi = T[(size_t)(nx - 2) + (size_t)nx * j]; }
            T[(size_t)nx * j] = lo;
            T[(size_t)(nx - 1) + (size_t)nx * j] = hi;
        }
    }
}

// (lap - c) T with separate hx, hy, true divisions (krylov.jl:7-13). Interior only.
__global__ void __launch_bounds__(kMGBX) mg_matvec_kernel(const double *__restrict__ T, double hx, double hy, double c,
                                                          double *__restrict__ out, int nx, int ny, int rows)
{
    const int i = blockIdx.x * kMGBX + threadIdx.x;
    if (i < 1 || i > nx - 2) return;
    const int j0 = max(1, blockIdx.y * rows), j1 = min(blockIdx.y * rows + rows, ny - 1);
    for (int j = j0; j < j1; ++j) {
        const size_t p = (size_t)i + (size_t)nx * j;
        out[p] = ((T[p + 1] - 2 * T[p] + T[p - 1]) / (hx * hx) + (T[p + nx] - 2 * T[p] + T[p - nx]) / (hy * hy)) - c * T[p];
    }
}

// r = f - q on the interior, 0 on the frame (initial residual of the MG-preconditioned CG)
__global__ void __launch_bounds__(kMGBX) mg_pcg_residual_kernel(const double *__restrict__ f, const double *__restrict__ q,
                                                                double *__restrict__ r, int nx, int ny, int rows)
{
    const int i = blockIdx.x * kMGBX + threadIdx.x;
    if (i >= nx) return;
    const int j0 = blockIdx.y * rows, j1 = min(j0 + rows, ny);
    for (int j = j0; j < j1; ++j) {
        const size_t p = (size_t)i + (size_t)nx * j;
        r[p] = (i >= 1 && i <= nx - 2 && j >= 1 && j <= ny - 2) ? f[p] - q[p] : 0.0;
    }
}

// Deterministic reductions / vector updates of cg! and of the residual checks.
// op 0: sum x*y ; op 1: sum x*x.  Fixed grid (kReduceBlocks) -> fixed summation order.
constexpr int kReduceBlocks = 296, kReduceThreads = 256;
__global__ void __launch_bounds__(kReduceThreads) mg_reduce_kernel(const double *__restrict__ x, const double *__restrict__ y,
                                                                  size_t n, double *partials, unsigned int *ticket,
                                                                  double *out, const MGCall *cp, int use_cp_rhs)
{
    __shared__ double red[32];
    if (use_cp_rhs) { x = cp->rhs; y = cp->rhs; }
    double acc = 0.0;
    for (size_t p = (size_t)blockIdx.x * kReduceThreads + threadIdx.x; p < n; p += (size_t)kReduceBlocks * kReduceThreads)
        acc += x[p] * y[p];
    const double bsum = block_sum(acc, red);
    double total;
    if (grid_sum_last_block(bsum, partials, ticket, gridDim.x, blockIdx.x, red, &total)) *out = total;
}
// op 0: y += alpha*x ; op 1: y = x + beta*y ; op 2: y = y - x ; op 3: y[0] = y[2] + y[3] (red + black partial sums)
__global__ void mg_axpy_kernel(double s, const double *__restrict__ x, double *__restrict__ y, size_t n, int op)
{
    if (op == 3) {
        if (blockIdx.x == 0 && threadIdx.x == 0) y[0] = y[2] + y[3];
        return;
    }
    for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (size_t)gridDim.x * blockDim.x) {
        if (op == 0) y[p] += s * x[p];
        else if (op == 1) y[p] = x[p] + s * y[p];
        else y[p] = y[p] - x[p];
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Collapsed coarse hierarchy: every level that fits into shared memory is processed by ONE thread block -- the
// remaining restrict / smooth / coarsest solve / prolongate chain of the V-cycle without a single kernel launch or
// host synchronisation in between. The data-dependent coarsest solve (<= 20*cs Jacobi sweeps with the reference's
// exit test, or CG) runs inside one warp when the grid is tiny, so its ~50-100 dependent sweeps cost __syncwarp()s.
// ---------------------------------------------------------------------------------------------------------------
struct CoarseArgs {
    const MGCall *cp;
    int level0;               // global level index of the first shared-memory level
    int nlev;                 // number of shared-memory levels (>= 1); the last one is the coarsest
    int nx[kMaxLevels], ny[kMaxLevels];
    const double *rhs_in;     // global rhs of level0 (nullptr: cp->rhs, i.e. level0 == 0)
    double *u_io;             // global unknown of level0 (nullptr: cp->u)
    int u_is_input;           // 1: start from the values in u_io (top-level coarsest solve); 0: start from zero
    int coarse_solve_size, coarse_solver, smoother, restriction;
    double *sumsq_out;        // nullable: sum res^2 of the last sweep when level0 == 0 has no finer level
    long long *prof;          // nullable: clock64() stamps of the phases (B2S_MG_PROF=1), [0] = count
};

__device__ __forceinline__ void coarse_stamp(const CoarseArgs &a)
{
    if (a.prof != nullptr && threadIdx.x == 0) {
        const long long n = a.prof[0];
        if (n < 60) { a.prof[1 + n] = clock64(); a.prof[0] = n + 1; }
    }
}

struct BlockGroup {
    double *red;
    __device__ __forceinline__ int rank() const { return threadIdx.x; }
    __device__ __forceinline__ int size() const { return blockDim.x; }
    __device__ __forceinline__ void sync() const { __syncthreads(); }
    __device__ __forceinline__ double sum(double v) const
    {
        double r = block_sum(v, red);
        __shared__ double bc;
        if (threadIdx.x == 0) bc = r;
        __syncthreads();
        r = bc;
        __syncthreads();
        return r;
    }
};
struct WarpGroup {
    __device__ __forceinline__ int rank() const { return threadIdx.x & 31; }
    __device__ __forceinline__ int size() const { return 32; }
    __device__ __forceinline__ void sync() const { __syncwarp(); }
    __device__ __forceinline__ double sum(double v) const { return warp_sum(v); }
};

// (i, j) of a strided flat index, advanced without a division per point
struct Idx2 {
    int p, i, j, di, dj, nx, stride;
    __device__ __forceinline__ Idx2(int rank, int stride_, int nx_) : p(rank), nx(nx_), stride(stride_)
    {
        j = rank / nx_; i = rank - j * nx_;
        dj = stride_ / nx_; di = stride_ - dj * nx_;
    }
    __device__ __forceinline__ void next()
    {
        p += stride; i += di; j += dj;
        if (i >= nx) { i -= nx; ++j; }
    }
};

template <class G>
__device__ __forceinline__ void sm_fill(const G &g, double *a, int n, double v)
{
    for (int p = g.rank(); p < n; p += g.size()) a[p] = v;
}

// out = u + w*res(u) (frame copied); returns sum res^2 if want_norm (valid on every thread of the group)
template <class G>
__device__ __forceinline__ double sm_jacobi(const G &g, const double *u, const double *rhs, double *out, int nx, int ny,
                                            const Coef &k, bool want_norm)
{
    double acc = 0.0;
    const int n = nx * ny;
    for (Idx2 q(g.rank(), g.size(), nx); q.p < n; q.next()) {
        const int p = q.p, i = q.i, j = q.j;
        if (i >= 1 && i <= nx - 2 && j >= 1 && j <= ny - 2) {
            const double r = ((u[p + 1] + u[p - 1] + u[p + nx] + u[p - nx] - k.C * u[p]) * k._h2 - rhs[p]);
            acc += r * r;
            out[p] = u[p] + k.w * r;
        } else {
            out[p] = u[p];
        }
    }
    double tot = 0.0;
    if (want_norm) tot = g.sum(acc);
    g.sync();
    return tot;
}

// in-place red-black Gauss-Seidel sweep; returns (sum red) + (sum black) of the pre-update res^2
template <class G>
__device__ __forceinline__ double sm_rbgs(const G &g, double *u, const double *rhs, int nx, int ny, double h, double c,
                                          bool want_norm)
{
    const double C = 4.0 + c * (h * h), h2 = h * h, w = 1.0 * (h2 / C);
    const DivH2 dh = make_div_h2(h2);
    double tot = 0.0;
    const int n = nx * ny;
    for (int colour = 0; colour < 2; ++colour) {
        double acc = 0.0;
        for (Idx2 q(g.rank(), g.size(), nx); q.p < n; q.next()) {
            const int p = q.p, i = q.i, j = q.j;
            if (i >= 1 && i <= nx - 2 && j >= 1 && j <= ny - 2 && ((i + j + colour) & 1) == 0) {
                const double r = div_h2(u[p + 1] + u[p - 1] + u[p + nx] + u[p - nx] - C * u[p], dh) - rhs[p];
                u[p] = u[p] + w * r;
                acc += r * r;
            }
        }
        if (want_norm) tot += g.sum(acc);
        g.sync();
    }
    return tot;
}

template <class G>
__device__ __forceinline__ double sm_sumsq(const G &g, const double *a, int n)
{
    double acc = 0.0;
    for (int p = g.rank(); p < n; p += g.size()) acc += a[p] * a[p];
    return g.sum(acc);
}
template <class G>
__device__ __forceinline__ double sm_dot(const G &g, const double *a, const double *b, int n)
{
    double acc = 0.0;
    for (int p = g.rank(); p < n; p += g.size()) acc += a[p] * b[p];
    return g.sum(acc);
}

// cg!(x_in, b, hx, hy, c, tol, Nmax)  krylov.jl:55-91.  work = 4*n doubles (r, p, ph, x). Returns sum r^2.
template <class G>
__device__ __forceinline__ double sm_cg(const G &g, double *x_in, const double *b, double *work, int nx, int ny, double hx,
                                        double hy, double c, double tol, int Nmax, int *iters_out)
{
    const int n = nx * ny;
    double *r = work, *p = work + n, *ph = work + 2 * n, *x = work + 3 * n;
    const double normb = sqrt(sm_sumsq(g, b, n));
    const double tolb = tol * normb;
    for (int q = g.rank(); q < n; q += g.size()) { r[q] = b[q]; p[q] = b[q]; ph[q] = b[q]; x[q] = 0.0; }
    g.sync();
    double rho = sm_dot(g, r, r, n);
    int it = 0;
    for (int k = 1; k <= Nmax; ++k) {
        it = k;
        for (int q = g.rank(); q < n; q += g.size()) {
            const int i = q % nx, j = q / nx;
            if (i >= 1 && i <= nx - 2 && j >= 1 && j <= ny - 2)
                ph[q] = ((p[q + 1] - 2 * p[q] + p[q - 1]) / (hx * hx) + (p[q + nx] - 2 * p[q] + p[q - nx]) / (hy * hy)) - c * p[q];
        }
        g.sync();
        const double alpha = rho / sm_dot(g, p, ph, n);
        for (int q = g.rank(); q < n; q += g.size()) { x[q] += alpha * p[q]; r[q] -= alpha * ph[q]; }
        g.sync();
        const double rr = sm_sumsq(g, r, n);
        if (sqrt(rr) < tolb) break;
        const double rho_old = rho;
        rho = rr;
        const double beta = rho / rho_old;
        for (int q = g.rank(); q < n; q += g.size()) p[q] = r[q] + beta * p[q];
        g.sync();
    }
    for (int q = g.rank(); q < n; q += g.size()) x_in[q] = x[q];
    g.sync();
    if (iters_out != nullptr && g.rank() == 0) *iters_out = it;
    return sm_sumsq(g, r, n);
}

// Exact threshold for the coarsest-level exit test: the reference stops when sqrt(ss/N) < T (multigrid.jl:150-156).
// Correctly rounded division and sqrt are monotone, so {ss : sqrt(fl(ss/N)) < T} = [0, S*) for one double S*; it is
// found once per solve (a few ulp steps around T*T*N) and the per-sweep test becomes `ss < S*` -- the same decisions
// without a division and a square root on the critical path of every sweep.
__device__ __forceinline__ double exit_threshold(double T, double N)
{
    if (!(T > 0.0)) return T == T ? 0.0 : T;  // T <= 0: never true (ss >= 0); NaN: comparisons stay false
    double s = T * T * N;
    if (!(s > 0.0) || s != s || s > 1.0e300) return s;  // underflow / overflow: keep the plain product
    for (int i = 0; i < 64 && sqrt(s / N) < T; ++i) s = __longlong_as_double(__double_as_longlong(s) + 1);
    for (int i = 0; i < 64; ++i) {
        const double sp = __longlong_as_double(__double_as_longlong(s) - 1);
        if (sqrt(sp / N) < T) break;
        s = sp;
    }
    return s;
}

// One Jacobi sweep of a tiny grid inside one warp with the points of each lane precomputed (M points per lane).
template <int M>
struct WarpPoints {
    int p[M];
    bool valid[M], interior[M], red[M];
    __device__ __forceinline__ WarpPoints(int nx, int ny)
    {
        const int lane = threadIdx.x & 31, n = nx * ny;
#pragma unroll
        for (int m = 0; m < M; ++m) {
            p[m] = lane + 32 * m;
            valid[m] = p[m] < n;
            const int j = p[m] / nx, i = p[m] - j * nx;
            interior[m] = valid[m] && i >= 1 && i <= nx - 2 && j >= 1 && j <= ny - 2;
            red[m] = ((i + j) & 1) == 0;
        }
    }
    // Branch-free (frame points evaluate the stencil at a safe interior index and discard it), so that the M dependent
    // chains of a lane overlap instead of running one after the other in separate divergent regions.
    __device__ __forceinline__ double sweep(const double *src, const double *rhs, double *dst, int nx, const Coef &k) const
    {
        double acc = 0.0;
#pragma unroll
        for (int m = 0; m < M; ++m) {
            const int q = interior[m] ? p[m] : nx + 1;
            const double own = valid[m] ? src[p[m]] : 0.0;
            const double r = ((src[q + 1] + src[q - 1] + src[q + nx] + src[q - nx] - k.C * src[q]) * k._h2 - rhs[q]);
            const double vn = src[q] + k.w * r;
            acc += interior[m] ? r * r : 0.0;
            if (valid[m]) dst[p[m]] = interior[m] ? vn : own;
        }
        return acc;
    }
    // One red-black Gauss-Seidel sweep in place (variant B): red half step, __syncwarp, black half step; returns this
    // lane's share of the pre-update res^2 of both colours. Same branch-free scheme.
    __device__ __forceinline__ double sweep_rb(double *u, const double *rhs, int nx, double C, double w, const DivH2 &dh) const
    {
        double acc = 0.0;
#pragma unroll
        for (int colour = 0; colour < 2; ++colour) {
#pragma unroll
            for (int m = 0; m < M; ++m) {
                const bool act = interior[m] && red[m] == (colour == 0);
                const int q = act ? p[m] : nx + 1;
                const double r = div_h2(u[q + 1] + u[q - 1] + u[q + nx] + u[q - nx] - C * u[q], dh) - rhs[q];
                const double vn = u[q] + w * r;
                acc += act ? r * r : 0.0;
                if (act) u[p[m]] = vn;
            }
            __syncwarp();
        }
        return acc;
    }
};

// Reduction of the per-lane res^2 of the (up to 8) sweeps of a batch, stored as accbuf[sweep][lane] by the whole warp:
// lanes 4b..4b+3 sum sweep b (8 loads and a 3-level add tree each, then 2 shuffle levels). Returns in every lane the
// index of the first sweep b < nb whose total is below sstar (-1: none) and, through tot_of_hit, that total (or the
// total of sweep nb-1 when none passes). All 32 lanes must call.
__device__ __forceinline__ int batch_first_hit(const double (*accbuf)[32], int nb, double sstar, double *tot_of_hit)
{
    const int lane = threadIdx.x & 31, rb = lane >> 2, rp = (lane & 3) * 8;
    const double *q = &accbuf[rb][rp];
    double t = ((q[0] + q[1]) + (q[2] + q[3])) + ((q[4] + q[5]) + (q[6] + q[7]));
    t += __shfl_xor_sync(0xffffffffu, t, 1);
    t += __shfl_xor_sync(0xffffffffu, t, 2);
    const unsigned ok = __ballot_sync(0xffffffffu, rb < nb && t < sstar);
    const int hit = ok ? ((__ffs(ok) - 1) >> 2) : -1;
    *tot_of_hit = __shfl_sync(0xffffffffu, t, 4 * (hit >= 0 ? hit : nb - 1));
    return hit;
}

// Coarsest Jacobi solve of a grid with 33..128 points in one warp (M points per lane, state in two shared-memory
// buffers). Like the register version below, the exit test of the reference (res_rms < tol_rhs after every sweep,
// multigrid.jl:150-156) is taken off the dependent chain: sweeps run speculatively in batches of 8, each sweep stores its
// per-lane res^2, one reduction per batch finds the first sweep that passes; the state at the start of the batch is kept
// in registers, and if the batch ran past the exit it is restored and the counted sweeps are redone (identical
// arithmetic). Returns sum res^2 of the last counted sweep; the solution ends up in u.
template <int M>
__device__ __noinline__ double warp_coarsest_jacobi(double *u, const double *rhs, double *tmp, int nx, int ny, const Coef &kref,
                                                    double sstar, int iters, int *sweeps_out)
{
    __shared__ double accbuf[8][32];
    const Coef k = kref;
    const WarpPoints<M> wp(nx, ny);
    const int n = nx * ny, lane = threadIdx.x & 31;
    double *cur = u, *oth = tmp;
    double tot = 0.0;
    int s = 0;
    for (;;) {
        double snap[M];
#pragma unroll
        for (int m = 0; m < M; ++m) snap[m] = wp.valid[m] ? cur[wp.p[m]] : 0.0;
        double *const start = cur;
        const int nb = min(8, iters - s);
#pragma unroll 1
        for (int b = 0; b < nb; ++b) {
            accbuf[b][lane] = wp.sweep(cur, rhs, oth, nx, k);
            __syncwarp();
            double *t_ = cur; cur = oth; oth = t_;
        }
        const int hit = batch_first_hit(accbuf, nb, sstar, &tot);
        __syncwarp();
        if (hit < 0 && s + nb < iters) { s += nb; continue; }
        const int last = hit >= 0 ? hit : nb - 1;  // the cap 20*coarse_solve_size ends the loop otherwise
        if (last != nb - 1) {  // the batch ran past the exit: restore its start state and redo the counted sweeps
            cur = start; oth = (start == u) ? tmp : u;
#pragma unroll
            for (int m = 0; m < M; ++m)
                if (wp.valid[m]) cur[wp.p[m]] = snap[m];
            __syncwarp();
            for (int b = 0; b <= last; ++b) {
                wp.sweep(cur, rhs, oth, nx, k);
                __syncwarp();
                double *t_ = cur; cur = oth; oth = t_;
            }
        }
        s += last + 1;
        break;
    }
    if (cur != u) {
        for (int q = lane; q < n; q += 32) u[q] = cur[q];
        __syncwarp();
    }
    if (sweeps_out != nullptr && lane == 0) *sweeps_out = s;
    return tot;
}

// The same for the red-black Gauss-Seidel smoother (variant B; sweeps are in place, so there is no second buffer).
template <int M>
__device__ __noinline__ double warp_coarsest_rbgs(double *u, const double *rhs, int nx, int ny, double h, double c, double sstar,
                                                  int iters, int *sweeps_out)
{
    __shared__ double accbuf[8][32];
    const WarpPoints<M> wp(nx, ny);
    const int lane = threadIdx.x & 31;
    const double C = 4.0 + c * (h * h), w = 1.0 * ((h * h) / C);
    const DivH2 dh = make_div_h2(h * h);
    double tot = 0.0;
    int s = 0;
    for (;;) {
        double snap[M];
#pragma unroll
        for (int m = 0; m < M; ++m) snap[m] = wp.valid[m] ? u[wp.p[m]] : 0.0;
        const int nb = min(8, iters - s);
#pragma unroll 1
        for (int b = 0; b < nb; ++b) accbuf[b][lane] = wp.sweep_rb(u, rhs, nx, C, w, dh);
        __syncwarp();
        const int hit = batch_first_hit(accbuf, nb, sstar, &tot);
        __syncwarp();
        if (hit < 0 && s + nb < iters) { s += nb; continue; }
        const int last = hit >= 0 ? hit : nb - 1;
        if (last != nb - 1) {  // the batch ran past the exit: restore its start state and redo the counted sweeps
#pragma unroll
            for (int m = 0; m < M; ++m)
                if (wp.valid[m]) u[wp.p[m]] = snap[m];
            __syncwarp();
            for (int b = 0; b <= last; ++b) wp.sweep_rb(u, rhs, nx, C, w, dh);
        }
        s += last + 1;
        break;
    }
    if (sweeps_out != nullptr && lane == 0) *sweeps_out = s;
    return tot;
}

// Coarsest solve of a grid with <= 32 points (5x5, the default coarse_solve_size) in the registers of one warp: one point
// per lane, neighbours by shuffle. Measured on the B200 (scripts/microbench): FP64 add/mul latency 8 cycles, a 64-bit
// shuffle 25, so one Jacobi sweep is a ~120-cycle dependent chain (4 shuffles + 8 FP64 operations + select) that nothing
// can shorten -- but everything else can be taken off it. The reference's exit test (res_rms < tol_rhs after every
// sweep, multigrid.jl:150-156) needs a warp-wide sum per sweep: sweeps therefore run speculatively in batches of 8, each
// sweep drops its per-lane res^2 and its state into shared memory (stores nobody waits for), and once per batch all 32
// lanes reduce the 8 sums together (4 lanes per sweep: 8 loads, a 3-level add tree, 2 shuffle levels, one ballot). If
// sweep j of the batch is the first to pass the test, the state after sweep j is simply read back. The arithmetic per
// sweep is that of the sequential loop, so solution and sweep count are identical to it.
// RB: red-black Gauss-Seidel (variant B) instead of damped Jacobi: two half steps per sweep.
template <bool RB>
struct RegSweep {
    double f, C, s2, kw;   // s2: 1/h^2 (Jacobi) or the exact reciprocal / h^2 itself (RB, see DivH2)
    bool interior, red, exact;
    int up, dn;
    struct Nb { double e, w, n, s; };
    __device__ __forceinline__ Nb gather(double v) const
    {
        Nb q;
        q.e = __shfl_down_sync(0xffffffffu, v, 1); q.w = __shfl_up_sync(0xffffffffu, v, 1);
        q.n = __shfl_sync(0xffffffffu, v, up); q.s = __shfl_sync(0xffffffffu, v, dn);
        return q;
    }
    __device__ __forceinline__ double update(const Nb &q, double v, bool active, double &acc) const
    {
        const double t = q.e + q.w + q.n + q.s - C * v;
        double r;
        if (RB) r = (exact ? t * s2 : t / s2) - f;   // (...) / h^2 - f   multigrid.jl:279-283
        else r = (t * s2 - f);                       // (...) * _h2 - f   multigrid.jl:183-187
        const double vn = v + kw * r;
        if (active) acc = r * r;
        return active ? vn : v;
    }
    __device__ __forceinline__ double sweep(double v, double &acc) const
    {
        acc = 0.0;
        if (!RB) return update(gather(v), v, interior, acc);
        v = update(gather(v), v, interior && red, acc);
        return update(gather(v), v, interior && !red, acc);
    }
};

#ifdef B2S_COARSEST_STAMPS
__device__ long long g_stamps[64];
#endif
template <bool RB>
__device__ __noinline__ double warp_coarsest_reg(double *u, const double *rhs, int nx, int ny, double h, double c, double sstar,
                                                 int iters, int *sweeps_out)
{
    __shared__ double accbuf[8][32], vbuf[8][32];
#ifdef B2S_COARSEST_STAMPS
    __shared__ long long s_stamps[64];
    int nst = 0;
#endif
    const int lane = threadIdx.x & 31, n = nx * ny;
    const bool valid = lane < n;
    const int j = lane / nx, i = lane - j * nx;
    RegSweep<RB> sw;
    sw.interior = valid && i >= 1 && i <= nx - 2 && j >= 1 && j <= ny - 2;
    sw.red = ((i + j) & 1) == 0;
    sw.up = lane + nx; sw.dn = lane - nx;
    sw.f = valid ? rhs[lane] : 0.0;
    sw.exact = true;
    if (RB) {
        const DivH2 dh = make_div_h2(h * h);
        sw.C = 4.0 + c * (h * h); sw.kw = 1.0 * ((h * h) / sw.C);
        sw.exact = dh.exact; sw.s2 = dh.exact ? dh.inv : dh.h2;
    } else {
        const Coef k = make_coef(h, c, 4.0 / 5.0);
        sw.C = k.C; sw.s2 = k._h2; sw.kw = k.w;
    }
    double v = valid ? u[lane] : 0.0;
    double tot = 0.0;
    int s = 0;
    for (;;) {
        const int nb = min(8, iters - s);
#ifdef B2S_COARSEST_STAMPS
        if (nst < 60) s_stamps[nst++] = clock64();
#endif
        if (nb == 8) {
#pragma unroll
            for (int b = 0; b < 8; ++b) {
                double a;
                v = sw.sweep(v, a);
                accbuf[b][lane] = a;
                vbuf[b][lane] = v;
            }
        } else {
#pragma unroll 1
            for (int b = 0; b < nb; ++b) {
                double a;
                v = sw.sweep(v, a);
                accbuf[b][lane] = a;
                vbuf[b][lane] = v;
            }
        }
#ifdef B2S_COARSEST_STAMPS
        if (nst < 60) s_stamps[nst++] = clock64();
#endif
        __syncwarp();
        int hit = batch_first_hit(accbuf, nb, sstar, &tot);  // first sweep of the batch that passes res_rms < tol_rhs
        __syncwarp();
        if (hit < 0 && s + nb < iters) { s += nb; continue; }
        if (hit < 0) hit = nb - 1;  // the cap 20*coarse_solve_size ends the loop
        v = vbuf[hit][lane];  // state after the last counted sweep
        s += hit + 1;
        break;
    }
    if (valid) u[lane] = v;
    __syncwarp();
#ifdef B2S_COARSEST_STAMPS
    if (lane == 0) for (int q = 0; q < 64; ++q) g_stamps[q] = q < nst ? s_stamps[q] : 0;
#endif
    if (sweeps_out != nullptr && lane == 0) *sweeps_out = s;
    return tot;
}

// Coarsest-level solve (multigrid.jl:145-167). u in place; tmp = scratch of the level; work = CG scratch.
template <class G>
__device__ __forceinline__ double sm_coarsest(const G &g, double *u, const double *rhs, double *tmp, double *work, int nx,
                                              int ny, double h, const CoarseArgs &a, double c, double tol, int *sweeps_out)
{
    const int iters = 20 * a.coarse_solve_size;
    const int n = nx * ny;
    double ss = 0.0;
    if (a.coarse_solver == B2S_COARSE_JACOBI) {
        const double N = (double)nx * ny;
        const double tol_rhs = tol * sqrt(sm_sumsq(g, rhs, n) / N);
        const double sstar = exit_threshold(tol_rhs, N);  // res_rms < tol_rhs  <=>  ss < sstar
        const Coef k = make_coef(h, c, 4.0 / 5.0);
        if (g.size() == 32 && n <= 32 && nx >= 3) {  // register-resident solve (shuffles wrap correctly only for nx >= 2)
            if (a.smoother == B2S_SMOOTH_RBGS) return warp_coarsest_reg<true>(u, rhs, nx, ny, h, c, sstar, iters, sweeps_out);
            return warp_coarsest_reg<false>(u, rhs, nx, ny, h, c, sstar, iters, sweeps_out);
        }
        if (a.smoother == B2S_SMOOTH_RBGS && g.size() == 32 && n <= 128 && nx >= 3 && ny >= 3) {
            if (n <= 64) return warp_coarsest_rbgs<2>(u, rhs, nx, ny, h, c, sstar, iters, sweeps_out);
            if (n <= 96) return warp_coarsest_rbgs<3>(u, rhs, nx, ny, h, c, sstar, iters, sweeps_out);
            return warp_coarsest_rbgs<4>(u, rhs, nx, ny, h, c, sstar, iters, sweeps_out);
        }
        if (a.smoother == B2S_SMOOTH_JACOBI && g.size() == 32 && n <= 128) {
            if (n <= 32) return warp_coarsest_jacobi<1>(u, rhs, tmp, nx, ny, k, sstar, iters, sweeps_out);
            if (n <= 64) return warp_coarsest_jacobi<2>(u, rhs, tmp, nx, ny, k, sstar, iters, sweeps_out);
            if (n <= 96) return warp_coarsest_jacobi<3>(u, rhs, tmp, nx, ny, k, sstar, iters, sweeps_out);
            return warp_coarsest_jacobi<4>(u, rhs, tmp, nx, ny, k, sstar, iters, sweeps_out);
        }
        int sweeps = 0;
        double *src = u, *dst = tmp;
        for (int s = 1; s <= iters; ++s) {
            if (a.smoother == B2S_SMOOTH_RBGS) {
                ss = sm_rbgs(g, u, rhs, nx, ny, h, c, true);
            } else {
                ss = sm_jacobi(g, src, rhs, dst, nx, ny, k, true);
                double *t = src; src = dst; dst = t;
            }
            ++sweeps;
            if (ss < sstar) break;
        }
        if (src != u) {  // odd number of Jacobi sweeps: result sits in tmp
            for (int p = g.rank(); p < n; p += g.size()) u[p] = src[p];
            g.sync();
        }
        if (sweeps_out != nullptr && g.rank() == 0) *sweeps_out = sweeps;
    } else {
        ss = sm_cg(g, u, rhs, work, nx, ny, h, h, c, tol, iters, sweeps_out);
    }
    return ss;
}

// The chain of one thread block over levels that live in ITS shared memory: downward leg, coarsest solve, upward leg
// (multigrid.jl:121-167 for every level of the block). U[0] and F[0] hold the unknown and the right-hand side of the first
// level on entry; U[0] holds the result on exit. Returns the sum of res^2 of the last sweep of the first level when
// `top_norm` (valid on every thread). All threads of the block must call.
__device__ __forceinline__ double coarse_chain(const CoarseArgs &a, double *const *U, double *const *F, double *const *T,
                                               const LevelCoef *slev, double *red, double c, double tol, int apply_bcs,
                                               bool top_norm)
{
    const MGCall *cp = a.cp;
    const bool fw = a.restriction == B2S_RESTRICT_FW;
    BlockGroup bg{red};
    auto coef_of = [&](int l) {
        Coef k;
        k.C = slev[l].C; k._h2 = slev[l]._h2; k.w = slev[l].wJ;
        return k;
    };
    double last_ss = 0.0;
    // ---- downward leg ---------------------------------------------------------------------------------------
    for (int l = 0; l + 1 < a.nlev; ++l) {
        const int nx = a.nx[l], ny = a.ny[l], nxc = a.nx[l + 1], nyc = a.ny[l + 1];
        const double h = slev[l].h;
        const Coef k = coef_of(l);
        if (a.smoother == B2S_SMOOTH_RBGS) {
            sm_rbgs(bg, U[l], F[l], nx, ny, h, c, false);
            sm_rbgs(bg, U[l], F[l], nx, ny, h, c, false);
        } else {
            sm_jacobi(bg, U[l], F[l], T[l], nx, ny, k, false);
            sm_jacobi(bg, T[l], F[l], U[l], nx, ny, k, false);
        }
        const Coef kr = k;  // the residual uses C and 1/h^2 only
        for (Idx2 q(threadIdx.x, blockDim.x, nxc); q.p < nxc * nyc; q.next()) {
            const int p = q.p, J = q.j;
            int I = q.i;
            if (apply_bcs) {
                if (I == 0) I = 1;
                else if (I == nxc - 1) I = nxc - 2;
            }
            F[l + 1][p] = coarse_value<true>(U[l], F[l], nx, I, J, nxc, nyc, fw, kr);
            U[l + 1][p] = 0.0;
        }
        __syncthreads();
        coarse_stamp(a);
    }
    // ---- coarsest solve ---------------------------------------------------------------------------------------
    {
        const int l = a.nlev - 1;
        const int nx = a.nx[l], ny = a.ny[l];
        const double h = slev[l].h;
        if (nx * ny <= 1024) {
            if (threadIdx.x < 32) {
                WarpGroup wg;
                last_ss = sm_coarsest(wg, U[l], F[l], T[l], U[a.nlev], nx, ny, h, a, c, tol, cp->coarse_sweeps);
            }
            __syncthreads();
        } else {
            last_ss = sm_coarsest(bg, U[l], F[l], T[l], U[a.nlev], nx, ny, h, a, c, tol, cp->coarse_sweeps);
        }
        coarse_stamp(a);
    }
    // ---- upward leg -------------------------------------------------------------------------------------------
    for (int l = a.nlev - 2; l >= 0; --l) {
        const int nx = a.nx[l], ny = a.ny[l], nxc = a.nx[l + 1], nyc = a.ny[l + 1];
        const double h = slev[l].h;
        const Coef k = coef_of(l);
        for (Idx2 q(threadIdx.x, blockDim.x, nx); q.p < nx * ny; q.next())
            U[l][q.p] = U[l][q.p] - prolong_value_bc(U[l + 1], nxc, nyc, nx, q.i, q.j, apply_bcs);
        __syncthreads();
        const bool top = (l == 0 && top_norm);
        if (a.smoother == B2S_SMOOTH_RBGS) {
            sm_rbgs(bg, U[l], F[l], nx, ny, h, c, false);
            last_ss = sm_rbgs(bg, U[l], F[l], nx, ny, h, c, top);
        } else {
            sm_jacobi(bg, U[l], F[l], T[l], nx, ny, k, false);
            last_ss = sm_jacobi(bg, T[l], F[l], U[l], nx, ny, k, top);
        }
        coarse_stamp(a);
    }
    return last_ss;
}

__global__ void __launch_bounds__(1024) mg_coarse_kernel(const CoarseArgs a)
{
    extern __shared__ double sm[];
    __shared__ double red[32];
    const MGCall *cp = a.cp;
    if (cp->done) return;
    const double c = cp->c, tol = cp->tol;
    const int apply_bcs = cp->apply_bcs;
    // carve shared memory: per level u, rhs, tmp; then CG scratch for the coarsest
    double *U[kMaxLevels], *F[kMaxLevels], *T[kMaxLevels];
    {
        double *ptr = sm;
        for (int l = 0; l < a.nlev; ++l) {
            const int n = a.nx[l] * a.ny[l];
            U[l] = ptr; ptr += n;
            F[l] = ptr; ptr += n;
            T[l] = ptr; ptr += n;
        }
        U[a.nlev] = ptr;  // CG scratch (4 * n_coarsest)
    }
    const double *rhs_g = a.rhs_in != nullptr ? a.rhs_in : cp->rhs;
    double *u_g = a.u_io != nullptr ? a.u_io : cp->u;
    __shared__ LevelCoef slev[kMaxLevels];  // per-level constants of this call (host-computed), fetched once up front
    {
        if (threadIdx.x < a.nlev) slev[threadIdx.x] = *level_consts(cp, a.level0 + threadIdx.x);
        const int n = a.nx[0] * a.ny[0];
        for (int p = threadIdx.x; p < n; p += blockDim.x) {
            F[0][p] = rhs_g[p];
            U[0][p] = a.u_is_input ? u_g[p] : 0.0;
        }
        __syncthreads();
    }
    if (a.prof != nullptr && threadIdx.x == 0) a.prof[0] = 0;
    coarse_stamp(a);
    const double last_ss = coarse_chain(a, U, F, T, slev, red, c, tol, apply_bcs, a.sumsq_out != nullptr);
    {
        const int n = a.nx[0] * a.ny[0];
        for (int p = threadIdx.x; p < n; p += blockDim.x) u_g[p] = U[0][p];
        if (a.sumsq_out != nullptr && threadIdx.x == 0) *a.sumsq_out = last_ss;
    }
    __syncthreads();
    coarse_stamp(a);
}

// ---------------------------------------------------------------------------------------------------------------
// Thread-block-cluster kernel for the latency-bound middle of the V-cycle (variant A: damped Jacobi + injection).
// Levels of ~1e3 .. 2e4 points (33^2 .. 129^2 at the 1025^2 bench shape) are chains of dependent sweeps that no amount of
// SMs speeds up: as separate kernels each level costs two dependent launches (~3.5 us each). Here ONE cluster of NC
// thread blocks keeps those levels in DISTRIBUTED SHARED MEMORY -- every block owns a band of rows of every such level
// (u, tmp with one halo row on each side, rhs) -- and walks down and up the hierarchy as a DATAFLOW between neighbours:
//   * a sweep computes the band's first and last row first; the thread that produces such a point also sends it into
//     the neighbour block's halo row with st.async (shared::cluster store that completes transaction bytes on an mbarrier
//     in the RECEIVING block: no read latency, no fence, no flag);
//   * at the end of the sweep every block waits on its own mbarrier for the two rows it is owed (they were sent at the
//     START of the neighbours' sweep, so the wait is normally over when it begins) -- there is no cluster-wide barrier
//     after the kernel prologue (measured: barrier.cluster costs ~970 cycles incl. its gpu-scope fence, x18 per cycle);
//   * write-after-read on a halo row needs no extra synchronisation: a neighbour can only be one sweep ahead, because
//     its next sweep needs the row this block sends after it has read the halo in question.
// The levels below (<= 17^2) are the one-block chain of mg_coarse_kernel, run by block 0 on the right-hand side all
// blocks send into its shared memory; it broadcasts its correction back the same way.
// Every point is produced by the arithmetic of the unfused kernels: bit-identical results.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kMidThreads = 512;
constexpr int kMidMaxDist = 6;
struct MidArgs {
    CoarseArgs ser;             // the serial tail run by block 0 (ser.level0 = global index of its first level)
    int level0;                 // global index of the first (finest) distributed level
    int ndist;                  // distributed levels
    int nx[kMidMaxDist], ny[kMidMaxDist];
    int base;                   // rows per block on the LAST distributed level: ny[ndist-1] - 1 == NC * base
    const double *rhs_in;       // global rhs of level0 (written by the fused downward kernel of the level above)
    double *u_out;              // global unknown of level0 (read by the fused upward kernel of the level above)
};

// shared-memory doubles per block: bands of the distributed levels + block 0's serial levels + the copy of the first
// serial level's correction every block receives (every block uses the same offsets)
__host__ __device__ inline size_t mid_dist_doubles(const MidArgs &m)
{
    size_t n = 0;
    for (int d = 0; d < m.ndist; ++d) {
        const size_t rows = ((size_t)m.base << (m.ndist - 1 - d)) + 1;
        n += (2 * (rows + 2) + rows) * (size_t)m.nx[d];
    }
    return n;
}

__device__ __forceinline__ uint32_t mapa_u32(uint32_t smem_addr, uint32_t cta_rank)
{
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(cta_rank));
    return r;
}
// 8-byte store into another block's shared memory that completes 8 transaction bytes on an mbarrier of THAT block
__device__ __forceinline__ void st_async_f64(uint32_t remote_addr, double v, uint32_t remote_mbar)
{
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.f64 [%0], %1, [%2];" ::"r"(remote_addr), "d"(v),
                 "r"(remote_mbar)
                 : "memory");
}

struct MidLevelS {
    int nx, ny, r0, rows, offU, offT, offF, per;
};
struct MidSend {  // where a band's first / last row goes: shared::cluster addresses in the neighbour blocks
    uint32_t lo_addr, lo_bar, hi_addr, hi_bar;
    bool has_lo, has_hi;
};

// one damped-Jacobi sweep over this block's band of a level: src/dst have a halo row on each side (local row lr + 1),
// F has the band only. Band rows are visited first row, last row, then the rest; points of the first / last row are also
// sent to the neighbours.
__device__ __forceinline__ void mid_sweep(const double *__restrict__ src, const double *__restrict__ F, double *__restrict__ dst,
                                          int nx, int ny, int r0, int rows, const Coef &k, const MidSend &sd)
{
    const int n = rows * nx;
    for (Idx2 q(threadIdx.x, kMidThreads, nx); q.p < n; q.next()) {
        const int i = q.i;
        const int lr = q.j == 0 ? 0 : (q.j == 1 ? rows - 1 : q.j - 1);
        const int j = r0 + lr;
        const int s = (lr + 1) * nx + i;
        double v = src[s];
        if (i >= 1 && i <= nx - 2 && j >= 1 && j <= ny - 2) {
            const double r = ((src[s + 1] + src[s - 1] + src[s + nx] + src[s - nx] - k.C * v) * k._h2 - F[lr * nx + i]);
            v = v + k.w * r;
        }
        dst[s] = v;
        if (lr == 0 && sd.has_lo) st_async_f64(sd.lo_addr + 8u * (uint32_t)i, v, sd.lo_bar);
        if (lr == rows - 1 && sd.has_hi) st_async_f64(sd.hi_addr + 8u * (uint32_t)i, v, sd.hi_bar);
    }
}

// bilinear prolongation value at fine point (i, j) from a coarse array held with local row crow0 at index 0; the coarse
// boundary ring counts as 0 (prolong_value_bc with 32-bit indices)
__device__ __forceinline__ double mid_prolong(const double *__restrict__ ecl, int crow0, int nxc, int nyc, int nx, int i, int j,
                                              int apply_bcs)
{
    if (apply_bcs) {
        if (i == 0) i = 1;
        else if (i == nx - 1) i = nx - 2;
    }
    const int I = i >> 1, J = j >> 1;
    const bool io = i & 1, jo = j & 1;
    auto at = [&](int II, int JJ) -> double {
        return (II >= 1 && II <= nxc - 2 && JJ >= 1 && JJ <= nyc - 2) ? ecl[(JJ - crow0) * nxc + II] : 0.0;
    };
    if (!io && !jo) return at(I, J);
    if (io && !jo) return 0.5 * at(I, J) + 0.5 * at(I + 1, J);
    if (!io && jo) return 0.5 * at(I, J) + 0.5 * at(I, J + 1);
    return ((0.25 * at(I, J) + 0.25 * at(I + 1, J)) + 0.25 * at(I, J + 1)) + 0.25 * at(I + 1, J + 1);
}

__global__ void __launch_bounds__(kMidThreads) mg_mid_cluster_kernel(const MidArgs m)
{
    namespace cgr = cooperative_groups;
    cgr::cluster_group cluster = cgr::this_cluster();
    extern __shared__ double sm[];
    __shared__ double red[32];
    __shared__ LevelCoef slev[kMaxLevels];
    __shared__ MidLevelS lv[kMidMaxDist];
    // [0] / [1]: halo rows of the sweeps that write tmp / u; [2]: block 0 only, the gathered rhs of the first serial level;
    // [3]: the broadcast correction of the first serial level
    __shared__ __align__(8) uint64_t bars[4];
    const MGCall *cp = m.ser.cp;
    if (cp->done) return;  // the same value in every block of the cluster
    const int NC = (int)cluster.num_blocks(), rank = (int)cluster.block_rank();
    const int tid = threadIdx.x;
    const double c = cp->c, tol = cp->tol;
    const int apply_bcs = cp->apply_bcs;
    const int nd = m.ndist;
    const int nxs = m.ser.nx[0], nys = m.ser.ny[0];
    // ---- carve-up (identical offsets in every block) ----------------------------------------------------------
    int off = 0;
    for (int d = 0; d < nd; ++d) {
        const int per = m.base << (nd - 1 - d), rowsmax = per + 1;
        if (tid == 0) {
            MidLevelS L;
            L.nx = m.nx[d]; L.ny = m.ny[d]; L.per = per;
            L.r0 = rank * per;
            L.rows = (rank == NC - 1 ? m.ny[d] : (rank + 1) * per) - L.r0;
            L.offU = off; L.offT = off + (rowsmax + 2) * m.nx[d]; L.offF = off + 2 * (rowsmax + 2) * m.nx[d];
            lv[d] = L;
        }
        off += (2 * (rowsmax + 2) + rowsmax) * m.nx[d];
    }
    double *U[kMaxLevels], *F[kMaxLevels], *T[kMaxLevels];
    {
        double *ptr = sm + off;
        for (int l = 0; l < m.ser.nlev; ++l) {
            const int n = m.ser.nx[l] * m.ser.ny[l];
            U[l] = ptr; ptr += n;
            F[l] = ptr; ptr += n;
            T[l] = ptr; ptr += n;
        }
        U[m.ser.nlev] = ptr;  // CG scratch of the coarsest level (4 x its size)
        ptr += 4 * m.ser.nx[m.ser.nlev - 1] * m.ser.ny[m.ser.nlev - 1];
        T[m.ser.nlev] = ptr;  // this block's copy of block 0's correction on the first serial level
    }
    double *Ecopy = T[m.ser.nlev];
    if (tid < nd + m.ser.nlev) slev[tid] = *level_consts(cp, m.level0 + tid);
    if (tid == 0) {
        for (int b_ = 0; b_ < 4; ++b_) mbar_init(&bars[b_], 1);
        fence_mbar_init();
        // the two one-shot transfers are armed right away
        if (rank == 0) mbar_arrive_expect_tx(&bars[2], (uint32_t)(nxs * nys * 8));
        mbar_arrive_expect_tx(&bars[3], (uint32_t)(nxs * nys * 8));
    }
    __syncthreads();
    // ---- load: rhs band of the first level, u = 0 ------------------------------------------------------------------
    {
        const MidLevelS L = lv[0];
        const double *g = m.rhs_in + (size_t)L.nx * L.r0;
        double *Fd = sm + L.offF, *Ud = sm + L.offU;
        for (int p = tid; p < L.rows * L.nx; p += kMidThreads) Fd[p] = g[p];
        for (int p = tid; p < (L.rows + 2) * L.nx; p += kMidThreads) Ud[p] = 0.0;
        if (rank == 0)
            for (int p = tid; p < nxs * nys; p += kMidThreads) U[0][p] = 0.0;
    }
    cluster.sync();  // every block of the cluster runs and has initialised its mbarriers before anyone sends to it
    if (rank == 0) {
        if (m.ser.prof != nullptr && tid == 0) m.ser.prof[0] = 0;
        coarse_stamp(m.ser);
    }
    uint32_t phase[2] = {0u, 0u};
    const uint32_t bars_addr = smem_u32(bars), sm_addr = smem_u32(sm);
    // two sweeps of level d: u -> tmp -> u, halo rows exchanged with the neighbours
    auto two_sweeps = [&](int d) {
        const MidLevelS L = lv[d];
        Coef k;
        k.C = slev[d].C; k._h2 = slev[d]._h2; k.w = slev[d].wJ;
        double *Ud = sm + L.offU, *Td = sm + L.offT;
        const double *Fd = sm + L.offF;
        const bool has_lo = rank > 0, has_hi = rank < NC - 1;
        const uint32_t bytes = (uint32_t)((has_lo ? 1 : 0) + (has_hi ? 1 : 0)) * (uint32_t)L.nx * 8u;
#pragma unroll
        for (int sw = 0; sw < 2; ++sw) {
            const int dstoff = sw == 0 ? L.offT : L.offU;
            MidSend sd;
            sd.has_lo = has_lo; sd.has_hi = has_hi;
            // my first row -> the lower neighbour's halo row above its band (its band has `per` rows: it is never the last block)
            sd.lo_addr = has_lo ? mapa_u32(sm_addr + 8u * (uint32_t)(dstoff + (L.per + 1) * L.nx), (uint32_t)(rank - 1)) : 0u;
            sd.lo_bar = has_lo ? mapa_u32(bars_addr + 8u * (uint32_t)sw, (uint32_t)(rank - 1)) : 0u;
            // my last row -> the upper neighbour's halo row 0
            sd.hi_addr = has_hi ? mapa_u32(sm_addr + 8u * (uint32_t)dstoff, (uint32_t)(rank + 1)) : 0u;
            sd.hi_bar = has_hi ? mapa_u32(bars_addr + 8u * (uint32_t)sw, (uint32_t)(rank + 1)) : 0u;
            if (tid == 0 && bytes != 0u) mbar_arrive_expect_tx(&bars[sw], bytes);
            mid_sweep(sw == 0 ? Ud : Td, Fd, sw == 0 ? Td : Ud, L.nx, L.ny, L.r0, L.rows, k, sd);
            if (bytes != 0u) {
                mbar_wait(&bars[sw], phase[sw]);
                phase[sw] ^= 1u;
            }
            __syncthreads();
            if (rank == 0) coarse_stamp(m.ser);
        }
    };
    // injected residual of the smoothed u at coarse point (I, J) of the next level, from this block's band of level d
    auto coarse_rhs = [&](const MidLevelS &L, const Coef &k, int I, int J, int nxc, int nyc) -> double {
        if (apply_bcs) {
            if (I == 0) I = 1;
            else if (I == nxc - 1) I = nxc - 2;
        }
        if (I < 1 || I > nxc - 2 || J < 1 || J > nyc - 2) return 0.0;
        const int nx = L.nx;
        const int s = (2 * J - L.r0 + 1) * nx + 2 * I;
        const double *u = sm + L.offU;
        return ((u[s + 1] + u[s - 1] + u[s + nx] + u[s - nx] - k.C * u[s]) * k._h2 - sm[L.offF + (2 * J - L.r0) * nx + 2 * I]);
    };
    // ---- downward leg over the distributed levels -----------------------------------------------------------------
    for (int d = 0; d < nd; ++d) {
        two_sweeps(d);
        const MidLevelS L = lv[d];
        Coef k;
        k.C = slev[d].C; k._h2 = slev[d]._h2; k.w = slev[d].wJ;
        if (d + 1 < nd) {
            const MidLevelS Lc = lv[d + 1];
            double *Fc = sm + Lc.offF, *Uc = sm + Lc.offU;
            for (Idx2 q(tid, kMidThreads, Lc.nx); q.p < Lc.rows * Lc.nx; q.next())
                Fc[q.p] = coarse_rhs(L, k, q.i, Lc.r0 + q.j, Lc.nx, Lc.ny);
            for (int p = tid; p < (Lc.rows + 2) * Lc.nx; p += kMidThreads) Uc[p] = 0.0;
            __syncthreads();
        } else {  // into block 0's copy of the first serial level: every block sends its rows (block 0 included)
            const int c0 = L.r0 >> 1, c1 = rank == NC - 1 ? nys : ((L.r0 + L.rows) >> 1);
            const uint32_t f0 = mapa_u32(smem_u32(F[0]), 0u), b2 = mapa_u32(bars_addr + 16u, 0u);
            for (Idx2 q(tid, kMidThreads, nxs); q.p < (c1 - c0) * nxs; q.next())
                st_async_f64(f0 + 8u * (uint32_t)((c0 + q.j) * nxs + q.i), coarse_rhs(L, k, q.i, c0 + q.j, nxs, nys), b2);
        }
        if (rank == 0) coarse_stamp(m.ser);
    }
    // ---- the serial tail in block 0, its correction broadcast to every block ---------------------------------------
    if (rank == 0) {
        mbar_wait(&bars[2], 0u);
        __syncthreads();
        coarse_chain(m.ser, U, F, T, slev + nd, red, c, tol, apply_bcs, false);
        for (int p = tid; p < nxs * nys; p += kMidThreads) {
            const double v = U[0][p];
            for (int r = 0; r < NC; ++r)
                st_async_f64(mapa_u32(smem_u32(Ecopy) + 8u * (uint32_t)p, (uint32_t)r), v, mapa_u32(bars_addr + 24u, (uint32_t)r));
        }
    }
    mbar_wait(&bars[3], 0u);
    __syncthreads();
    if (rank == 0) coarse_stamp(m.ser);
    // ---- upward leg ------------------------------------------------------------------------------------------------------
    for (int d = nd - 1; d >= 0; --d) {
        const MidLevelS L = lv[d];
        const double *ecl;
        int crow0, nxc, nyc;
        if (d + 1 < nd) {
            const MidLevelS Lc = lv[d + 1];
            ecl = sm + Lc.offU; crow0 = Lc.r0 - 1; nxc = Lc.nx; nyc = Lc.ny;
        } else {
            ecl = Ecopy; crow0 = 0; nxc = nxs; nyc = nys;
        }
        // u -= P(e) on the band and on its two halo rows (their coarse neighbours are in the coarse band's halo rows)
        const int jlo = rank == 0 ? L.r0 : L.r0 - 1, jhi = rank == NC - 1 ? L.r0 + L.rows : L.r0 + L.rows + 1;
        double *Ud = sm + L.offU;
        for (Idx2 q(tid, kMidThreads, L.nx); q.p < (jhi - jlo) * L.nx; q.next()) {
            const int j = jlo + q.j;
            const int s = (j - L.r0 + 1) * L.nx + q.i;
            Ud[s] = Ud[s] - mid_prolong(ecl, crow0, nxc, nyc, L.nx, q.i, j, apply_bcs);
        }
        __syncthreads();
        if (rank == 0) coarse_stamp(m.ser);
        two_sweeps(d);
    }
    {
        const MidLevelS L = lv[0];
        double *g = m.u_out + (size_t)L.nx * L.r0;
        const double *u = sm + L.offU + L.nx;
        for (int p = tid; p < L.rows * L.nx; p += kMidThreads) g[p] = u[p];
    }
    // every block has waited for everything that was sent to it: nobody's shared memory is written after it exits
}

// tol_rhs = tol * sqrt(sum(rhs.^2)/(nx*ny)) and the exit threshold of a global-memory coarsest Jacobi solve
__global__ void mg_coarse_loop_init_kernel(CoarseLoop *loop, const double *sumsq_rhs, const MGCall *cp, double N, int iters)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        const double tol_rhs = cp->tol * sqrt(*sumsq_rhs / N);
        loop->sstar = exit_threshold(tol_rhs, N);
        loop->sweeps = 0;
        loop->done = 0;
        loop->iters = iters;
    }
}

// One red-black sweep of a global-memory coarsest solve is done (two mg_rbgs_kernel launches): res_rms < tol_rhs -> break
// (multigrid.jl:150-156) with sum res^2 = red + black.
__global__ void mg_coarse_loop_rb_step_kernel(CoarseLoop *loop, double *sumsq)
{
    if (threadIdx.x != 0 || blockIdx.x != 0 || loop->done) return;
    const double total = sumsq[2] + sumsq[3];
    sumsq[0] = total;
    const int sw = loop->sweeps + 1;
    loop->sweeps = sw;
    if (total < loop->sstar || sw >= loop->iters) loop->done = 1;
}

// Stand-alone CG for grids that fit into shared memory (test/krylov.jl shape: 66^2).
__global__ void __launch_bounds__(1024) mg_cg_smem_kernel(double *x, const double *b, double hx, double hy, double c, double tol,
                                                          int nmax, int nx, int ny, double *ss_out, int *iters_out)
{
    extern __shared__ double sm[];
    __shared__ double red[32];
    BlockGroup bg{red};
    const int n = nx * ny;
    double *bs = sm, *xs = sm + n, *work = sm + 2 * n;
    for (int p = threadIdx.x; p < n; p += blockDim.x) { bs[p] = b[p]; xs[p] = x[p]; }
    __syncthreads();
    const double ss = sm_cg(bg, xs, bs, work, nx, ny, hx, hy, c, tol, nmax, iters_out);
    for (int p = threadIdx.x; p < n; p += blockDim.x) x[p] = xs[p];
    if (threadIdx.x == 0) *ss_out = ss;
}

}  // namespace b2s
