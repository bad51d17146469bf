/*
 * b200stencil.h -- C ABI of libb200stencil.so (hand-written CUDA for sm_100a).
 *
 * This is the drop-in boundary for the two stencil hot paths of ntselepidis/FinalProjectRepo.jl.
 * The reference has no FFI (its kernels are Julia functions JIT-compiled by CUDA.jl through
 * ParallelStencil), so each entry point below names the reference function / call site it replaces
 * (file:line relative to the reference repository).  The reference-side binding is a Julia `ccall`
 * (julia/); the executable harness in this repository binds the same symbols with
 * Python ctypes (finalprojectrepo.jl_b200/_capi.py).
 *
 * Conventions
 *   - All arrays are dense Float64, column-major (x fastest), exactly as a Julia Array/CuArray stores
 *     them: element (ix,iy,iz) 1-based  <->  offset (ix-1) + nx*((iy-1) + ny*(iz-1)).
 *   - "dev" pointers are CUDA device pointers (Julia: CuPtr{Float64}); "host" pointers are host memory.
 *   - L0 functions (kernel level) never allocate user-visible memory, are asynchronous on `stream`
 *     (a cudaStream_t passed as void*; NULL = the legacy default stream) unless they return a scalar to
 *     the host, in which case they synchronise that stream (mirrors @synchronize, multigrid.jl:65,
 *     krylov.jl:51).
 *   - L1 functions work on opaque handles that own device memory, CUDA graphs and peer mappings.
 *     One host thread per handle; calls on one handle are not re-entrant.
 *   - Every function returns an int status (B2S_OK == 0). b2s_last_error() gives a thread-local message.
 *     Non-convergence is NOT an error (the reference only warns, multigrid.jl:78-82, or silently stops
 *     at iter_max, part1_kernel_programming.jl:179).
 *   - There is no CPU fallback: without a CUDA device every compute entry point returns
 *     B2S_ERR_NO_DEVICE / B2S_ERR_CUDA.
 */
#ifndef B200STENCIL_H
#define B200STENCIL_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2S_VERSION 100

/* ---- status codes -------------------------------------------------------------------------------- */
#define B2S_OK 0
#define B2S_ERR_BAD_SIZE 1        /* multigrid.jl:95-97 "ERROR:not a power of 2"; asserts multigrid.jl:45-46 */
#define B2S_ERR_BAD_ARG 2
#define B2S_ERR_CUDA 3
#define B2S_ERR_NOT_IMPLEMENTED 4 /* execution_policy == serial -> error(), multigrid.jl:233-236, krylov.jl:46-49 */
#define B2S_ERR_NO_DEVICE 5
#define B2S_ERR_STATE 6

const char *b2s_last_error(void);
int b2s_version(void);
int b2s_device_count(int *count);
/* Frees the library's lazily allocated per-device scratch (used by L0 reductions). */
int b2s_shutdown(void);

/* ==================================================================================================
 * PATH 1 -- 3-D dual-time / pseudo-transient diffusion
 * ================================================================================================== */

/* Halo semantics across z-slabs. */
#define B2S_HALO_REFERENCE_LAG2 0 /* update_halo!(Htau) on the buffer just READ: part1_kernel_programming.jl:182,187 */
#define B2S_HALO_CONSISTENT 1     /* halo of the buffer just written: part1_array_programming.jl:66-67 */
/* Dirichlet-face handling at set-up. */
#define B2S_BC_LITERAL 0 /* part1_utils.jl:14-34 as executed with ImplicitGlobalGrid's 0-based coords */
#define B2S_BC_PROPER 1  /* zero the faces on the physical boundary */
/* Kernel variants of the PT step (all bit-identical in their results). */
#define B2S_KERNEL_AUTO 0
#define B2S_KERNEL_DIRECT 1 /* one thread per cell pair, neighbours through L1/L2 (reference-shaped; correctness anchor) */
#define B2S_KERNEL_TMA 2    /* 2.5-D z-marching, TMA-staged planes in an mbarrier ring, register z-queue */
/* Arithmetic of the PT update. */
#define B2S_ARITH_KERNEL 0 /* part1_kernel_programming.jl:12-20,46-58: fluxes with D_dx = D/dx, residual with _dx = 1/dx */
#define B2S_ARITH_ARRAY 1  /* part1_array_programming.jl:9-18: q = D*d(Htau)/dx, divisions by dx, dy, dz, dt, Htau updated in
                            * place (its frame keeps Ht's values); direct kernel only */

/*
 * L0: one launch of the fused flux/residual/update kernel.
 * Replaces: @parallel diffusion_3D_step_tau(Ht, Htau, Htau2, dHdtau, dtau, _dt, _dx, _dy, _dz, D_dx, D_dy, D_dz)
 *           part1_kernel_programming.jl:46-58 (call sites :181 and :186) and its shared-memory twin :75-97.
 * Interior cells only; boundary cells of Htau2 / dHdtau are left untouched.
 *   dHdtau_dev   nullable: when NULL the residual field is not materialised.
 *   sumsq_dev    nullable: receives sum over interior cells of (R*norm_scale)^2, i.e. the local part of
 *                dist_norm_L2(residual_H*dt)^2 (part1_kernel_programming.jl:191, part1_utils.jl:36-40) with
 *                norm_scale = dt; deterministic (fixed-order two-stage reduction).
 */
int b2s_diffusion3d_step_tau(const double *Ht_dev, const double *Htau_dev, double *Htau2_dev, double *dHdtau_dev,
                             int nx, int ny, int nz, double dtau, double _dt, double _dx, double _dy, double _dz,
                             double D_dx, double D_dy, double D_dz, double norm_scale, double *sumsq_dev,
                             int kernel_variant, void *stream);

/* L1: the whole solver of diffusion_3D_kernel_programming (part1_kernel_programming.jl:99-228). */
typedef struct b2s_diff3d b2s_diff3d;

typedef struct {
    int nx, ny, nz;          /* LOCAL grid of one z-slab (= one reference MPI rank), incl. the 2-cell overlap */
    int nslabs_total;        /* dims = (1,1,nslabs_total): number of z-slabs of the global grid */
    int slab_begin;          /* first slab hosted by this handle */
    int slab_count;          /* slabs hosted by this handle (in-process: nslabs_total; one process per GPU: 1) */
    const int *devices;      /* slab_count CUDA device ordinals (may repeat: several slabs on one GPU) */
    int halo_mode;           /* B2S_HALO_* */
    int bc_mode;             /* B2S_BC_* */
    int scale_physical_size; /* part1_kernel_programming.jl:110-114 */
    int kernel_variant;      /* B2S_KERNEL_* */
    int batch;               /* PT iterations enqueued per host poll; 0 = automatic */
    int dimx, dimy;          /* general Cartesian decomposition dims = (dimx, dimy, nslabs_total/(dimx*dimy)) like
                              * ImplicitGlobalGrid's init_global_grid (part1_kernel_programming.jl:117); 0 or 1 = z-slabs.
                              * Rank r has coords (r / (dimy*dimz), (r / dimz) % dimy, r % dimz) (MPI Cartesian order).
                              * update_halo! then runs as separate plane copies in ImplicitGlobalGrid's order x, y, z after
                              * every step (in-process or one process per GPU). */
    int arithmetic;          /* B2S_ARITH_* (0 = the kernel-programming version) */
} b2s_diff3d_config;

/* Derived numerics of part1_kernel_programming.jl:117-152. */
typedef struct {
    double lx, ly, lz, dx, dy, dz, dt, dtau;
    double total_N; /* prod(dims)*nx*ny*nz, :124 */
    int nx_g, ny_g, nz_g;
} b2s_diff3d_params;

int b2s_diff3d_create(b2s_diff3d **h, const b2s_diff3d_config *cfg);
int b2s_diff3d_destroy(b2s_diff3d *h);
int b2s_diff3d_get_params(const b2s_diff3d *h, b2s_diff3d_params *out);
/* Same numbers without a handle or a GPU (pure host arithmetic; used to size host buffers before create). */
int b2s_diff3d_params_for(const b2s_diff3d_config *cfg, b2s_diff3d_params *out);

/* Ht = init_local_gaussian; apply_boundary_conditions!; Htau = copy(Ht); Htau2 = 0
 * (part1_kernel_programming.jl:137-142, part1_utils.jl:1-34). Evaluated on the host like the reference. */
int b2s_diff3d_init_gaussian(b2s_diff3d *h);
/* Same, from caller-provided LOCAL arrays (host, nx*ny*nz per hosted slab, slab-major).
 * One process per GPU: every rank initialises, then b2s_diff3d_ipc_connect, then a barrier across the ranks (the caller's:
 * MPI / torch.distributed) before the first iteration; a later re-initialisation needs the same barrier after it. */
int b2s_diff3d_set_initial(b2s_diff3d *h, const double *Ht_host);

/* Multi-process z-slabs (one process per GPU): export the CUDA IPC handles of this handle's halo mailboxes and
 * reduction slots, exchange the blobs out of band (torch.distributed / MPI), then connect.
 * blob size = b2s_diff3d_ipc_blob_bytes(). all_blobs = nslabs_total blobs in slab order. */
size_t b2s_diff3d_ipc_blob_bytes(void);
int b2s_diff3d_ipc_export(b2s_diff3d *h, void *blob_out);
int b2s_diff3d_ipc_connect(b2s_diff3d *h, const void *all_blobs, int nblobs);
/* There is no separate initial halo exchange: the reference has none either (update_halo! only runs inside the PT loop,
 * part1_kernel_programming.jl:182,187) and the fused exchange reproduces it from the first iteration on. */

/* while err > tol && iter < iter_max  (part1_kernel_programming.jl:177-193) -- device-resident loop:
 * the exit test runs on the GPU, the host polls a flag once per batch.  iters/err are the reference's
 * iter_inner and err at exit. */
int b2s_diff3d_solve_timestep(b2s_diff3d *h, double tol, int iter_max, int *iters, double *err);
/* Fixed number of PT iterations (benchmark / per-iteration parity); err_hist nullable, host, n entries. */
int b2s_diff3d_iterate(b2s_diff3d *h, int n, double *err_hist);
/* Ht .= Htau  (part1_kernel_programming.jl:203) */
int b2s_diff3d_advance_time(b2s_diff3d *h);
/* Whole time loop "for t in 0:dt:ttot-dt" (:166-204); iters_per_step nullable (capacity cap). Returns the
 * number of outer steps in *nsteps. */
int b2s_diff3d_run(b2s_diff3d *h, double ttot, double tol, int iter_max, int *iters_per_step, int cap, int *nsteps);
/* Field access: which = 0 Ht, 1 Htau (current), 2 Htau2 (other buffer). Host buffer nx*ny*nz; slab is the
 * GLOBAL slab index and must be hosted by this handle. Halo planes are materialised as the reference holds them. */
int b2s_diff3d_get_field(b2s_diff3d *h, int slab, int which, double *host_out);
/* gather!(Array(Ht), H_g) for the hosted slabs: host buffer nx*ny*(nz*slab_count) (:144,223). */
int b2s_diff3d_gather(b2s_diff3d *h, double *H_g_host);
/* Device pointers of a hosted slab (for zero-copy interop): which as in get_field. */
int b2s_diff3d_device_ptr(b2s_diff3d *h, int slab, int which, double **dev_out);
/* Host<->device transfer of the evolving state through the public API (used by the end-to-end benchmark).
 * upload_state starts a NEW JOB on the handle: Ht := host data, Htau := Ht, Htau2 := 0, ping-pong parity reset -- the
 * state of a fresh handle after set_initial, so a job's result does not depend on what ran before (call it for every
 * hosted slab; one process per GPU: all ranks must have returned from the previous job's last call first, which
 * solve_timestep / iterate guarantee because their final norm needs every rank). download_state copies Htau out. */
int b2s_diff3d_upload_state(b2s_diff3d *h, int slab, const double *Ht_host);
int b2s_diff3d_download_state(b2s_diff3d *h, int slab, double *Htau_host);
/* Pipelined variant for back-to-back jobs: the result is copied aside on the device (an extra nx*ny*nz array, allocated
 * on first use) and transferred from there on a separate copy stream, so the handle is free for the next job at once.
 * Htau_host (pinned) is valid after b2s_diff3d_sync(). */
int b2s_diff3d_download_state_async(b2s_diff3d *h, int slab, double *Htau_host);
/* Double-buffered upload for back-to-back jobs: _async copies the NEXT job's state into a staging array on the copy
 * stream (it overlaps with the iterations of the current job; the staging array, nx*ny*nz doubles, is allocated on
 * first use); _commit makes it the current state (Ht := Htau := staged) on the compute stream once it has arrived. */
int b2s_diff3d_upload_state_async(b2s_diff3d *h, int slab, const double *Ht_host);
int b2s_diff3d_commit_upload(b2s_diff3d *h, int slab);
int b2s_diff3d_sync(b2s_diff3d *h);
/* Bookkeeping for gpu_launches / timing: kernels launched so far; device time (ms, CUDA events on the
 * launching stream) of the last solve_timestep / iterate call, max over hosted slabs. */
int b2s_diff3d_stats(const b2s_diff3d *h, long long *kernel_launches, double *last_call_ms);

/* ==================================================================================================
 * PATH 2 -- 2-D matrix-free geometric multigrid for (lap - c) u = f, CG, boundary conditions
 * ================================================================================================== */

#define B2S_COARSE_JACOBI 0 /* CoarseSolver_t jacobi, multigrid.jl:10-13 */
#define B2S_COARSE_CG 1     /* conjugate_gradient */
#define B2S_SMOOTH_JACOBI 0 /* variant A (reference): damped Jacobi, alpha = 0.8, multigrid.jl:245-258 */
#define B2S_SMOOTH_RBGS 1   /* variant B (north-star extension): red-black Gauss-Seidel, alpha = 1 */
#define B2S_RESTRICT_INJECT 0 /* variant A: multigrid.jl:330-337 */
#define B2S_RESTRICT_FW 1     /* variant B: full weighting */
/* ExecutionPolicy_t, part2_utils.jl:4-8 (parallel and parallel_shmem map to the same CUDA kernels) */
#define B2S_POLICY_SERIAL 0
#define B2S_POLICY_PARALLEL 1
#define B2S_POLICY_PARALLEL_SHMEM 2

/* ---- L0, 1:1 with the reference's call sites ------------------------------------------------------- */
/* residual_2DPoisson_wrapper!(u, f, h, c, res, policy)  multigrid.jl:173-238.  Interior only. */
int b2s_residual2d(const double *u_dev, const double *f_dev, double h, double c, double *res_dev, int nx, int ny,
                   int policy, void *stream);
/* r_rms = iteration_2DPoisson!(u, f, h, c, res, policy; alpha)  multigrid.jl:245-258 (in place; synchronises). */
int b2s_iteration2d(double *u_dev, const double *f_dev, double h, double c, double *res_dev, int nx, int ny,
                    double alpha, int policy, double *r_rms_host, void *stream);
/* NOT provided: iteration_2DPoisson_gs! (multigrid.jl:269-297), the serial lexicographic Gauss-Seidel sweep. It has no call
 * site in the reference (dead code upstream) and is inherently sequential; the CPU oracle restates it (orc_gs2d_lex) so that
 * variant B's on-the-fly residual definition can be checked against it, the library does not. */
/* Variant B smoother: one red-black Gauss-Seidel sweep in place (alpha = 1); r_rms from pre-update residuals. */
int b2s_rbgs2d(double *u_dev, const double *f_dev, double h, double c, int nx, int ny, double *r_rms_host,
               void *stream);
/* restrict_wrapper!(fine, coarse, apply_BCs, policy)  multigrid.jl:330-358 (zero + injection + Neumann). */
int b2s_restrict_inject2d(const double *fine_dev, double *coarse_dev, int nx, int ny, int apply_bcs, void *stream);
int b2s_restrict_fw2d(const double *fine_dev, double *coarse_dev, int nx, int ny, int apply_bcs, void *stream);
/* prolongate_wrapper!(coarse, fine, apply_BCs, policy)  multigrid.jl:403-472, as a deterministic gather with the
 * reference's CPU arrival order (no atomics). nx, ny are the FINE sizes. */
int b2s_prolongate2d(const double *coarse_dev, double *fine_dev, int nx, int ny, int apply_bcs, void *stream);
/* matrix_free_matvec_prod_wrapper!(p, hx, hy, c, p_hat)  krylov.jl:7-52 */
int b2s_matvec2d(const double *T_dev, double hx, double hy, double c, double *out_dev, int nx, int ny, int policy,
                 void *stream);
/* apply_boundary_conditions!/_dirichlet!/_neumann!  part2_utils.jl:21-39. kind: 0 both, 1 Dirichlet, 2 Neumann */
int b2s_apply_bc2d(double *T_dev, int nx, int ny, int kind, void *stream);
/* Reductions / vector updates of cg! and the residual checks (krylov.jl:57-90, multigrid.jl:53,150,252):
 * deterministic warp-shuffle + block + fixed-order final stage; results to host (synchronise). */
int b2s_dot(const double *x_dev, const double *y_dev, size_t n, double *out_host, void *stream);
int b2s_sumsq(const double *x_dev, size_t n, double *out_host, void *stream);
int b2s_axpy(double alpha, const double *x_dev, double *y_dev, size_t n, void *stream); /* y += alpha*x */
int b2s_xpby(const double *x_dev, double beta, double *y_dev, size_t n, void *stream);  /* y = x + beta*y */

/* ---- L1: solver handle ------------------------------------------------------------------------------ */
typedef struct b2s_mg b2s_mg;

typedef struct {
    int nx, ny;            /* finest grid, (2^k*lambda)+1 per side (multigrid.jl:25-29,95-97) */
    int coarse_solve_size; /* MGOpt.coarse_solve_size, multigrid.jl:17,21 (default 5) */
    int coarse_solver;     /* B2S_COARSE_* */
    int smoother;          /* B2S_SMOOTH_* */
    int restriction;       /* B2S_RESTRICT_* */
    int device;            /* CUDA device ordinal */
    int use_graph;         /* 1: V-cycle captured once in a CUDA graph and replayed */
    int smem_levels;       /* 1: all levels that fit are collapsed into one shared-memory-resident kernel */
    int fuse_sweeps;       /* temporal blocking on the fine levels (2 sweeps + transfer operator per kernel; variant A
                              only, bit-identical to the unfused kernels): 0 off, 1 automatic (streaming y-marching
                              kernels on large levels, shared-memory tile kernels on small ones), 2 tiles everywhere,
                              3 block-wide streaming (one column per thread) everywhere, 4 block-wide streaming (two
                              columns per thread, producer warp) everywhere, 5 one-warp-per-strip streaming everywhere */
} b2s_mg_config;

/* preallocate_buffers(nx, ny)  multigrid.jl:25-38 (+ level table, graphs). */
int b2s_mg_create(b2s_mg **h, const b2s_mg_config *cfg);
int b2s_mg_destroy(b2s_mg *h);
/* r_rms = MGsolve_2DPoisson!(u, f, h, c, tol, niters, apply_BCs; opt, prealloc_dict)  multigrid.jl:41-84.
 * u_dev is updated in place. ncycles = V-cycles executed; rel_hist nullable (host, niters): r_rms/f_rms per cycle
 * (the reference's verbose print, :68). */
int b2s_mg_solve(b2s_mg *h, double *u_dev, const double *f_dev, double hgrid, double c, double tol, int niters,
                 int apply_bcs, double *r_rms, int *ncycles, double *rel_hist);
/* res_rms = Vcycle_2DPoisson!(u, rhs, h, c, tol, ...)  multigrid.jl:91-170: exactly one V-cycle. */
int b2s_mg_vcycle(b2s_mg *h, double *u_dev, const double *rhs_dev, double hgrid, double c, double tol,
                  int apply_bcs, double *res_rms);
/* Fixed number of V-cycles without the exit test (benchmark). Device time of the call in *ms (CUDA events). */
int b2s_mg_cycles(b2s_mg *h, double *u_dev, const double *f_dev, double hgrid, double c, double tol, int ncycles,
                  int apply_bcs, double *r_rms_last, double *ms);
/* MG-preconditioned CG (north-star extension; the reference has no such solver): solves
 * (lap - c) u = f on the interior (frame of u = Dirichlet data) by CG on matrix_free_matvec_prod! (krylov.jl:7-13) with
 * ONE V-cycle of this handle (started from zero) as preconditioner. Needs a symmetric cycle: restriction must be
 * B2S_RESTRICT_FW (with injection the V-cycle is not symmetric and CG stagnates -> B2S_ERR_BAD_ARG).
 * Exit: sqrt(sum r^2/(nx ny)) < tol * (the same norm of the initial residual). iters = CG iterations = V-cycles. */
int b2s_mg_pcg_solve(b2s_mg *h, double *u_dev, const double *f_dev, double hgrid, double c, double tol, int maxit,
                     double *r_rms, int *iters);
/* The same with a selectable stopping criterion: relative to the initial residual (as above), or MGsolve_2DPoisson!'s
 * r_rms < tol * f_rms with f_rms over all entries of f (multigrid.jl:53,70-75) -- what navier_stokes_2D needs to swap
 * solvers without changing the accuracy it asks for. */
#define B2S_PCG_TOL_INITIAL_RESIDUAL 0
#define B2S_PCG_TOL_RHS 1
int b2s_mg_pcg_solve2(b2s_mg *h, double *u_dev, const double *f_dev, double hgrid, double c, double tol, int maxit,
                      int tol_mode, double *r_rms, int *iters);
/* Measurement aid (no reference counterpart): average device time (ms, CUDA events on the handle's stream, `reps` back-to-back
 * launches after a warm-up) of every kernel of one V-cycle of the fused paths in isolation: per global-memory level l the
 * downward kernel (2 sweeps + residual + restriction) and the upward kernel (prolongation + correction + 2 sweeps), and the
 * one kernel that handles all levels below them (ms_tail). nlevels = number of global-memory levels; the ms_* / level_*
 * arrays need capacity 24. */
int b2s_mg_profile_kernels(b2s_mg *h, double *u_dev, const double *f_dev, double hgrid, double c, int reps, int *nlevels,
                           double *ms_down, double *ms_up, double *ms_tail, int *level_nx, int *level_ny);
/* Sweeps / iterations the coarsest-level solver used in the last V-cycle. */
int b2s_mg_last_coarse_sweeps(const b2s_mg *h, int *sweeps);
int b2s_mg_stats(const b2s_mg *h, long long *kernel_launches, double *last_call_ms);

/* res_rms = cg!(x_in, b, hx, hy, c, tol, Nmax; execution_policy)  krylov.jl:55-91. */
int b2s_cg_solve(double *x_dev, const double *b_dev, double hx, double hy, double c, double tol, int nmax, int nx,
                 int ny, int policy, double *res_rms, int *iters, void *stream);

/* ---- Navier-Stokes driver around the solves (part2.jl:140-262) --------------------------------------- */
typedef struct b2s_ns2d b2s_ns2d;
typedef struct {
    double k, Ra, Pr; /* SimIn_t, part2.jl:30-46 */
    int nx, ny;
    double ttot, beta;
    int niters;
    double tol, a_dif, a_adv;
} b2s_ns2d_params;
typedef struct {
    double dt;
    int cycles_S, cycles_T, cycles_W;
    double r_S, r_T, r_W;
} b2s_ns2d_stepinfo;

int b2s_ns2d_create(b2s_ns2d **h, const b2s_ns2d_params *p, const b2s_mg_config *mg);
int b2s_ns2d_destroy(b2s_ns2d *h);
/* Solver of the two Dirichlet solves of a step (S: part2.jl:187, W: :226). The reference iterates plain V-cycles
 * (B2S_NS_SOLVER_VCYCLE, the parity default). B2S_NS_SOLVER_MG_PCG (BASELINE configs[3], north-star extension) solves them
 * with MG-preconditioned CG under the same stopping criterion; it needs a handle whose multigrid configuration restricts
 * with full weighting (symmetric cycle). The T solve (:221) applies boundary conditions inside the cycle -- not an SPD
 * system -- and always iterates V-cycles. cycles_S / cycles_W of the step info then count CG iterations (= V-cycles). */
#define B2S_NS_SOLVER_VCYCLE 0
#define B2S_NS_SOLVER_MG_PCG 1
int b2s_ns2d_set_solver(b2s_ns2d *h, int solver);
/* which: 0 T, 1 W, 2 S; host arrays nx*ny column-major */
int b2s_ns2d_set_field(b2s_ns2d *h, int which, const double *host);
int b2s_ns2d_get_field(b2s_ns2d *h, int which, double *host);
/* init_array!(M, cosine, h, width)  part2.jl:58-63 (host evaluation like the reference) */
int b2s_ns2d_init_cosine(b2s_ns2d *h, int which);
/* One pass of the while-loop body part2.jl:181-250. */
int b2s_ns2d_step(b2s_ns2d *h, b2s_ns2d_stepinfo *info);
/* aux fields of the last step for parity checks: 0 vx, 1 vy, 2 Ra_dTdx, 3 dT2, 4 dW2 */
int b2s_ns2d_get_aux(b2s_ns2d *h, int which, double *host);

#ifdef __cplusplus
}
#endif
#endif /* B200STENCIL_H */
