/*
 * ORACLE -- TEST INFRASTRUCTURE ONLY.  Not shipped, not on the product path.
 *
 * CPU restatement (plain C + OpenMP, no FMA contraction: build with
 * -ffp-contract=off) of the reference's 3-D dual-time / pseudo-transient
 * diffusion solver, including an emulation of its MPI ranks so that the
 * published multi-rank iteration counts can be reproduced on one host.
 *
 * Reference lines followed (all under /root/reference):
 *   scripts-part1/part1_kernel_programming.jl:12-20   flux macros @qx/@qy/@qz
 *   scripts-part1/part1_kernel_programming.jl:46-58   diffusion_3D_step_tau
 *   scripts-part1/part1_kernel_programming.jl:99-228  solver loop, constants, accounting
 *   scripts-part1/part1_utils.jl:1-12                 init_local_gaussian
 *   scripts-part1/part1_utils.jl:14-34                apply_boundary_conditions!
 *   scripts-part1/part1_utils.jl:36-40                dist_norm_L2
 *   scripts-part1/part1_array_programming.jl:9-18,61-83  array-programming version (orc_diff3d_set_array)
 * Un-vendored upstream behaviour restated (ImplicitGlobalGrid.jl, unpinned):
 *   overlap 2, nx_g = dims*(n-2)+2, x_g = (coords*(n-2)+i)*dx (0-based i),
 *   update_halo! per axis in order x,y,z on whole planes, 0-based coords.
 *
 * Parity is PINNED (see tests/test_oracle_diffusion.py): published point
 * values to 17 digits, published single- and multi-rank iteration counts,
 * and test/reftest-files/test_1.bson.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this file's shared object.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_HALO_REFERENCE_LAG2 0 /* update_halo!(Htau): the buffer just READ  (part1_kernel_programming.jl:182,187) */
#define ORC_HALO_CONSISTENT 1     /* exchange the buffer just WRITTEN (part1_array_programming.jl:67 semantics) */
#define ORC_BC_LITERAL 0          /* coords[d]==1 / coords[d]==dims[d] with 0-based coords (part1_utils.jl:14-34) */
#define ORC_BC_PROPER 1           /* zero the faces that lie on the physical boundary */

typedef struct {
    int nx, ny, nz;      /* local size per rank */
    int dims[3];         /* rank grid */
    int nranks;
    int halo_mode, bc_mode;
    int unfused_norm;    /* 1: materialise R*dt and (.)^2 temporaries like the reference (timing structure) */
    int array;           /* 1: the array-programming version (in-place update of Htau, flux arrays, divisions) */
    double *qx, *qy, *qz; /* its flux arrays (nx-1)(ny-2)(nz-2), (nx-2)(ny-1)(nz-2), (nx-2)(ny-2)(nz-1) */
    double lx, ly, lz, dx, dy, dz, dt, dtau;
    double _dt, _dx, _dy, _dz, D_dx, D_dy, D_dz;
    double total_N;
    double **Ht, **A, **B, **R; /* per rank */
    double *tmp1, *tmp2;        /* temporaries for the un-fused norm */
    long iters_total;
} orc_diff3d;

static size_t cells(const orc_diff3d *s) { return (size_t)s->nx * s->ny * s->nz; }

static void rank_coords(const orc_diff3d *s, int r, int c[3])
{
    /* MPI_Cart_create ordering is row-major: last dimension varies fastest. */
    c[2] = r % s->dims[2];
    c[1] = (r / s->dims[2]) % s->dims[1];
    c[0] = r / (s->dims[2] * s->dims[1]);
}
static int coords_rank(const orc_diff3d *s, const int c[3])
{
    return (c[0] * s->dims[1] + c[1]) * s->dims[2] + c[2];
}

void orc_diff3d_destroy(orc_diff3d *s)
{
    if (!s) return;
    for (int r = 0; r < s->nranks; ++r) {
        if (s->Ht) free(s->Ht[r]);
        if (s->A) free(s->A[r]);
        if (s->B) free(s->B[r]);
        if (s->R) free(s->R[r]);
    }
    free(s->Ht); free(s->A); free(s->B); free(s->R); free(s->tmp1); free(s->tmp2);
    free(s->qx); free(s->qy); free(s->qz);
    free(s);
}

/* part1_kernel_programming.jl:100-152 + part1_utils.jl:1-34 */
orc_diff3d *orc_diff3d_create(int nx, int ny, int nz, int dimx, int dimy, int dimz,
                              int halo_mode, int bc_mode, int scale_physical_size, int unfused_norm)
{
    orc_diff3d *s = (orc_diff3d *)calloc(1, sizeof(*s));
    s->nx = nx; s->ny = ny; s->nz = nz;
    s->dims[0] = dimx; s->dims[1] = dimy; s->dims[2] = dimz;
    s->nranks = dimx * dimy * dimz;
    s->halo_mode = halo_mode; s->bc_mode = bc_mode; s->unfused_norm = unfused_norm;
    const double D = 1.0;
    if (scale_physical_size) { s->lx = dimx * 10.0; s->ly = dimy * 10.0; s->lz = dimz * 10.0; }
    else { s->lx = s->ly = s->lz = 10.0; }
    const int nxg = dimx * (nx - 2) + 2, nyg = dimy * (ny - 2) + 2, nzg = dimz * (nz - 2) + 2;
    s->dx = s->lx / nxg; s->dy = s->ly / nyg; s->dz = s->lz / nzg;
    s->total_N = (double)s->nranks * nx * ny * nz;
    s->dt = 0.2;
    double m = fmin(s->dx, fmin(s->dy, s->dz));
    s->dtau = m * m / D / 8.1;
    s->_dt = 1.0 / s->dt; s->_dx = 1.0 / s->dx; s->_dy = 1.0 / s->dy; s->_dz = 1.0 / s->dz;
    s->D_dx = D / s->dx; s->D_dy = D / s->dy; s->D_dz = D / s->dz;

    size_t n = cells(s);
    s->Ht = (double **)calloc(s->nranks, sizeof(double *));
    s->A = (double **)calloc(s->nranks, sizeof(double *));
    s->B = (double **)calloc(s->nranks, sizeof(double *));
    s->R = (double **)calloc(s->nranks, sizeof(double *));
    if (unfused_norm) {
        s->tmp1 = (double *)malloc(n * sizeof(double));
        s->tmp2 = (double *)malloc(n * sizeof(double));
    }
    const double cx = s->lx / 2, cy = s->ly / 2, cz = s->lz / 2;
    for (int r = 0; r < s->nranks; ++r) {
        int c[3]; rank_coords(s, r, c);
        double *Ht = (double *)malloc(n * sizeof(double));
        s->Ht[r] = Ht;
        s->A[r] = (double *)malloc(n * sizeof(double));
        s->B[r] = (double *)calloc(n, sizeof(double)); /* Htau2 = @zeros */
        s->R[r] = (double *)calloc(n, sizeof(double)); /* residual_H = @zeros */
#pragma omp parallel for schedule(static)
        for (int k = 0; k < nz; ++k)
            for (int j = 0; j < ny; ++j)
                for (int i = 0; i < nx; ++i) {
                    double xg = (double)(c[0] * (nx - 2) + i) * s->dx;
                    double yg = (double)(c[1] * (ny - 2) + j) * s->dy;
                    double zg = (double)(c[2] * (nz - 2) + k) * s->dz;
                    double ax = xg + s->dx / 2 - cx, ay = yg + s->dy / 2 - cy, az = zg + s->dz / 2 - cz;
                    Ht[(size_t)i + (size_t)nx * (j + (size_t)ny * k)] =
                        2 * exp(-1.0 * ((ax * ax + ay * ay) + az * az));
                }
        /* apply_boundary_conditions!(Ht, coords, dims) */
        for (int d = 0; d < 3; ++d) {
            int zero_lo, zero_hi;
            if (bc_mode == ORC_BC_LITERAL) { zero_lo = (c[d] == 1); zero_hi = (c[d] == s->dims[d]); }
            else { zero_lo = (c[d] == 0); zero_hi = (c[d] == s->dims[d] - 1); }
            int nd[3] = {nx, ny, nz};
            for (int side = 0; side < 2; ++side) {
                if (!(side == 0 ? zero_lo : zero_hi)) continue;
                int fixed = side == 0 ? 0 : nd[d] - 1;
                for (int k = 0; k < nz; ++k)
                    for (int j = 0; j < ny; ++j)
                        for (int i = 0; i < nx; ++i) {
                            int idx[3] = {i, j, k};
                            if (idx[d] == fixed) Ht[(size_t)i + (size_t)nx * (j + (size_t)ny * k)] = 0.0;
                        }
            }
        }
        memcpy(s->A[r], Ht, n * sizeof(double)); /* Htau = copy(Ht) */
    }
    return s;
}

/* part1_kernel_programming.jl:46-58, evaluation order exactly as written. */
static void step_tau_rank(const orc_diff3d *s, const double *restrict Ht, const double *restrict A,
                          double *restrict B, double *restrict R)
{
    const int nx = s->nx, ny = s->ny, nz = s->nz;
    const size_t sy = nx, sz = (size_t)nx * ny;
    const double D_dx = s->D_dx, D_dy = s->D_dy, D_dz = s->D_dz;
    const double _dx = s->_dx, _dy = s->_dy, _dz = s->_dz, _dt = s->_dt, dtau = s->dtau;
#pragma omp parallel for schedule(static)
    for (int k = 1; k < nz - 1; ++k)
        for (int j = 1; j < ny - 1; ++j) {
            size_t p0 = (size_t)nx * (j + (size_t)ny * k);
            for (int i = 1; i < nx - 1; ++i) {
                size_t p = p0 + i;
                double c = A[p];
                double r = ((-D_dx * (A[p + 1] - c)) - (-D_dx * (c - A[p - 1]))) * _dx +
                           ((-D_dy * (A[p + sy] - c)) - (-D_dy * (c - A[p - sy]))) * _dy +
                           ((-D_dz * (A[p + sz] - c)) - (-D_dz * (c - A[p - sz]))) * _dz +
                           (c - Ht[p]) * _dt;
                R[p] = r;
                B[p] = c - dtau * r;
            }
        }
}

/* Switches the solver to diffusion_3D_array_programming (part1_array_programming.jl:20-92): update_halo!(Htau) after the
 * in-place update (:66-67), i.e. consistent halos. Call right after create. */
void orc_diff3d_set_array(orc_diff3d *s)
{
    s->array = 1;
    s->halo_mode = ORC_HALO_CONSISTENT;
    s->qx = (double *)calloc((size_t)(s->nx - 1) * (s->ny - 2) * (s->nz - 2), sizeof(double));
    s->qy = (double *)calloc((size_t)(s->nx - 2) * (s->ny - 1) * (s->nz - 2), sizeof(double));
    s->qz = (double *)calloc((size_t)(s->nx - 2) * (s->ny - 2) * (s->nz - 1), sizeof(double));
}

/* part1_array_programming.jl:9-18 with the statements as whole-array operations (the order in which the course's
 * array-programming model defines them): all of qx, qy, qz from the current Htau, then dHdtau on the inner points, then
 * the in-place update @inn(Htau) += dHdtau*dtau. R receives dHdtau at the inner points (its frame stays 0), so that
 * dist_norm_L2(dHdt*dt) (:68) is the same sum as in the kernel version. */
static void array_step_rank(const orc_diff3d *s, const double *restrict Ht, double *restrict H, double *restrict R)
{
    const int nx = s->nx, ny = s->ny, nz = s->nz;
    const size_t sy = nx, sz = (size_t)nx * ny;
    const double D = 1.0, dx = s->dx, dy = s->dy, dz = s->dz, dt = s->dt, dtau = s->dtau;
    double *qx = s->qx, *qy = s->qy, *qz = s->qz;
    const size_t ax = nx - 1, bx = nx - 2, ay = ny - 2, by = ny - 1;
    /* @all(qx) = D * @d_xi(Htau) / dx  : d_xi = differences in x of the y,z-inner points */
#pragma omp parallel for schedule(static)
    for (int k = 0; k < nz - 2; ++k)
        for (int j = 0; j < ny - 2; ++j)
            for (int i = 0; i < nx - 1; ++i) {
                size_t p = (size_t)i + sy * (j + 1) + sz * (k + 1);
                qx[i + ax * (j + ay * k)] = D * (H[p + 1] - H[p]) / dx;
            }
#pragma omp parallel for schedule(static)
    for (int k = 0; k < nz - 2; ++k)
        for (int j = 0; j < ny - 1; ++j)
            for (int i = 0; i < nx - 2; ++i) {
                size_t p = (size_t)(i + 1) + sy * j + sz * (k + 1);
                qy[i + bx * (j + by * k)] = D * (H[p + sy] - H[p]) / dy;
            }
#pragma omp parallel for schedule(static)
    for (int k = 0; k < nz - 1; ++k)
        for (int j = 0; j < ny - 2; ++j)
            for (int i = 0; i < nx - 2; ++i) {
                size_t p = (size_t)(i + 1) + sy * (j + 1) + sz * k;
                qz[i + bx * (j + ay * k)] = D * (H[p + sz] - H[p]) / dz;
            }
#pragma omp parallel for schedule(static)
    for (int k = 0; k < nz - 2; ++k)
        for (int j = 0; j < ny - 2; ++j)
            for (int i = 0; i < nx - 2; ++i) {
                size_t p = (size_t)(i + 1) + sy * (j + 1) + sz * (k + 1);
                double dH = -(H[p] - Ht[p]) / dt +
                            ((qx[(i + 1) + ax * (j + ay * k)] - qx[i + ax * (j + ay * k)]) / dx +
                             (qy[i + bx * ((j + 1) + by * k)] - qy[i + bx * (j + by * k)]) / dy +
                             (qz[i + bx * (j + ay * (k + 1))] - qz[i + bx * (j + ay * k)]) / dz);
                R[p] = dH;
                H[p] = H[p] + dH * dtau;
            }
}

/* sum((R*dt).^2) over the whole local array; per-plane partials combined in plane order. */
static double sumsq_rank(const orc_diff3d *s, const double *restrict R)
{
    const int nz = s->nz;
    const size_t plane = (size_t)s->nx * s->ny;
    const double dt = s->dt;
    double *part = (double *)malloc(nz * sizeof(double));
    if (s->unfused_norm) {
        /* Reference structure (part1_kernel_programming.jl:191, part1_utils.jl:37):
         * two temporaries are materialised, then summed. */
        size_t n = cells(s);
#pragma omp parallel for schedule(static)
        for (size_t p = 0; p < n; ++p) s->tmp1[p] = R[p] * dt;
#pragma omp parallel for schedule(static)
        for (size_t p = 0; p < n; ++p) s->tmp2[p] = s->tmp1[p] * s->tmp1[p];
#pragma omp parallel for schedule(static)
        for (int k = 0; k < nz; ++k) {
            double a = 0.0;
            const double *q = s->tmp2 + plane * k;
            for (size_t p = 0; p < plane; ++p) a += q[p];
            part[k] = a;
        }
    } else {
#pragma omp parallel for schedule(static)
        for (int k = 0; k < nz; ++k) {
            double a = 0.0;
            const double *q = R + plane * k;
            for (size_t p = 0; p < plane; ++p) { double v = q[p] * dt; a += v * v; }
            part[k] = a;
        }
    }
    double tot = 0.0;
    for (int k = 0; k < nz; ++k) tot += part[k];
    free(part);
    return tot;
}

static void copy_plane(const orc_diff3d *s, int axis, const double *src, int sp, double *dst, int dp)
{
    const int nx = s->nx, ny = s->ny, nz = s->nz;
    if (axis == 0) {
        for (int k = 0; k < nz; ++k)
            for (int j = 0; j < ny; ++j)
                dst[dp + (size_t)nx * (j + (size_t)ny * k)] = src[sp + (size_t)nx * (j + (size_t)ny * k)];
    } else if (axis == 1) {
        for (int k = 0; k < nz; ++k)
            memcpy(dst + (size_t)nx * (dp + (size_t)ny * k), src + (size_t)nx * (sp + (size_t)ny * k),
                   nx * sizeof(double));
    } else {
        memcpy(dst + (size_t)nx * ny * dp, src + (size_t)nx * ny * sp, (size_t)nx * ny * sizeof(double));
    }
}

/* ImplicitGlobalGrid update_halo!: per axis x,y,z; low rank's plane n-2 -> high rank's plane 0,
 * high rank's plane 1 -> low rank's plane n-1; whole planes including edges. */
static void update_halo(const orc_diff3d *s, double **F)
{
    const int nd[3] = {s->nx, s->ny, s->nz};
    for (int axis = 0; axis < 3; ++axis) {
        if (s->dims[axis] == 1) continue;
        for (int r = 0; r < s->nranks; ++r) {
            int c[3]; rank_coords(s, r, c);
            if (c[axis] + 1 >= s->dims[axis]) continue;
            int ch[3] = {c[0], c[1], c[2]}; ch[axis] += 1;
            int rh = coords_rank(s, ch);
            int n = nd[axis];
            copy_plane(s, axis, F[r], n - 2, F[rh], 0);
            copy_plane(s, axis, F[rh], 1, F[r], n - 1);
        }
    }
}

/* One PT iteration on all emulated ranks; returns err (part1_kernel_programming.jl:181-191). */
double orc_diff3d_iterate_once(orc_diff3d *s)
{
    double sq = 0.0;
    if (s->array) { /* part1_array_programming.jl:65-68: step in place, update_halo!(Htau), norm */
        for (int r = 0; r < s->nranks; ++r) array_step_rank(s, s->Ht[r], s->A[r], s->R[r]);
        update_halo(s, s->A);
    } else {
        for (int r = 0; r < s->nranks; ++r) step_tau_rank(s, s->Ht[r], s->A[r], s->B[r], s->R[r]);
        update_halo(s, s->halo_mode == ORC_HALO_REFERENCE_LAG2 ? s->A : s->B);
        double **t = s->A; s->A = s->B; s->B = t;
    }
    for (int r = 0; r < s->nranks; ++r) sq += sumsq_rank(s, s->R[r]); /* MPI.Allreduce!(+) in rank order */
    s->iters_total += 1;
    return sqrt(sq) / sqrt(s->total_N);
}

/* Fixed number of iterations; err_hist (nullable) receives n values. */
void orc_diff3d_iterate(orc_diff3d *s, int n, double *err_hist)
{
    for (int i = 0; i < n; ++i) {
        double e = orc_diff3d_iterate_once(s);
        if (err_hist) err_hist[i] = e;
    }
}

/* while err > tol && iter < iter_max (part1_kernel_programming.jl:177-193). Returns iterations. */
int orc_diff3d_solve_timestep(orc_diff3d *s, double tol, int iter_max, double *err_out)
{
    int it = 0;
    double err = 2 * tol;
    while (err > tol && it < iter_max) { err = orc_diff3d_iterate_once(s); ++it; }
    if (err_out) *err_out = err;
    return it;
}

/* Ht .= Htau (part1_kernel_programming.jl:203) */
void orc_diff3d_advance_time(orc_diff3d *s)
{
    for (int r = 0; r < s->nranks; ++r) memcpy(s->Ht[r], s->A[r], cells(s) * sizeof(double));
}

/* which: 0 Ht, 1 Htau (current), 2 Htau2 (other buffer), 3 residual_H */
void orc_diff3d_get(const orc_diff3d *s, int rank, int which, double *out)
{
    const double *src = which == 0 ? s->Ht[rank] : which == 1 ? s->A[rank] : which == 2 ? s->B[rank] : s->R[rank];
    memcpy(out, src, cells(s) * sizeof(double));
}

/* gather!(Array(Ht), H_g): H_g is (nx*dimx, ny*dimy, nz*dimz), each rank's whole local array in its block
 * (part1_kernel_programming.jl:144,223). */
void orc_diff3d_gather(const orc_diff3d *s, double *H_g)
{
    const int nx = s->nx, ny = s->ny, nz = s->nz;
    const size_t gx = (size_t)nx * s->dims[0], gy = (size_t)ny * s->dims[1];
    for (int r = 0; r < s->nranks; ++r) {
        int c[3]; rank_coords(s, r, c);
        for (int k = 0; k < nz; ++k)
            for (int j = 0; j < ny; ++j)
                memcpy(H_g + (size_t)c[0] * nx + gx * ((size_t)c[1] * ny + j + gy * ((size_t)c[2] * nz + k)),
                       s->Ht[r] + (size_t)nx * (j + (size_t)ny * k), nx * sizeof(double));
    }
}

void orc_diff3d_params(const orc_diff3d *s, double *out8)
{
    out8[0] = s->dx; out8[1] = s->dy; out8[2] = s->dz; out8[3] = s->dt;
    out8[4] = s->dtau; out8[5] = s->lx; out8[6] = s->ly; out8[7] = s->lz;
}

/* Number of elements of the Julia range 0:dt:ttot-dt (part1_kernel_programming.jl:166). */
int orc_diff3d_num_timesteps(double ttot, double dt)
{
    double stop = ttot - dt;
    if (stop < 0) return 0;
    return (int)floor(stop / dt + 1e-9) + 1;
}

/* Whole run: returns total iterations; iters_per_step must hold num_timesteps entries. */
long orc_diff3d_run(orc_diff3d *s, double ttot, double tol, int iter_max, int *iters_per_step)
{
    int nt = orc_diff3d_num_timesteps(ttot, s->dt);
    long tot = 0;
    for (int t = 0; t < nt; ++t) {
        int it = orc_diff3d_solve_timestep(s, tol, iter_max, NULL);
        if (iters_per_step) iters_per_step[t] = it;
        tot += it;
        orc_diff3d_advance_time(s);
    }
    return tot;
}

int orc_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
