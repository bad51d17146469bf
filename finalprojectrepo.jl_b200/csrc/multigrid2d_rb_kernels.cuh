// multigrid2d_rb_kernels.cuh -- variant B of hot path 2 (red-black Gauss-Seidel smoother + full-weighting restriction),
// temporally blocked: one kernel per level for the downward leg and one for the upward leg, like the variant-A tile
// kernels of multigrid2d_kernels.cuh.
//
// The split red/black layout lives where the sweeps happen -- in SHARED MEMORY. The arrays in HBM keep the caller's
// natural column-major layout (a Julia CuArray can be passed as it is) and are read once and written once per leg with
// fully coalesced accesses; while a tile is staged, every element is routed to one of two per-colour planes
// (red: i+j even, black: i+j odd; element (c, r) of the staged window -> plane[(c+r)&1][r][c>>1]). A half sweep then
// reads and writes unit-stride runs of one plane and unit-stride runs of the other (no 2-way bank conflicts, which a
// natural layout would cost every half sweep). The right-hand side is split the same way; after the pre-smoothing its
// planes are overwritten in place by the residual (res[p] needs rhs[p] only), from which the full-weighting stencil is
// gathered.
//
//   mg_down_rb_kernel:  u_s = BR(BR(u)) (B = black, R = red half sweep, red first) ; rc = FW(residual(u_s)) (+ Neumann) ;
//                       ec = 0                                            [reads u, rhs; writes u_s, rc, ec]
//   mg_up_rb_kernel:    u = BR(BR(u_s - P(ec))) ; sum of the pre-update res^2 of the last red + black half sweeps
//                                                                         [reads u_s, rhs, ec; writes u]
// Semantics: iteration_2DPoisson_gs! (scripts-part2/multigrid.jl:269-297) restricted to one colour per half sweep
// (alpha = 1, residual written with "/ h^2" as there), residual_2DPoisson! (:173-188), the V-cycle order of
// Vcycle_2DPoisson! (:121-143). Every point value is produced by exactly the arithmetic of the unfused kernels
// (mg_rbgs_kernel, mg_restrict_kernel, mg_prolong_kernel); halo points are recomputed redundantly by neighbouring
// blocks (the halo shrinks by one per half sweep), so results are bit-identical to the unfused path and to the oracle.
//
// Tile shapes: 52x32 / 52x16 (staged width 64) and 20x16 / 20x8 (staged width 32): one window row holds at most 32 points
// of a colour, so a half sweep is done row-wise by lane groups (no per-point index arithmetic); wider tiles fall back to
// a flat loop. The per-level constants (C, h^2, its exact reciprocal when h^2 is a power of two, the Gauss-Seidel weight)
// are computed on the host and read from the call block (LevelCoef).
#pragma once
#include "multigrid2d_kernels.cuh"

namespace b2s {

template <int TW, int TH, int HALO>
struct RbCfg {
    static constexpr int kW = TW + 2 * HALO;     // staged columns (even)
    static constexpr int kRows = TH + 2 * HALO;  // staged rows
    static constexpr int kP2 = kW / 2;           // row pitch of one colour plane
    static constexpr int kPlane = kRows * kP2 + 4;  // +4: planes start 32 bytes apart in bank space
    static constexpr int kCW = TW / 2 + 5, kCH = TH / 2 + 5;  // coarse window of the upward kernel (HALO = 4)
    static constexpr size_t kSmemDown = (size_t)4 * kPlane * sizeof(double);
    static constexpr size_t kSmemUp = ((size_t)4 * kPlane + (size_t)kCW * kCH) * sizeof(double);
};

struct RbCoef {
    double C, h2, w, inv_h2;  // inv_h2: exact reciprocal when h^2 is a power of two (DivH2), unused otherwise
};
__device__ __forceinline__ RbCoef level_rb_coef(const MGCall *cp, int level)
{
    const LevelCoef *L = level_consts(cp, level);
    RbCoef k;
    k.C = L->C; k.h2 = L->h2; k.w = L->wGS; k.inv_h2 = L->inv_h2;
    return k;
}

// Stage a window of a natural-layout global array into the two colour planes (asynchronous 8-byte copies; elements
// outside the domain are zero-filled). (gx0, gy0) = global coordinates of window element (0,0); gx0 + gy0 is even, so
// the local parity (c + r) & 1 is the global colour. Elements outside [c0, c1) x [r0, r1) are skipped.
template <int kW, int kRows, int kP2>
__device__ __forceinline__ void rb_stage(double *__restrict__ red, double *__restrict__ blk, const double *__restrict__ g, int gx0,
                                         int gy0, int nx, int ny, int c0, int c1, int r0, int r1)
{
    for (int r = r0 + (threadIdx.x >> 5); r < r1; r += kTileThreads / 32) {  // one warp per staged row, lanes along x
        const int j = gy0 + r;
        const bool jin = j >= 0 && j < ny;
        const size_t rowoff = jin ? (size_t)nx * j : 0;
        double *rowR = red + r * kP2, *rowB = blk + r * kP2;
#pragma unroll
        for (int c = c0 + (threadIdx.x & 31); c < c1; c += 32) {
            const int i = gx0 + c;
            const bool in = jin && i >= 0 && i < nx;
            cp_async8((((c + r) & 1) ? rowB : rowR) + (c >> 1), g + (in ? rowoff + i : 0), in);
        }
    }
}

// One half sweep of colour X (0 = red) over the window [HALO-K, HALO+TW+K) x [HALO-K, HALO+TH+K) of the staged tile,
// in place in the colour planes. NORM: accumulate the pre-update res^2 of the points inside the tile proper.
// DIV: true = divide by h^2 as written in the reference, false = multiply by the exact reciprocal (same bits, DivH2).
template <int TW, int TH, int HALO, int K, bool CHECKED, bool NORM, bool DIV>
__device__ __forceinline__ double rb_half_sweep(double *__restrict__ own, const double *__restrict__ oth,
                                                const double *__restrict__ F, int X, int gx0, int gy0, int nx, int ny,
                                                const RbCoef &k)
{
    using Cf = RbCfg<TW, TH, HALO>;
    constexpr int kP2 = Cf::kP2;
    constexpr int w0 = HALO - K, W = TW + 2 * K, H = TH + 2 * K;
    constexpr int HW = (W + 1) / 2;  // points of one colour per window row (W is even)
    double acc = 0.0;
    auto point = [&](int r, int c, int par) {
        if (CHECKED) {
            const int i = gx0 + c, j = gy0 + r;
            if (i < 1 || j < 1 || i > nx - 2 || j > ny - 2) return;
        }
        const int s = r * kP2 + (c >> 1);
        const double v = own[s];
        // (uE + uW + uN + uS - C u) / h^2 - f     multigrid.jl:279-283; E/W neighbours sit at s+par / s-1+par of the
        // other plane, N/S neighbours at the same index of the rows above and below
        const double t = oth[s + par] + oth[s - 1 + par] + oth[s + kP2] + oth[s - kP2] - k.C * v;
        const double res = (DIV ? t / k.h2 : t * k.inv_h2) - F[s];
        own[s] = v + k.w * res;
        if (NORM) {
            if (K == 0 || (c >= HALO && c < HALO + TW && r >= HALO && r < HALO + TH)) acc += res * res;
        }
    };
    if constexpr (HW <= 32) {
        // row-wise: a group of LPR lanes owns a window row (no per-point division; parity and row offsets are per row)
        constexpr int LPR = HW <= 8 ? 8 : (HW <= 16 ? 16 : 32), RPW = 32 / LPR;
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        const int sub = lane / LPR, kk = lane % LPR;
        if (kk < HW) {
            for (int rr = warp * RPW + sub; rr < H; rr += (kTileThreads / 32) * RPW) {
                const int r = w0 + rr;
                const int par = (r + X) & 1;                   // parity of the columns of colour X in this row
                point(r, w0 + ((w0 ^ par) & 1) + 2 * kk, par); // kk-th column >= w0 of that parity
            }
        }
    } else {
        for (int idx = threadIdx.x; idx < H * HW; idx += kTileThreads) {
            const int rr = idx / HW, kk = idx - rr * HW;
            const int r = w0 + rr;
            const int par = (r + X) & 1;
            point(r, w0 + ((w0 ^ par) & 1) + 2 * kk, par);
        }
    }
    return acc;
}

// tile proper -> natural-layout global array, one warp per row (coalesced; the two planes are read alternately)
template <int TW, int TH, int HALO>
__device__ __forceinline__ void rb_store_tile(const double *__restrict__ UR, const double *__restrict__ UB, double *__restrict__ out,
                                              int X0, int Y0, int nx, int ny)
{
    constexpr int kP2 = RbCfg<TW, TH, HALO>::kP2;
    for (int r = threadIdx.x >> 5; r < TH; r += kTileThreads / 32) {
        const int j = Y0 + r;
        if (j >= ny) continue;
        const int lr = r + HALO;
        double *orow = out + (size_t)nx * j + X0;
        for (int c = threadIdx.x & 31; c < TW; c += 32) {
            if (X0 + c >= nx) continue;
            const int lc = c + HALO;
            orow[c] = (((lc + lr) & 1) ? UB : UR)[lr * kP2 + (lc >> 1)];
        }
    }
}

// the four half sweeps of the downward leg (pre-smoothing: red, black, red, black; the valid window shrinks by one each)
template <int TW, int TH, bool CHECKED, bool DIV>
__device__ __forceinline__ void rb_down_sweeps(double *UR, double *UB, const double *FR, const double *FB, int gx0, int gy0, int nx,
                                               int ny, const RbCoef &k)
{
    rb_half_sweep<TW, TH, 6, 5, CHECKED, false, DIV>(UR, UB, FR, 0, gx0, gy0, nx, ny, k);
    __syncthreads();
    rb_half_sweep<TW, TH, 6, 4, CHECKED, false, DIV>(UB, UR, FB, 1, gx0, gy0, nx, ny, k);
    __syncthreads();
    rb_half_sweep<TW, TH, 6, 3, CHECKED, false, DIV>(UR, UB, FR, 0, gx0, gy0, nx, ny, k);
    __syncthreads();
    rb_half_sweep<TW, TH, 6, 2, CHECKED, false, DIV>(UB, UR, FB, 1, gx0, gy0, nx, ny, k);
}
// the four half sweeps of the upward leg; returns this thread's share of the last full sweep's pre-update res^2
template <int TW, int TH, bool CHECKED, bool DIV>
__device__ __forceinline__ double rb_up_sweeps(double *UR, double *UB, const double *FR, const double *FB, int gx0, int gy0, int nx,
                                               int ny, const RbCoef &k)
{
    rb_half_sweep<TW, TH, 4, 3, CHECKED, false, DIV>(UR, UB, FR, 0, gx0, gy0, nx, ny, k);
    __syncthreads();
    rb_half_sweep<TW, TH, 4, 2, CHECKED, false, DIV>(UB, UR, FB, 1, gx0, gy0, nx, ny, k);
    __syncthreads();
    double acc = rb_half_sweep<TW, TH, 4, 1, CHECKED, true, DIV>(UR, UB, FR, 0, gx0, gy0, nx, ny, k);
    __syncthreads();
    acc += rb_half_sweep<TW, TH, 4, 0, CHECKED, true, DIV>(UB, UR, FB, 1, gx0, gy0, nx, ny, k);
    return acc;
}

template <int TW, int TH>
__global__ void __launch_bounds__(kTileThreads) mg_down_rb_kernel(const TileArgs a)
{
    constexpr int HALO = 6;
    using Cf = RbCfg<TW, TH, HALO>;
    constexpr int kW = Cf::kW, kRows = Cf::kRows, kP2 = Cf::kP2, kPlane = Cf::kPlane;
    extern __shared__ __align__(16) double tsm[];
    double *UR = tsm, *UB = tsm + kPlane, *FR = tsm + 2 * kPlane, *FB = tsm + 3 * kPlane;
    const MGCall *cp = a.cp;
    // below the finest level the staging loads are issued before the call block has arrived (see mg_down_kernel)
    const double *u = a.u_in, *rhs = a.rhs;
    if (a.level == 0) {
        if (cp->done) return;
        u = cp->u; rhs = cp->rhs;
    }
    const int nx = a.nx, ny = a.ny;
    const int X0 = blockIdx.x * TW, Y0 = blockIdx.y * TH;
    const int gx0 = X0 - HALO, gy0 = Y0 - HALO;
    rb_stage<kW, kRows, kP2>(UR, UB, u, gx0, gy0, nx, ny, 0, kW, 0, kRows);
    rb_stage<kW, kRows, kP2>(FR, FB, rhs, gx0, gy0, nx, ny, 1, kW - 1, 1, kRows - 1);
    const int done = a.level == 0 ? 0 : cp->done;
    const int apply_bcs = cp->apply_bcs;
    const RbCoef k = level_rb_coef(cp, a.level);
    const Coef kr = level_coef(cp, a.level);  // C and 1/h^2 of the residual (the weight is not used)
    cp_async_wait_all();
    __syncthreads();
    if (done) return;
    // block-uniform: does the widest window (halo 5) stay strictly inside the domain?
    const bool inner = X0 - 5 >= 1 && Y0 - 5 >= 1 && X0 + TW + 4 <= nx - 2 && Y0 + TH + 4 <= ny - 2;
    const bool div = !level_consts(cp, a.level)->exact;
    if (inner && !div) rb_down_sweeps<TW, TH, false, false>(UR, UB, FR, FB, gx0, gy0, nx, ny, k);
    else if (inner) rb_down_sweeps<TW, TH, false, true>(UR, UB, FR, FB, gx0, gy0, nx, ny, k);
    else if (!div) rb_down_sweeps<TW, TH, true, false>(UR, UB, FR, FB, gx0, gy0, nx, ny, k);
    else rb_down_sweeps<TW, TH, true, true>(UR, UB, FR, FB, gx0, gy0, nx, ny, k);
    __syncthreads();
    // smoothed u out (natural layout, coalesced): one warp per tile row
    rb_store_tile<TW, TH, HALO>(UR, UB, a.u_out, X0, Y0, nx, ny);
    // residual of the smoothed u on tile+1, in place of the right-hand side (multigrid.jl:173-188); interior points only
    // -- the full-weighting stencil of an interior coarse point never reaches the fine frame
    {
        constexpr int W = TW + 2, H = TH + 2, w0 = HALO - 1, HW = W / 2;
        const int lane = threadIdx.x & 31;
        for (int rr = threadIdx.x >> 5; rr < H; rr += kTileThreads / 32) {
            const int r = w0 + rr, j = gy0 + r;
            if (j < 1 || j > ny - 2) continue;
#pragma unroll
            for (int X = 0; X < 2; ++X) {
                const int par = (r + X) & 1;  // column parity of colour X in this row
                const double *own = X ? UB : UR, *oth = X ? UR : UB;
                double *F = X ? FB : FR;
                for (int kk = lane; kk < HW; kk += 32) {
                    const int c = w0 + ((w0 ^ par) & 1) + 2 * kk, i = gx0 + c;
                    if (i < 1 || i > nx - 2) continue;
                    const int s = r * kP2 + (c >> 1);
                    F[s] = ((oth[s + par] + oth[s - 1 + par] + oth[s + kP2] + oth[s - kP2] - kr.C * own[s]) * kr._h2 - F[s]);
                }
            }
        }
    }
    __syncthreads();
    // coarse rhs = full weighting of the residual (+ Neumann copies), coarse unknown = 0: one warp per coarse row
    const int nxc = a.nxc, nyc = a.nyc;
    for (int rr = threadIdx.x >> 5; rr < TH / 2; rr += kTileThreads / 32) {
        const int J = Y0 / 2 + rr;
        if (J >= nyc) continue;
        const bool jint = J >= 1 && J <= nyc - 2;
        for (int cc = threadIdx.x & 31; cc < TW / 2; cc += 32) {
            const int I = X0 / 2 + cc;
            if (I >= nxc) continue;
            const size_t pc = (size_t)I + (size_t)nxc * J;
            a.ec[pc] = 0.0;
            if (jint && I >= 1 && I <= nxc - 2) {
                // fine point (2I, 2J): local (even, even) -> red plane; its x neighbours are black at m-1, m; the diagonal
                // neighbours are red at m-1, m of the rows below and above
                const int s = (2 * rr + HALO) * kP2 + (cc + HALO / 2);
                const double corners = (FR[s - kP2 - 1] + FR[s - kP2]) + (FR[s + kP2 - 1] + FR[s + kP2]);
                const double edges = (FB[s - 1] + FB[s]) + (FB[s - kP2] + FB[s + kP2]);
                const double v = ((corners + 2.0 * edges) + 4.0 * FR[s]) * 0.0625;
                a.rc[pc] = v;
                if (apply_bcs) {  // coarse[0,:] = coarse[1,:] ; coarse[nxc-1,:] = coarse[nxc-2,:]
                    if (I == 1) a.rc[(size_t)0 + (size_t)nxc * J] = v;
                    if (I == nxc - 2) a.rc[(size_t)(nxc - 1) + (size_t)nxc * J] = v;
                }
            } else if (!(apply_bcs && (I == 0 || I == nxc - 1) && jint)) {
                a.rc[pc] = 0.0;
            }
        }
    }
}

template <int TW, int TH>
__global__ void __launch_bounds__(kTileThreads) mg_up_rb_kernel(const TileArgs a)
{
    constexpr int HALO = 4;
    using Cf = RbCfg<TW, TH, HALO>;
    constexpr int kW = Cf::kW, kRows = Cf::kRows, kP2 = Cf::kP2, kPlane = Cf::kPlane, kCW = Cf::kCW, kCH = Cf::kCH;
    extern __shared__ __align__(16) double tsm[];
    __shared__ double red[32];
    double *UR = tsm, *UB = tsm + kPlane, *FR = tsm + 2 * kPlane, *FB = tsm + 3 * kPlane, *Cw = tsm + 4 * kPlane;
    const MGCall *cp = a.cp;
    const double *rhs = a.rhs;
    double *out = a.u_out;
    if (a.level == 0) {
        if (cp->done) return;
        rhs = cp->rhs; out = cp->u;
    }
    const int nx = a.nx, ny = a.ny, nxc = a.nxc, nyc = a.nyc;
    const int X0 = blockIdx.x * TW, Y0 = blockIdx.y * TH;
    const int gx0 = X0 - HALO, gy0 = Y0 - HALO;
    const int cx0 = X0 / 2 - 2, cy0 = Y0 / 2 - 2;
    rb_stage<kW, kRows, kP2>(UR, UB, a.u_in, gx0, gy0, nx, ny, 0, kW, 0, kRows);
    rb_stage<kW, kRows, kP2>(FR, FB, rhs, gx0, gy0, nx, ny, 1, kW - 1, 1, kRows - 1);
    for (int r = threadIdx.x >> 5; r < kCH; r += kTileThreads / 32) {
        const int J = cy0 + r;
        const bool jin = J >= 1 && J <= nyc - 2;  // the boundary ring counts as 0
        const size_t rowoff = jin ? (size_t)nxc * J : 0;
        for (int c = threadIdx.x & 31; c < kCW; c += 32) {
            const int I = cx0 + c;
            const bool in = jin && I >= 1 && I <= nxc - 2;
            cp_async8(Cw + r * kCW + c, a.ec + (in ? rowoff + I : 0), in);
        }
    }
    const int done = a.level == 0 ? 0 : cp->done;
    const int apply_bcs = cp->apply_bcs;
    const RbCoef k = level_rb_coef(cp, a.level);
    cp_async_wait_all();
    __syncthreads();
    if (done) return;
    // u_f .= u_f - corr_f on tile+4 (multigrid.jl:136-139): one warp per staged row, one column parity per pass (the
    // interpolation formula and the colour plane are then uniform across the warp)
    for (int r = threadIdx.x >> 5; r < kRows; r += kTileThreads / 32) {
        const int j = gy0 + r;
        if (j < 0 || j >= ny) continue;
#pragma unroll
        for (int cpar = 0; cpar < 2; ++cpar) {
            double *row = (((cpar + r) & 1) ? UB : UR) + r * kP2;
            for (int m = threadIdx.x & 31; 2 * m + cpar < kW; m += 32) {
                const int i = gx0 + 2 * m + cpar;
                if (i < 0 || i >= nx) continue;
                row[m] = row[m] - prolong_from_window<kCW>(Cw, cx0, cy0, nx, i, j, apply_bcs);
            }
        }
    }
    __syncthreads();
    const bool inner = X0 - 3 >= 1 && Y0 - 3 >= 1 && X0 + TW + 2 <= nx - 2 && Y0 + TH + 2 <= ny - 2;
    const bool div = !level_consts(cp, a.level)->exact;
    double acc;
    if (inner && !div) acc = rb_up_sweeps<TW, TH, false, false>(UR, UB, FR, FB, gx0, gy0, nx, ny, k);
    else if (inner) acc = rb_up_sweeps<TW, TH, false, true>(UR, UB, FR, FB, gx0, gy0, nx, ny, k);
    else if (!div) acc = rb_up_sweeps<TW, TH, true, false>(UR, UB, FR, FB, gx0, gy0, nx, ny, k);
    else acc = rb_up_sweeps<TW, TH, true, true>(UR, UB, FR, FB, gx0, gy0, nx, ny, k);
    __syncthreads();
    rb_store_tile<TW, TH, HALO>(UR, UB, out, X0, Y0, nx, ny);
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

}  // namespace b2s
