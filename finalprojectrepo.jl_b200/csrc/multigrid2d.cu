// multigrid2d.cu -- host side of hot path 2: L0 wrappers (1:1 with the reference's call sites) and the L1 multigrid
// handle (level table allocated once, fine levels as fused global-memory kernels, all coarse levels collapsed into one
// shared-memory kernel, the whole V-cycle captured in a CUDA graph whose per-call arguments live in device memory).
//
// Reference entry points mirrored: MGsolve_2DPoisson! / Vcycle_2DPoisson! (scripts-part2/multigrid.jl:41-170),
// cg! (scripts-part2/krylov.jl:55-91).
#include "multigrid2d_kernels.cuh"
#include "multigrid2d_rb_kernels.cuh"

#include <math.h>
#include <stdlib.h>

#include <algorithm>
#include <vector>

using namespace b2s;

namespace {

inline int rows_for(int nx, int ny)
{
    // ~16 resident blocks of 128 threads per SM (148 SMs) so the loads in flight cover the HBM latency, >= 4 rows/thread
    const int bx = (nx + kMGBX - 1) / kMGBX;
    int want = std::max(1, (148 * 16) / bx);
    int rows = std::max(4, (ny + want - 1) / want);
    return std::min(rows, std::max(ny, 1));
}

inline dim3 sweep_grid(int nx, int ny, int rows) { return dim3((nx + kMGBX - 1) / kMGBX, (ny + rows - 1) / rows, 1); }

inline bool pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

constexpr size_t kCoarseSmemLimit = 200 * 1024;
// device / pinned layout of a call: [MGCall][LevelCoef x kMaxLevels]; the pinned copy is followed by the read-back MGCall
constexpr size_t kCallBlockBytes = sizeof(MGCall) + kMaxLevels * sizeof(LevelCoef);
inline size_t smem_level_max_points()
{
    const char *e = getenv("B2S_MG_SMEM_MAXPTS");
    return (e && *e) ? (size_t)atoll(e) : (size_t)600;  // measured: 33^2 and 65^2 are faster as small-tile kernels
}

}  // namespace

struct b2s_mg {
    b2s_mg_config cfg;
    int nlev = 0;  // total levels; level nlev-1 is the coarsest
    int nx[kMaxLevels], ny[kMaxLevels];
    double *u[kMaxLevels] = {}, *rhs[kMaxLevels] = {}, *tmp[kMaxLevels] = {};  // u/rhs for l >= 1, tmp for all
    int first_smem = 0;  // first level handled by the collapsed kernel
    size_t coarse_smem = 0;
    // thread-block-cluster kernel for the latency-bound middle levels (variant A, fused): levels mid_first .. first_smem-1
    // live in the distributed shared memory of one cluster of mid_nc blocks, first_smem.. in block 0 (mg_mid_cluster_kernel)
    int mid_first = 0;   // == first_smem: disabled
    int mid_nc = 0, mid_base = 0;
    size_t mid_smem = 0;
    MGCall *call_dev = nullptr, *call_pin = nullptr;  // call_pin: upload staging [MGCall][LevelCoef...], then the read-back MGCall
    double *hist_dev = nullptr, *hist_pin = nullptr;
    double *sumsq_dev = nullptr;  // [0] last sweep, [1] f, [2],[3] rbgs colours
    double *sumsq_pin = nullptr;
    int *sweeps_dev = nullptr;
    double *partials = nullptr;
    unsigned int *ticket = nullptr;
    cudaStream_t stream = nullptr;
    cudaGraphExec_t graph[2] = {nullptr, nullptr};  // [bc_before]
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    long long kernel_launches = 0;
    long long launches_per_cycle[2] = {0, 0};
    double last_ms = 0.0;
    int tile_choice = -1;
    int tile_min_blocks = 400;
    int rb_tile_choice = -1;
    int stream_ch = 0;
    bool stream_warp = false;            // automatic mode: block-wide two-column streaming kernels (measured faster than the
                                         // one-warp-per-strip variant; B2S_MG_STREAM_KIND=warp selects the latter)
    size_t stream_min_points = 1500000;  // levels above this use the streaming kernels (B2S_MG_STREAM_MIN)
    long long *prof_dev = nullptr;  // B2S_MG_PROF=1: phase stamps of the collapsed coarse kernel
    bool coarse_global = false;     // coarsest level too large for shared memory: solved by global-memory kernels
    CoarseLoop *loop_dev = nullptr, *loop_pin = nullptr;
    double *cg_work = nullptr;      // 4 arrays of the coarsest size (global CG)
    double *pcg_work = nullptr;     // 4 arrays of the finest size (MG-preconditioned CG), allocated on first use
    int last_sweeps_host = -1;
    int last_ncycles[2][2] = {{0, 0}, {0, 0}};  // [apply_BCs][c != 0]: V-cycles the last solve of that kind needed
};

namespace {

// tile shapes of the temporally blocked kernels (B2S_MG_TILE selects; 0 is the tuned default)
template <int TW, int TH>
void launch_tile_t(bool up, const TileArgs &t, cudaStream_t st)
{
    dim3 g((t.nx + TW - 1) / TW, (t.ny + TH - 1) / TH, 1);
    if (up) mg_up_kernel<TW, TH><<<g, kTileThreads, TileCfg<TW, TH>::kSmemBytes, st>>>(t);
    else mg_down_kernel<TW, TH><<<g, kTileThreads, TileCfg<TW, TH>::kSmemBytes, st>>>(t);
}
void launch_tile(int choice, bool up, const TileArgs &t, cudaStream_t st)
{
    switch (choice) {
    default:
    case 0: launch_tile_t<64, 16>(up, t, st); break;
    case 1: launch_tile_t<32, 32>(up, t, st); break;
    case 2: launch_tile_t<64, 32>(up, t, st); break;
    case 3: launch_tile_t<128, 16>(up, t, st); break;
    case 4: launch_tile_t<32, 16>(up, t, st); break;
    case 5: launch_tile_t<32, 8>(up, t, st); break;
    case 6: launch_tile_t<16, 8>(up, t, st); break;
    }
}
constexpr int kTileChoices = 7;
constexpr int kTileW[kTileChoices] = {64, 32, 64, 128, 32, 32, 16}, kTileH[kTileChoices] = {16, 32, 32, 16, 16, 8, 8};
template <int TW, int TH>
cudaError_t tile_set_attr_t()
{
    cudaError_t e = cudaFuncSetAttribute(mg_down_kernel<TW, TH>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)TileCfg<TW, TH>::kSmemBytes);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(mg_up_kernel<TW, TH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TileCfg<TW, TH>::kSmemBytes);
}
cudaError_t tile_set_attr(int choice)
{
    switch (choice) {
    default:
    case 0: return tile_set_attr_t<64, 16>();
    case 1: return tile_set_attr_t<32, 32>();
    case 2: return tile_set_attr_t<64, 32>();
    case 3: return tile_set_attr_t<128, 16>();
    case 4: return tile_set_attr_t<32, 16>();
    case 5: return tile_set_attr_t<32, 8>();
    case 6: return tile_set_attr_t<16, 8>();
    }
}

// variant B (red-black Gauss-Seidel + full weighting) tile kernels; B2S_MG_RB_TILE selects the shape
template <int TW, int TH>
void launch_rb_tile_t(bool up, const TileArgs &t, cudaStream_t st)
{
    dim3 g((t.nx + TW - 1) / TW, (t.ny + TH - 1) / TH, 1);
    if (up) mg_up_rb_kernel<TW, TH><<<g, kTileThreads, RbCfg<TW, TH, 4>::kSmemUp, st>>>(t);
    else mg_down_rb_kernel<TW, TH><<<g, kTileThreads, RbCfg<TW, TH, 6>::kSmemDown, st>>>(t);
}
void launch_rb_tile(int choice, bool up, const TileArgs &t, cudaStream_t st)
{
    switch (choice) {
    default:
    case 0: launch_rb_tile_t<64, 32>(up, t, st); break;
    case 1: launch_rb_tile_t<64, 16>(up, t, st); break;
    case 2: launch_rb_tile_t<32, 32>(up, t, st); break;
    case 3: launch_rb_tile_t<128, 16>(up, t, st); break;
    case 4: launch_rb_tile_t<128, 32>(up, t, st); break;
    case 5: launch_rb_tile_t<32, 16>(up, t, st); break;
    case 6: launch_rb_tile_t<16, 16>(up, t, st); break;
    case 7: launch_rb_tile_t<16, 8>(up, t, st); break;
    case 8: launch_rb_tile_t<52, 32>(up, t, st); break;   // staged width 64: the half sweeps run row-wise
    case 9: launch_rb_tile_t<52, 16>(up, t, st); break;
    case 10: launch_rb_tile_t<20, 16>(up, t, st); break;  // staged width 32
    case 11: launch_rb_tile_t<20, 8>(up, t, st); break;
    }
}
constexpr int kRbTileChoices = 12;
constexpr int kRbTileW[kRbTileChoices] = {64, 64, 32, 128, 128, 32, 16, 16, 52, 52, 20, 20},
              kRbTileH[kRbTileChoices] = {32, 16, 32, 16, 32, 16, 16, 8, 32, 16, 16, 8};
template <int TW, int TH>
cudaError_t rb_tile_set_attr_t()
{
    cudaError_t e = cudaFuncSetAttribute(mg_down_rb_kernel<TW, TH>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)RbCfg<TW, TH, 6>::kSmemDown);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(mg_up_rb_kernel<TW, TH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RbCfg<TW, TH, 4>::kSmemUp);
}
cudaError_t rb_tile_set_attr(int choice)
{
    switch (choice) {
    default:
    case 0: return rb_tile_set_attr_t<64, 32>();
    case 1: return rb_tile_set_attr_t<64, 16>();
    case 2: return rb_tile_set_attr_t<32, 32>();
    case 3: return rb_tile_set_attr_t<128, 16>();
    case 4: return rb_tile_set_attr_t<128, 32>();
    case 5: return rb_tile_set_attr_t<32, 16>();
    case 6: return rb_tile_set_attr_t<16, 16>();
    case 7: return rb_tile_set_attr_t<16, 8>();
    case 8: return rb_tile_set_attr_t<52, 32>();
    case 9: return rb_tile_set_attr_t<52, 16>();
    case 10: return rb_tile_set_attr_t<20, 16>();
    case 11: return rb_tile_set_attr_t<20, 8>();
    }
}

// fused level kernels exist for Jacobi + injection (variant A) and red-black Gauss-Seidel + full weighting (variant B)
inline bool fused_variant_a(const b2s_mg_config &c)
{
    return c.fuse_sweeps && c.smoother == B2S_SMOOTH_JACOBI && c.restriction == B2S_RESTRICT_INJECT;
}
inline bool fused_variant_b(const b2s_mg_config &c)
{
    return c.fuse_sweeps && c.smoother == B2S_SMOOTH_RBGS && c.restriction == B2S_RESTRICT_FW;
}

int coarse_global_solve(b2s_mg *h, cudaStream_t st, long long *count);
int cg_global(double *x_in, const double *b, double *work, double hx, double hy, double c, double tol, int nmax, int nx, int ny,
              cudaStream_t st, double *ss_out, int *iters_out, long long *count);

// Enqueues one V-cycle (multigrid.jl:91-170) on `st`; counts kernel launches. only_level >= 0 (profiling, fused paths):
// emit just one kernel of the cycle -- the downward (only_dir 0) or upward (1) kernel of that level, or the kernel that
// handles everything below the global-memory levels (2: collapsed coarse kernel / cluster kernel).
int enqueue_vcycle(b2s_mg *h, cudaStream_t st, long long *count, bool with_bc, int only_level = -1, int only_dir = 0)
{
    auto skip = [&](int l, int dir) { return only_level >= 0 && !(dir == only_dir && (dir == 2 || l == only_level)); };
    const b2s_mg_config &c = h->cfg;
    const MGCall *cp = h->call_dev;
    long long n = 0;
    const bool rb = c.smoother == B2S_SMOOTH_RBGS;
    // apply_boundary_conditions!(u) before the cycle (multigrid.jl:60-62); calls without it use a graph without the node
    if (with_bc && only_level < 0) {
        const int t = h->nx[0] + h->ny[0];
        mg_bc_kernel<<<(t + 255) / 256, 256, 0, st>>>(cp, nullptr, h->nx[0], h->ny[0], 0);
        ++n;
    }
    auto smooth2 = [&](int l, bool norm_on_last) {
        const int nx = h->nx[l], ny = h->ny[l];
        const int rows = rows_for(nx, ny);
        if (rb) {
            for (int s = 0; s < 2; ++s)
                for (int colour = 0; colour < 2; ++colour) {
                    RbgsArgs a = {};
                    a.cp = cp; a.level = l; a.u = h->u[l]; a.rhs = h->rhs[l]; a.nx = nx; a.ny = ny; a.rows = rows;
                    a.colour = colour; a.want_norm = (norm_on_last && s == 1);
                    a.partials = h->partials; a.ticket = h->ticket; a.sumsq_out = h->sumsq_dev + 2;
                    dim3 g(((nx + 1) / 2 + kMGBX - 1) / kMGBX, (ny + rows - 1) / rows, 1);
                    mg_rbgs_kernel<<<g, kMGBX, 0, st>>>(a);
                    ++n;
                }
        } else {
            for (int s = 0; s < 2; ++s) {
                SweepArgs a = {};
                a.cp = cp; a.level = l; a.nx = nx; a.ny = ny; a.rows = rows; a.alpha = 4.0 / 5.0; a.mode = 1;
                a.rhs = h->rhs[l];
                a.u = s == 0 ? h->u[l] : h->tmp[l];
                a.out = s == 0 ? h->tmp[l] : h->u[l];
                a.swap_io = s;
                a.want_norm = (norm_on_last && s == 1);
                a.partials = h->partials; a.ticket = h->ticket; a.sumsq_out = h->sumsq_dev;
                mg_sweep_kernel<<<sweep_grid(nx, ny, rows), kMGBX, 0, st>>>(a);
                ++n;
            }
        }
    };
    const bool use_mid = !h->coarse_global && h->mid_nc > 0 && h->mid_first < h->first_smem;
    const int fs = h->coarse_global ? h->nlev - 1 : (use_mid ? h->mid_first : h->first_smem);
    const bool fused_b = fused_variant_b(c);
    const bool fused = fused_variant_a(c) || fused_b;
    // the finest level's fused upward kernel finishes the cycle itself (norm -> r_rms -> exit test) when there is one
    const bool end_fused = fused && fs > 0;
    auto tile_args = [&](int l) {
        TileArgs t = {};
        t.cp = cp; t.level = l; t.rhs = h->rhs[l]; t.nx = h->nx[l]; t.ny = h->ny[l]; t.nxc = h->nx[l + 1]; t.nyc = h->ny[l + 1];
        t.partials = h->partials; t.ticket = h->ticket; t.sumsq_out = h->sumsq_dev;
        return t;
    };
    // variant-A tile shape per level: the levels below ~0.3 M points are latency-bound chains of block-wide phases, so they
    // get the largest tile that still yields >= tile_min_blocks blocks (more, smaller blocks = shorter per-block chains;
    // the extra halo traffic stays in L2). B2S_MG_TILE pins one shape for every level.
    auto tile_choice_for = [&](int l) {
        if (h->tile_choice >= 0) return h->tile_choice;
        const int order[4] = {0, 4, 5, 6};  // 64x16, 32x16, 32x8, 16x8
        for (int k = 0; k < 4; ++k) {
            const int c_ = order[k];
            const long nb = (long)((h->nx[l] + kTileW[c_] - 1) / kTileW[c_]) * ((h->ny[l] + kTileH[c_] - 1) / kTileH[c_]);
            if (nb >= h->tile_min_blocks) return c_;
        }
        return 6;
    };
    // variant-B tile shape: 32x32 on the latency-bound (L2-resident) levels, 64x32 (less halo redundancy) above -- measured
    auto rb_choice = [&](int l) {
        if (h->rb_tile_choice >= 0) return h->rb_tile_choice;
        if ((size_t)h->nx[l] * h->ny[l] > 1500000) return 8;
        const int order[4] = {8, 9, 10, 11};  // 52x32, 52x16, 20x16, 20x8 (row-wise half sweeps)
        for (int k = 0; k < 4; ++k) {
            const int c_ = order[k];
            const long nb = (long)((h->nx[l] + kRbTileW[c_] - 1) / kRbTileW[c_]) * ((h->ny[l] + kRbTileH[c_] - 1) / kRbTileH[c_]);
            if (nb >= h->tile_min_blocks) return c_;
        }
        return 11;
    };
    // downward leg on the global-memory levels
    // fuse_sweeps: 1 = automatic (streaming kernels for large levels, where their lower instruction count wins; tile
    // kernels for levels <= ~1.5 M points, which are latency-bound and prefer 3 barriers to a 24-step pipeline),
    // 2 = tiles everywhere, 3 = streaming everywhere
    auto use_streaming = [&](int l) {
        if (c.fuse_sweeps == 2) return false;
        if (c.fuse_sweeps == 3 || c.fuse_sweeps == 4 || c.fuse_sweeps == 5) return true;
        return (size_t)h->nx[l] * h->ny[l] > h->stream_min_points;
    };
    const bool two_col = c.fuse_sweeps != 3;  // 1 (auto) and 4: two columns per thread; 3: one column per thread
    const bool warp_kind = c.fuse_sweeps == 5 || (c.fuse_sweeps == 1 && h->stream_warp);  // one warp per strip
    auto warp_rows = [&](int l) {
        const int bx = (h->nx[l] + kWW - 1) / kWW, ny = h->ny[l], slots = 148 * 28;
        int ch = 16;
        for (int w = 1; w <= 64; ++w) {
            const int chunks = std::max(1, (w * slots) / bx);
            ch = (ny + chunks - 1) / chunks;
            if (ch <= 256) break;
        }
        if (h->stream_ch > 0) ch = h->stream_ch;
        return std::max(16, (ch + 1) & ~1);
    };
    auto warp_grid = [&](int l, int ch) { return dim3((h->nx[l] + kWW - 1) / kWW, (h->ny[l] + ch - 1) / ch, 1); };
    auto stream2_rows = [&](int l) {
        const int bx = (h->nx[l] + kS2W - 1) / kS2W, ny = h->ny[l], slots = 148 * 4;
        int ch = 16;
        for (int w = 1; w <= 64; ++w) {
            const int chunks = std::max(1, (w * slots) / bx);
            ch = (ny + chunks - 1) / chunks;
            if (ch <= 256) break;
        }
        if (h->stream_ch > 0) ch = h->stream_ch;
        return std::max(16, (ch + 1) & ~1);
    };
    auto stream2_grid = [&](int l, int ch) { return dim3((h->nx[l] + kS2W - 1) / kS2W, (h->ny[l] + ch - 1) / ch, 1); };
    auto stream_rows = [&](int l) {
        // rows per chunk (even, >= 16, <= ~256): the grid should fill whole waves of 148 SMs x 6 resident blocks
        const int bx = (h->nx[l] + kSW - 1) / kSW, ny = h->ny[l], slots = 148 * B2S_STREAM_MINBLOCKS;
        int ch = 16;
        for (int w = 1; w <= 64; ++w) {
            const int chunks = std::max(1, (w * slots) / bx);
            ch = (ny + chunks - 1) / chunks;
            if (ch <= 256) break;
        }
        if (h->stream_ch > 0) ch = h->stream_ch;
        ch = std::max(16, (ch + 1) & ~1);
        return ch;
    };
    auto stream_grid = [&](int l, int ch) { return dim3((h->nx[l] + kSW - 1) / kSW, (h->ny[l] + ch - 1) / ch, 1); };
    for (int l = 0; l < fs && fused; ++l) {
        if (skip(l, 0)) continue;
        TileArgs t = tile_args(l);
        t.u_in = h->u[l]; t.u_out = h->tmp[l]; t.rc = h->rhs[l + 1]; t.ec = h->u[l + 1];
        if (fused_b) {
            launch_rb_tile(rb_choice(l), false, t, st);
        } else if (use_streaming(l) && warp_kind) {
            const int ch = warp_rows(l);
            mg_down_warp_kernel<<<warp_grid(l, ch), 32, 0, st>>>(t, ch);
        } else if (use_streaming(l) && two_col) {
            const int ch = stream2_rows(l);
            mg_down_stream2_kernel<<<stream2_grid(l, ch), kS2NT + 32, kS2SmemDown, st>>>(t, ch);
        } else if (use_streaming(l)) {
            const int ch = stream_rows(l);
            mg_down_stream_kernel<<<stream_grid(l, ch), kSNT, 0, st>>>(t, ch);
        } else {
            launch_tile(tile_choice_for(l), false, t, st);
        }
        ++n;
    }
    for (int l = 0; l < fs && !fused; ++l) {
        smooth2(l, false);
        RestrictArgs r = {};
        r.cp = cp; r.level = l; r.u = h->u[l]; r.rhs = h->rhs[l]; r.coarse = h->rhs[l + 1]; r.zero_out = h->u[l + 1];
        r.nx = h->nx[l]; r.ny = h->ny[l]; r.nxc = h->nx[l + 1]; r.nyc = h->ny[l + 1];
        r.from_res = 1; r.full_weighting = c.restriction == B2S_RESTRICT_FW;
        dim3 g((r.nxc + 63) / 64, (r.nyc + 3) / 4, 1);
        mg_restrict_kernel<<<g, dim3(64, 4, 1), 0, st>>>(r);
        ++n;
    }
    // coarsest level in global memory (too large for shared memory)
    if (skip(0, 2)) {
    } else if (h->coarse_global) {
        long long nn = 0;
        B2S_CHECK(coarse_global_solve(h, st, &nn));
        n += nn;
    } else if (use_mid) {
        // middle levels in the distributed shared memory of one thread-block cluster + the collapsed tail in its block 0
        MidArgs m = {};
        const int f2 = h->first_smem;
        m.ser.cp = cp; m.ser.level0 = f2; m.ser.nlev = h->nlev - f2;
        for (int l = f2; l < h->nlev; ++l) { m.ser.nx[l - f2] = h->nx[l]; m.ser.ny[l - f2] = h->ny[l]; }
        m.ser.coarse_solve_size = c.coarse_solve_size; m.ser.coarse_solver = c.coarse_solver;
        m.ser.smoother = c.smoother; m.ser.restriction = c.restriction;
        m.ser.prof = h->prof_dev;
        m.level0 = fs; m.ndist = f2 - fs; m.base = h->mid_base;
        for (int l = fs; l < f2; ++l) { m.nx[l - fs] = h->nx[l]; m.ny[l - fs] = h->ny[l]; }
        m.rhs_in = h->rhs[fs]; m.u_out = h->u[fs];
        cudaLaunchConfig_t lc = {};
        lc.gridDim = dim3(h->mid_nc, 1, 1); lc.blockDim = dim3(kMidThreads, 1, 1);
        lc.dynamicSmemBytes = h->mid_smem; lc.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = h->mid_nc; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        lc.attrs = at; lc.numAttrs = 1;
        B2S_CUDA(cudaLaunchKernelEx(&lc, mg_mid_cluster_kernel, m));
        ++n;
    } else
    // collapsed coarse hierarchy
    {
        CoarseArgs a = {};
        a.cp = cp; a.level0 = fs; a.nlev = h->nlev - fs;
        for (int l = fs; l < h->nlev; ++l) { a.nx[l - fs] = h->nx[l]; a.ny[l - fs] = h->ny[l]; }
        a.rhs_in = fs == 0 ? nullptr : h->rhs[fs];
        a.u_io = fs == 0 ? nullptr : h->u[fs];
        a.u_is_input = fs == 0 ? 1 : 0;
        a.coarse_solve_size = c.coarse_solve_size; a.coarse_solver = c.coarse_solver;
        a.smoother = c.smoother; a.restriction = c.restriction;
        a.sumsq_out = fs == 0 ? h->sumsq_dev : nullptr;
        a.prof = h->prof_dev;
        // one thread per point of the largest resident level (block-wide barriers get cheaper with fewer warps)
        int cthreads = 64;
        for (int l = fs; l < h->nlev; ++l) cthreads = std::max(cthreads, h->nx[l] * h->ny[l]);
        cthreads = std::min(1024, (cthreads + 31) & ~31);
        mg_coarse_kernel<<<1, cthreads, h->coarse_smem, st>>>(a);
        ++n;
    }
    // upward leg
    for (int l = fs - 1; l >= 0 && fused; --l) {
        if (skip(l, 1)) continue;
        TileArgs t = tile_args(l);
        t.u_in = h->tmp[l]; t.u_out = h->u[l]; t.ec = h->u[l + 1]; t.want_norm = (l == 0);
        t.fused_end = (l == 0 && end_fused && only_level < 0) ? 1 : 0;
        if (fused_b) {
            launch_rb_tile(rb_choice(l), true, t, st);
        } else if (use_streaming(l) && warp_kind) {
            const int ch = warp_rows(l);
            mg_up_warp_kernel<<<warp_grid(l, ch), 32, 0, st>>>(t, ch);
        } else if (use_streaming(l) && two_col) {
            const int ch = stream2_rows(l);
            mg_up_stream2_kernel<<<stream2_grid(l, ch), kS2NT + 32, kS2SmemUp, st>>>(t, ch);
        } else if (use_streaming(l)) {
            const int ch = stream_rows(l);
            mg_up_stream_kernel<<<stream_grid(l, ch), kSNT, 0, st>>>(t, ch);
        } else {
            launch_tile(tile_choice_for(l), true, t, st);
        }
        ++n;
    }
    for (int l = fs - 1; l >= 0 && !fused; --l) {
        const int nx = h->nx[l], ny = h->ny[l];
        const int rows = rows_for(nx, ny);
        if (rb) {
            ProlongArgs p = {};
            p.cp = cp; p.level = l; p.coarse = h->u[l + 1]; p.fine = h->u[l]; p.nx = nx; p.ny = ny;
            p.nxc = h->nx[l + 1]; p.nyc = h->ny[l + 1]; p.rows = rows; p.mode = 1;
            mg_prolong_kernel<<<sweep_grid(nx, ny, rows), kMGBX, 0, st>>>(p);
            ++n;
            smooth2(l, l == 0);
        } else {
            // fused prolongation + correction + first post-smoothing sweep, then the second sweep
            ProlongSmoothArgs p = {};
            p.cp = cp; p.level = l; p.coarse = h->u[l + 1]; p.u = h->u[l]; p.rhs = h->rhs[l]; p.out = h->tmp[l];
            p.nx = nx; p.ny = ny; p.nxc = h->nx[l + 1]; p.nyc = h->ny[l + 1]; p.rows = rows;
            mg_prolong_smooth_kernel<<<sweep_grid(nx, ny, rows), kMGBX, 0, st>>>(p);
            ++n;
            SweepArgs a = {};
            a.cp = cp; a.level = l; a.nx = nx; a.ny = ny; a.rows = rows; a.alpha = 4.0 / 5.0; a.mode = 1;
            a.rhs = h->rhs[l]; a.u = h->tmp[l]; a.out = h->u[l]; a.swap_io = 1;
            a.want_norm = (l == 0);
            a.partials = h->partials; a.ticket = h->ticket; a.sumsq_out = h->sumsq_dev;
            mg_sweep_kernel<<<sweep_grid(nx, ny, rows), kMGBX, 0, st>>>(a);
            ++n;
        }
    }
    if (!end_fused && only_level < 0) {
        mg_cycle_end_kernel<<<1, 32, 0, st>>>(h->call_dev);  // r_rms, exit test, bookkeeping (multigrid.jl:64-75)
        ++n;
    }
    B2S_CUDA(cudaGetLastError());
    if (count) *count = n;
    return B2S_OK;
}

// cg!(x_in, b, hx, hy, c, tol, Nmax) with global-memory kernels, host-driven like the reference (krylov.jl:55-91).
// work: 4*nx*ny doubles. Synchronises `st`.
int cg_global(double *x_in, const double *b, double *work, double hx, double hy, double c, double tol, int nmax, int nx, int ny,
              cudaStream_t st, double *ss_out, int *iters_out, long long *count)
{
    const size_t n = (size_t)nx * ny, bytes = n * sizeof(double);
    double *r = work, *p = work + n, *ph = work + 2 * n, *x = work + 3 * n;
    long long nl = 0;
    double v = 0.0;
    B2S_CHECK(b2s_sumsq(b, n, &v, st)); ++nl;
    const double normb = sqrt(v), tolb = tol * normb;
    B2S_CUDA(cudaMemcpyAsync(r, b, bytes, cudaMemcpyDeviceToDevice, st));
    B2S_CUDA(cudaMemcpyAsync(p, b, bytes, cudaMemcpyDeviceToDevice, st));
    B2S_CUDA(cudaMemcpyAsync(ph, b, bytes, cudaMemcpyDeviceToDevice, st));
    B2S_CUDA(cudaMemsetAsync(x, 0, bytes, st));
    double rho = 0.0;
    B2S_CHECK(b2s_dot(r, r, n, &rho, st)); ++nl;
    int it = 0;
    double rr = rho;
    for (int k = 1; k <= nmax; ++k) {
        it = k;
        B2S_CHECK(b2s_matvec2d(p, hx, hy, c, ph, nx, ny, B2S_POLICY_PARALLEL, st));
        double pAp = 0.0;
        B2S_CHECK(b2s_dot(p, ph, n, &pAp, st));
        const double alpha = rho / pAp;
        B2S_CHECK(b2s_axpy(alpha, p, x, n, st));
        B2S_CHECK(b2s_axpy(-alpha, ph, r, n, st));  // r .-= alpha .* p_hat  (r + (-alpha)*ph == r - alpha*ph exactly)
        B2S_CHECK(b2s_sumsq(r, n, &rr, st));
        nl += 5;
        if (sqrt(rr) < tolb) break;
        const double rho_old = rho;
        rho = rr;
        const double beta = rho / rho_old;
        B2S_CHECK(b2s_xpby(r, beta, p, n, st)); ++nl;
    }
    B2S_CUDA(cudaMemcpyAsync(x_in, x, bytes, cudaMemcpyDeviceToDevice, st));
    B2S_CUDA(cudaStreamSynchronize(st));
    if (ss_out) *ss_out = rr;
    if (iters_out) *iters_out = it;
    if (count) *count = nl;
    return B2S_OK;
}

// Coarsest-level solve in global memory (multigrid.jl:145-167) for levels that do not fit into shared memory:
// Jacobi sweeps with the exit test on the device, launched in batches and polled; or the host-driven CG above.
int coarse_global_solve(b2s_mg *h, cudaStream_t st, long long *count)
{
    const int l = h->nlev - 1;
    const int nx = h->nx[l], ny = h->ny[l];
    const size_t n = (size_t)nx * ny;
    const MGCall &m = *h->call_pin;
    double hl = m.h;
    for (int i = 0; i < l; ++i) hl = hl * 2;
    double *u = l == 0 ? m.u : h->u[l];
    const double *rhs = l == 0 ? m.rhs : h->rhs[l];
    const int iters = 20 * h->cfg.coarse_solve_size;
    long long nl = 0;
    if (h->cfg.coarse_solver == B2S_COARSE_CG) {
        double ss = 0.0;
        int it = 0;
        long long c2 = 0;
        B2S_CHECK(cg_global(u, rhs, h->cg_work, hl, hl, m.c, m.tol, iters, nx, ny, st, &ss, &it, &c2));
        nl += c2;
        h->last_sweeps_host = it;
        if (l == 0) B2S_CUDA(cudaMemcpyAsync(h->sumsq_dev, &ss, sizeof(double), cudaMemcpyHostToDevice, st));
    } else {
        mg_reduce_kernel<<<kReduceBlocks, kReduceThreads, 0, st>>>(rhs, rhs, n, h->partials, h->ticket, h->sumsq_dev + 4, nullptr, 0);
        mg_coarse_loop_init_kernel<<<1, 32, 0, st>>>(h->loop_dev, h->sumsq_dev + 4, h->call_dev, (double)nx * ny, iters);
        nl += 2;
        const int rows = rows_for(nx, ny);
        int launched = 0;
        CoarseLoop cur = {};
        const bool rb = h->cfg.smoother == B2S_SMOOTH_RBGS;
        while (!cur.done) {
            const int batch = std::min(32, iters - launched);
            for (int k = 0; k < batch && rb; ++k, ++launched) {  // red-black Gauss-Seidel sweeps, in place
                for (int colour = 0; colour < 2; ++colour) {
                    RbgsArgs a = {};
                    a.u = u; a.rhs = rhs; a.nx = nx; a.ny = ny; a.rows = rows; a.colour = colour; a.h = hl; a.c = m.c;
                    a.want_norm = 1; a.partials = h->partials; a.ticket = h->ticket; a.sumsq_out = h->sumsq_dev + 2;
                    a.loop = h->loop_dev;
                    dim3 g(((nx + 1) / 2 + kMGBX - 1) / kMGBX, (ny + rows - 1) / rows, 1);
                    mg_rbgs_kernel<<<g, kMGBX, 0, st>>>(a);
                }
                mg_coarse_loop_rb_step_kernel<<<1, 32, 0, st>>>(h->loop_dev, h->sumsq_dev);
                nl += 3;
            }
            for (int k = 0; k < batch && !rb; ++k, ++launched) {
                SweepArgs a = {};
                a.nx = nx; a.ny = ny; a.rows = rows; a.h = hl; a.c = m.c; a.alpha = 4.0 / 5.0; a.mode = 1;
                a.rhs = rhs;
                a.u = (launched & 1) ? h->tmp[l] : u;
                a.out = (launched & 1) ? u : h->tmp[l];
                a.want_norm = 1; a.partials = h->partials; a.ticket = h->ticket; a.sumsq_out = h->sumsq_dev;
                a.loop = h->loop_dev;
                mg_sweep_kernel<<<sweep_grid(nx, ny, rows), kMGBX, 0, st>>>(a);
                ++nl;
            }
            B2S_CUDA(cudaGetLastError());
            B2S_CUDA(cudaMemcpyAsync(h->loop_pin, h->loop_dev, sizeof(CoarseLoop), cudaMemcpyDeviceToHost, st));
            B2S_CUDA(cudaStreamSynchronize(st));
            cur = *h->loop_pin;
            launched = cur.sweeps;  // launches after the exit were no-ops
            if (launched >= iters) break;
        }
        if (!rb && (cur.sweeps & 1))  // odd number of Jacobi sweeps: the result sits in tmp
            B2S_CUDA(cudaMemcpyAsync(u, h->tmp[l], n * sizeof(double), cudaMemcpyDeviceToDevice, st));
        h->last_sweeps_host = cur.sweeps;
    }
    if (count) *count = nl;
    return B2S_OK;
}

int launch_cycle(b2s_mg *h)
{
    const int bc = h->call_pin->bc_before ? 1 : 0;  // two graphs: with / without the boundary-condition node
    if (h->cfg.use_graph) {
        if (!h->graph[bc]) {
            cudaGraph_t g = nullptr;
            B2S_CUDA(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
            long long n = 0;
            int rc = enqueue_vcycle(h, h->stream, &n, bc != 0);
            cudaError_t e = cudaStreamEndCapture(h->stream, &g);
            if (rc != B2S_OK) { if (g) cudaGraphDestroy(g); return rc; }
            B2S_CUDA(e);
            h->launches_per_cycle[bc] = n;
            B2S_CUDA(cudaGraphInstantiate(&h->graph[bc], g, 0));
            B2S_CUDA(cudaGraphDestroy(g));
        }
        B2S_CUDA(cudaGraphLaunch(h->graph[bc], h->stream));
        h->kernel_launches += h->launches_per_cycle[bc];
    } else {
        long long n = 0;
        B2S_CHECK(enqueue_vcycle(h, h->stream, &n, bc != 0));
        h->kernel_launches += n;
    }
    return B2S_OK;
}

inline MGCall *call_back(b2s_mg *h) { return reinterpret_cast<MGCall *>(reinterpret_cast<char *>(h->call_pin) + kCallBlockBytes); }

int set_call(b2s_mg *h, double *u, const double *f, double hgrid, double c, double tol, int apply_bcs, int bc_before,
             int niters, int check)
{
    MGCall &m = *h->call_pin;
    m.u = u; m.rhs = f; m.h = hgrid; m.c = c; m.tol = tol; m.apply_bcs = apply_bcs; m.bc_before = bc_before;
    m.sumsq = h->sumsq_dev; m.coarse_sweeps = h->sweeps_dev;
    m.done = niters > 0 ? 0 : 1; m.ncycles = 0; m.niters = niters; m.check = check;
    m.n_points = (double)h->nx[0] * h->ny[0];
    m.hist = h->hist_dev;
    const int fs = h->coarse_global ? h->nlev - 1 : h->first_smem;
    // unfused red-black sweeps deposit one sum per colour; the fused upward kernel sums both colours itself
    m.rb_combine = (h->cfg.smoother == B2S_SMOOTH_RBGS && fs > 0 && !fused_variant_b(h->cfg)) ? 1 : 0;
    m.pad = 0;
    // per-level constants, in the arithmetic of make_coef / make_rb_coef / make_div_h2 (this file is compiled without
    // floating-point contraction on the host side as well)
    LevelCoef *lev = reinterpret_cast<LevelCoef *>(h->call_pin + 1);
    m.lev = reinterpret_cast<const LevelCoef *>(h->call_dev + 1);
    double hl = hgrid;
    for (int l = 0; l < h->nlev; ++l) {
        LevelCoef &L = lev[l];
        L.h = hl;
        L.C = 4.0 + c * (hl * hl);
        L._h2 = 1 / (hl * hl);
        L.wJ = (4.0 / 5.0) * ((hl * hl) / (4.0 + c * (hl * hl)));
        L.h2 = hl * hl;
        L.wGS = 1.0 * ((hl * hl) / (4.0 + c * (hl * hl)));
        unsigned long long b;
        memcpy(&b, &L.h2, sizeof b);
        const int e = (int)((b >> 52) & 0x7ff);
        L.exact = ((long long)b > 0 && (b & 0x000fffffffffffffULL) == 0 && e >= 2 && e <= 2044) ? 1 : 0;
        L.inv_h2 = L.exact ? 1.0 / L.h2 : 0.0;
        L.pad = 0;
        hl = hl * 2;
    }
    B2S_CUDA(cudaMemcpyAsync(h->call_dev, h->call_pin, kCallBlockBytes, cudaMemcpyHostToDevice, h->stream));
    return B2S_OK;
}

int mg_destroy_impl(b2s_mg *h)
{
    if (!h) return B2S_OK;
    cudaSetDevice(h->cfg.device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    for (int g_ = 0; g_ < 2; ++g_)
        if (h->graph[g_]) cudaGraphExecDestroy(h->graph[g_]);
    for (int l = 0; l < kMaxLevels; ++l) {
        if (h->u[l]) cudaFree(h->u[l]);
        if (h->rhs[l]) cudaFree(h->rhs[l]);
        if (h->tmp[l]) cudaFree(h->tmp[l]);
    }
    if (h->call_dev) cudaFree(h->call_dev);
    if (h->call_pin) cudaFreeHost(h->call_pin);
    if (h->hist_dev) cudaFree(h->hist_dev);
    if (h->hist_pin) cudaFreeHost(h->hist_pin);
    if (h->sumsq_dev) cudaFree(h->sumsq_dev);
    if (h->sumsq_pin) cudaFreeHost(h->sumsq_pin);
    if (h->sweeps_dev) cudaFree(h->sweeps_dev);
    if (h->partials) cudaFree(h->partials);
    if (h->ticket) cudaFree(h->ticket);
    if (h->prof_dev) cudaFree(h->prof_dev);
    if (h->loop_dev) cudaFree(h->loop_dev);
    if (h->loop_pin) cudaFreeHost(h->loop_pin);
    if (h->cg_work) cudaFree(h->cg_work);
    if (h->pcg_work) cudaFree(h->pcg_work);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return B2S_OK;
}

// sum f^2 over all entries into sumsq_dev[1]
int enqueue_fnorm(b2s_mg *h)
{
    mg_reduce_kernel<<<kReduceBlocks, kReduceThreads, 0, h->stream>>>(nullptr, nullptr, (size_t)h->nx[0] * h->ny[0], h->partials,
                                                                      h->ticket, h->sumsq_dev + 1, h->call_dev, 1);
    B2S_CUDA(cudaGetLastError());
    h->kernel_launches += 1;
    return B2S_OK;
}

}  // namespace

// internal (ns2d.cu): the stream all work of a handle is ordered on
cudaStream_t b2s_mg_stream_internal(b2s_mg *h) { return h->stream; }
void b2s_mg_count_launches_internal(b2s_mg *h, long long n) { h->kernel_launches += n; }

extern "C" {

int b2s_mg_create(b2s_mg **out, const b2s_mg_config *cfg)
{
    B2S_REQUIRE(out && cfg, B2S_ERR_BAD_ARG, "NULL argument");
    *out = nullptr;
    const int nx = cfg->nx, ny = cfg->ny, cs = cfg->coarse_solve_size;
    B2S_REQUIRE(nx >= 3 && ny >= 3, B2S_ERR_BAD_SIZE, "grid %dx%d too small", nx, ny);
    // asserts multigrid.jl:45-46
    B2S_REQUIRE(cs <= std::min(nx, ny), B2S_ERR_BAD_SIZE, "coarse_solve_size %d > min(nx, ny) = %d", cs, std::min(nx, ny));
    B2S_REQUIRE(cs >= 2 && pow2(cs - 1), B2S_ERR_BAD_SIZE, "coarse_solve_size - 1 = %d is not a power of 2", cs - 1);
    B2S_REQUIRE(cfg->coarse_solver == B2S_COARSE_JACOBI || cfg->coarse_solver == B2S_COARSE_CG, B2S_ERR_BAD_ARG, "bad coarse_solver");
    B2S_REQUIRE(cfg->smoother == B2S_SMOOTH_JACOBI || cfg->smoother == B2S_SMOOTH_RBGS, B2S_ERR_BAD_ARG, "bad smoother");
    B2S_REQUIRE(cfg->restriction == B2S_RESTRICT_INJECT || cfg->restriction == B2S_RESTRICT_FW, B2S_ERR_BAD_ARG, "bad restriction");
    int ndev = 0;
    B2S_CHECK(b2s_device_count(&ndev));
    B2S_REQUIRE(ndev > 0, B2S_ERR_NO_DEVICE, "no CUDA device (there is no CPU fallback)");
    B2S_REQUIRE(cfg->device >= 0 && cfg->device < ndev, B2S_ERR_BAD_ARG, "device %d out of range", cfg->device);

    b2s_mg *h = new b2s_mg();
    h->cfg = *cfg;
    // level table: halve until min(nx, ny) <= coarse_solve_size; every halving needs even nx-1, ny-1 (multigrid.jl:95-97)
    int lx = nx, ly = ny, L = 0;
    for (;;) {
        if (L >= kMaxLevels) { delete h; set_error("too many levels"); return B2S_ERR_BAD_SIZE; }
        h->nx[L] = lx; h->ny[L] = ly; ++L;
        if (((lx - 1) & 1) || ((ly - 1) & 1)) {
            delete h;
            set_error("ERROR:not a power of 2 (level %d is %dx%d)", L - 1, lx, ly);
            return B2S_ERR_BAD_SIZE;
        }
        if (std::min(lx, ly) <= cs) break;
        lx = 1 + (lx - 1) / 2; ly = 1 + (ly - 1) / 2;
    }
    h->nlev = L;
    // which levels fit into the collapsed shared-memory kernel (3 arrays per level + CG scratch of the coarsest)
    {
        size_t bytes = 4 * (size_t)h->nx[L - 1] * h->ny[L - 1] * 8 + 64;
        int fs = L;
        for (int l = L - 1; l >= 0; --l) {
            const size_t add = 3 * (size_t)h->nx[l] * h->ny[l] * 8;
            if (bytes + add > kCoarseSmemLimit) break;
            if (!cfg->smem_levels && l < L - 1) break;
            // levels above this size run faster as tile kernels spread over many SMs than inside the one-block kernel
            if (l < L - 1 && (fused_variant_a(*cfg) || fused_variant_b(*cfg)) &&
                (size_t)h->nx[l] * h->ny[l] > smem_level_max_points()) break;
            bytes += add;
            fs = l;
        }
        if (fs == L) {  // coarsest level does not fit into shared memory: global-memory coarsest solve, host-polled
            h->coarse_global = true;
            h->cfg.use_graph = 0;  // the coarsest solve polls the device between batches of sweeps
            bytes = 0;
        }
        h->first_smem = fs;
        h->coarse_smem = bytes;
        h->mid_first = fs;
    }
    DeviceGuard guard;
    guard.set(cfg->device);
#define MG_CUDA(call)                                                                          \
    do {                                                                                       \
        cudaError_t e_ = (call);                                                               \
        if (e_ != cudaSuccess) {                                                               \
            set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_));   \
            mg_destroy_impl(h);                                                                \
            return B2S_ERR_CUDA;                                                               \
        }                                                                                      \
    } while (0)
    MG_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    MG_CUDA(cudaEventCreate(&h->ev0));
    MG_CUDA(cudaEventCreate(&h->ev1));
    for (int l = 0; l < L; ++l) {
        const size_t bytes = (size_t)h->nx[l] * h->ny[l] * 8;
        if (l < std::max(h->first_smem, 1)) {
            MG_CUDA(cudaMalloc(&h->tmp[l], bytes));
            MG_CUDA(cudaMemset(h->tmp[l], 0, bytes));
        }
        if (l >= 1 && l <= h->first_smem && l < L) {
            MG_CUDA(cudaMalloc(&h->u[l], bytes));
            MG_CUDA(cudaMalloc(&h->rhs[l], bytes));
            MG_CUDA(cudaMemset(h->u[l], 0, bytes));
            MG_CUDA(cudaMemset(h->rhs[l], 0, bytes));
        }
    }
    if (h->coarse_global) {
        MG_CUDA(cudaMalloc(&h->loop_dev, sizeof(CoarseLoop)));
        MG_CUDA(cudaMallocHost(&h->loop_pin, sizeof(CoarseLoop)));
        if (cfg->coarse_solver == B2S_COARSE_CG)
            MG_CUDA(cudaMalloc(&h->cg_work, 4 * (size_t)h->nx[L - 1] * h->ny[L - 1] * sizeof(double)));
    }
    MG_CUDA(cudaMalloc(&h->call_dev, kCallBlockBytes));
    MG_CUDA(cudaMallocHost(&h->call_pin, kCallBlockBytes + sizeof(MGCall)));
    MG_CUDA(cudaMalloc(&h->hist_dev, kMaxHist * sizeof(double)));
    MG_CUDA(cudaMallocHost(&h->hist_pin, kMaxHist * sizeof(double)));
    MG_CUDA(cudaMalloc(&h->sumsq_dev, 8 * sizeof(double)));
    MG_CUDA(cudaMemset(h->sumsq_dev, 0, 8 * sizeof(double)));
    MG_CUDA(cudaMallocHost(&h->sumsq_pin, 8 * sizeof(double)));
    MG_CUDA(cudaMalloc(&h->sweeps_dev, 4 * sizeof(int)));
    MG_CUDA(cudaMemset(h->sweeps_dev, 0, 4 * sizeof(int)));
    MG_CUDA(cudaMalloc(&h->partials, sizeof(double) * kMaxPartials));
    MG_CUDA(cudaMalloc(&h->ticket, 64));
    MG_CUDA(cudaMemset(h->ticket, 0, 64));
    MG_CUDA(cudaFuncSetAttribute(mg_coarse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kCoarseSmemLimit + 1024));
    // cluster kernel for the middle levels: the largest cluster (16 blocks: non-portable size, 8: portable) whose blocks can
    // hold their bands; the finest distributed level is the largest one that still fits (never level 0: its arrays are the
    // caller's and its upward kernel finishes the cycle). B2S_MG_CLUSTER = 0 disables, 8 / 16 pins the cluster size.
    if (fused_variant_a(*cfg) && !h->coarse_global && h->first_smem >= 2) {
        const char *ec = getenv("B2S_MG_CLUSTER");
        // Default: OFF. Measured on B200 (profiles/r02_mid_cluster_*): bit-identical, 13 -> 7 launches, but 0.0813 ms per
        // 1025^2 V-cycle against 0.0784 with the six small tile kernels -- inside a CUDA graph those cost ~1.85 us each, and a
        // sweep that exchanges rows between blocks cannot go below ~700 cycles (DSMEM store + mbarrier wake-up + block barrier).
        const int want = (ec && *ec) ? atoi(ec) : 0;
        const char *emp = getenv("B2S_MG_CLUSTER_MAXPTS");  // largest level kept in the cluster (larger ones are smem-bandwidth
        const size_t maxpts = (emp && *emp) ? (size_t)atoll(emp) : (size_t)20000;  // bound on 16 SMs: measured, 257^2 loses)
        const int last = h->first_smem - 1;
        const size_t limit = 200 * 1024;  // + 24 KB of static shared memory (the in-warp coarsest solver's batches) <= 227 KB
        const int tries[2] = {16, 8};
        for (int t_ = 0; t_ < 2 && want != 0 && h->mid_nc == 0; ++t_) {
            const int NC = tries[t_];
            if (want > 0 && want != NC) continue;
            if ((h->ny[last] - 1) % NC != 0 || (h->ny[last] - 1) / NC < 2) continue;
            MidArgs m = {};
            m.base = (h->ny[last] - 1) / NC;
            const size_t ser0 = (size_t)h->nx[h->first_smem] * h->ny[h->first_smem] * sizeof(double);  // the broadcast correction
            int top = -1;
            for (int cand = last; cand >= 1 && last - cand + 1 <= kMidMaxDist; --cand) {
                m.ndist = last - cand + 1;
                for (int l = cand; l <= last; ++l) { m.nx[l - cand] = h->nx[l]; m.ny[l - cand] = h->ny[l]; }
                if ((size_t)h->nx[cand] * h->ny[cand] > maxpts) break;
                if (mid_dist_doubles(m) * sizeof(double) + h->coarse_smem + ser0 > limit) break;
                top = cand;
            }
            if (top < 0) continue;
            m.ndist = last - top + 1;
            for (int l = top; l <= last; ++l) { m.nx[l - top] = h->nx[l]; m.ny[l - top] = h->ny[l]; }
            const size_t smem = mid_dist_doubles(m) * sizeof(double) + h->coarse_smem + ser0;
            MG_CUDA(cudaFuncSetAttribute(mg_mid_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)limit));
            if (NC > 8) MG_CUDA(cudaFuncSetAttribute(mg_mid_cluster_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
            cudaLaunchConfig_t lc = {};
            lc.gridDim = dim3(NC, 1, 1); lc.blockDim = dim3(kMidThreads, 1, 1); lc.dynamicSmemBytes = smem;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = NC; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            lc.attrs = at; lc.numAttrs = 1;
            int nclusters = 0;
            if (cudaOccupancyMaxActiveClusters(&nclusters, mg_mid_cluster_kernel, &lc) != cudaSuccess || nclusters < 1) {
                cudaGetLastError();
                continue;
            }
            h->mid_nc = NC; h->mid_base = m.base; h->mid_first = top; h->mid_smem = smem;
        }
    }
    {
        const char *e = getenv("B2S_MG_TILE");
        h->tile_choice = (e && *e) ? atoi(e) : -1;  // -1: per level (tile_choice_for)
        if (h->tile_choice < -1 || h->tile_choice >= kTileChoices) h->tile_choice = -1;
        for (int t_ = 0; t_ < kTileChoices; ++t_) MG_CUDA(tile_set_attr(t_));
        const char *e7 = getenv("B2S_MG_TILE_MINBLOCKS");
        if (e7 && *e7) h->tile_min_blocks = atoi(e7);
        const char *e6 = getenv("B2S_MG_RB_TILE");
        h->rb_tile_choice = (e6 && *e6) ? atoi(e6) : -1;  // -1: per level (rb_choice)
        if (h->rb_tile_choice < -1 || h->rb_tile_choice >= kRbTileChoices) h->rb_tile_choice = -1;
        for (int t_ = 0; t_ < kRbTileChoices; ++t_) MG_CUDA(rb_tile_set_attr(t_));
        MG_CUDA(cudaFuncSetAttribute(mg_down_stream2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kS2SmemDown));
        MG_CUDA(cudaFuncSetAttribute(mg_up_stream2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kS2SmemUp));
        const char *e2 = getenv("B2S_MG_CH");
        h->stream_ch = (e2 && *e2) ? atoi(e2) : 0;
        const char *e4 = getenv("B2S_MG_STREAM_MIN");
        if (e4 && *e4) h->stream_min_points = (size_t)atoll(e4);
        const char *e5 = getenv("B2S_MG_STREAM_KIND");
        if (e5 && *e5) h->stream_warp = (strcmp(e5, "warp") == 0);
        const char *e3 = getenv("B2S_MG_PROF");
        if (e3 && *e3 == '1') {
            MG_CUDA(cudaMalloc(&h->prof_dev, 64 * sizeof(long long)));
            MG_CUDA(cudaMemset(h->prof_dev, 0, 64 * sizeof(long long)));
        }
    }
#undef MG_CUDA
    *out = h;
    return B2S_OK;
}

int b2s_mg_destroy(b2s_mg *h)
{
    int prev = -1;
    cudaGetDevice(&prev);
    mg_destroy_impl(h);
    if (prev >= 0) cudaSetDevice(prev);
    return B2S_OK;
}

int b2s_mg_solve(b2s_mg *h, double *u, const double *f, double hgrid, double c, double tol, int niters, int apply_bcs,
                 double *r_rms_out, int *ncycles, double *rel_hist)
{
    B2S_REQUIRE(h && u && f, B2S_ERR_BAD_ARG, "NULL argument");
    B2S_REQUIRE(niters <= kMaxHist, B2S_ERR_BAD_ARG, "niters %d exceeds %d", niters, kMaxHist);
    DeviceGuard guard;
    guard.set(h->cfg.device);
    const double N = (double)h->nx[0] * h->ny[0];
    B2S_CUDA(cudaEventRecord(h->ev0, h->stream));
    B2S_CHECK(set_call(h, u, f, hgrid, c, tol, apply_bcs, apply_bcs, niters, 1));
    B2S_CHECK(enqueue_fnorm(h));  // f_rms = sqrt(sum(f.^2)/(nx*ny))   multigrid.jl:53
    // The loop "for iter = 1:niters ... break if r_rms < tolf" (multigrid.jl:58-76) runs on the device: the last kernel
    // of every cycle evaluates the test and raises a flag that turns all later cycles into no-ops. The host enqueues
    // cycles in batches sized from the observed contraction factor and polls one struct per batch.
    MGCall *back = call_back(h);
    back->done = niters > 0 ? 0 : 1;
    back->ncycles = 0;
    int launched = 0;
    // Solves of the same kind (same BC handling, Poisson or Helmholtz) tend to need the same number of cycles -- the three
    // solves of every Navier-Stokes step, repeated benchmark solves: the first batch is the count the last such solve
    // needed (cycles past convergence are no-ops), so a steady-state solve costs one host synchronisation.
    int &remembered = h->last_ncycles[apply_bcs ? 1 : 0][c != 0.0 ? 1 : 0];
    while (!back->done) {
        int batch = 1;
        if (!h->coarse_global) {
            const int k = back->ncycles;
            if (k == 0 && remembered > 0) batch = remembered;
            else if (k < 2) batch = 2 - k;
            else {  // predict the remaining cycles from the last contraction factor
                const double a = h->hist_pin[k - 1], b = h->hist_pin[k - 2];
                const double ratio = a / b;
                batch = 1;
                if (ratio > 0.0 && ratio < 0.9 && a > tol) batch = (int)ceil(log(tol / a) / log(ratio));
                batch = std::max(1, std::min(batch, 8));
            }
        }
        batch = std::min(batch, niters - launched);
        if (batch <= 0) batch = 1;
        for (int i = 0; i < batch; ++i) B2S_CHECK(launch_cycle(h));
        launched += batch;
        B2S_CUDA(cudaMemcpyAsync(back, h->call_dev, sizeof(MGCall), cudaMemcpyDeviceToHost, h->stream));
        B2S_CUDA(cudaMemcpyAsync(h->hist_pin, h->hist_dev, sizeof(double) * std::min(launched, kMaxHist), cudaMemcpyDeviceToHost,
                                 h->stream));
        B2S_CUDA(cudaStreamSynchronize(h->stream));  // @synchronize()   multigrid.jl:65
        launched = back->ncycles;  // cycles enqueued after the exit were no-ops
    }
    B2S_CUDA(cudaMemcpyAsync(h->sumsq_pin, h->sumsq_dev, 2 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    B2S_CUDA(cudaEventRecord(h->ev1, h->stream));
    B2S_CUDA(cudaEventSynchronize(h->ev1));
    float ms = 0.f;
    B2S_CUDA(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    h->last_ms = ms;
    const int n = back->ncycles;
    remembered = n;
    if (rel_hist)
        for (int i = 0; i < n; ++i) rel_hist[i] = h->hist_pin[i];
    if (r_rms_out) *r_rms_out = n > 0 ? sqrt(h->sumsq_pin[0] / N) : 0.0;
    if (ncycles) *ncycles = n;
    return B2S_OK;
}

int b2s_mg_vcycle(b2s_mg *h, double *u, const double *rhs, double hgrid, double c, double tol, int apply_bcs, double *res_rms)
{
    B2S_REQUIRE(h && u && rhs, B2S_ERR_BAD_ARG, "NULL argument");
    DeviceGuard guard;
    guard.set(h->cfg.device);
    B2S_CHECK(set_call(h, u, rhs, hgrid, c, tol, apply_bcs, 0, 1, 0));
    B2S_CHECK(launch_cycle(h));
    B2S_CUDA(cudaMemcpyAsync(h->sumsq_pin, h->sumsq_dev, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    B2S_CUDA(cudaStreamSynchronize(h->stream));
    if (res_rms) *res_rms = sqrt(h->sumsq_pin[0] / ((double)h->nx[0] * h->ny[0]));
    if (h->prof_dev) {
        long long p[64];
        B2S_CUDA(cudaMemcpy(p, h->prof_dev, sizeof(p), cudaMemcpyDeviceToHost));
        fprintf(stderr, "[b2s mg_coarse_kernel phases, SM cycles since entry]");
        for (long long i = 1; i < p[0] && i < 60; ++i) fprintf(stderr, " %lld", p[1 + i] - p[1]);
        fprintf(stderr, "\n");
    }
    return B2S_OK;
}

int b2s_mg_cycles(b2s_mg *h, double *u, const double *f, double hgrid, double c, double tol, int ncycles, int apply_bcs,
                  double *r_rms_last, double *ms_out)
{
    B2S_REQUIRE(h && u && f && ncycles >= 0, B2S_ERR_BAD_ARG, "bad argument");
    DeviceGuard guard;
    guard.set(h->cfg.device);
    B2S_CHECK(set_call(h, u, f, hgrid, c, tol, apply_bcs, apply_bcs, 1 << 30, 0));
    if (h->cfg.use_graph && !h->graph[apply_bcs ? 1 : 0] && ncycles > 0) {  // keep graph instantiation out of the timed region
        B2S_CHECK(launch_cycle(h));
        B2S_CUDA(cudaStreamSynchronize(h->stream));
        --ncycles;
    }
    B2S_CUDA(cudaEventRecord(h->ev0, h->stream));
    for (int i = 0; i < ncycles; ++i) B2S_CHECK(launch_cycle(h));
    B2S_CUDA(cudaEventRecord(h->ev1, h->stream));
    B2S_CUDA(cudaMemcpyAsync(h->sumsq_pin, h->sumsq_dev, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    B2S_CUDA(cudaStreamSynchronize(h->stream));
    float ms = 0.f;
    B2S_CUDA(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    h->last_ms = ms;
    if (ms_out) *ms_out = ms;
    if (r_rms_last) *r_rms_last = sqrt(h->sumsq_pin[0] / ((double)h->nx[0] * h->ny[0]));
    return B2S_OK;
}

int b2s_mg_pcg_solve(b2s_mg *h, double *u, const double *f, double hgrid, double c, double tol, int maxit, double *r_rms_out,
                     int *iters_out)
{
    return b2s_mg_pcg_solve2(h, u, f, hgrid, c, tol, maxit, B2S_PCG_TOL_INITIAL_RESIDUAL, r_rms_out, iters_out);
}

int b2s_mg_pcg_solve2(b2s_mg *h, double *u, const double *f, double hgrid, double c, double tol, int maxit, int tol_mode,
                      double *r_rms_out, int *iters_out)
{
    B2S_REQUIRE(h && u && f && maxit >= 0, B2S_ERR_BAD_ARG, "bad argument");
    B2S_REQUIRE(tol_mode == B2S_PCG_TOL_INITIAL_RESIDUAL || tol_mode == B2S_PCG_TOL_RHS, B2S_ERR_BAD_ARG, "bad tol_mode");
    B2S_REQUIRE(h->cfg.restriction == B2S_RESTRICT_FW, B2S_ERR_BAD_ARG,
                "MG-preconditioned CG needs a symmetric V-cycle: use restriction = B2S_RESTRICT_FW");
    DeviceGuard guard;
    guard.set(h->cfg.device);
    const int nx = h->nx[0], ny = h->ny[0];
    const size_t n = (size_t)nx * ny, bytes = n * sizeof(double);
    if (!h->pcg_work) B2S_CUDA(cudaMalloc(&h->pcg_work, 4 * bytes));
    double *r = h->pcg_work, *z = r + n, *p = r + 2 * n, *q = r + 3 * n;
    cudaStream_t st = h->stream;
    const double N = (double)nx * ny;
    const int rows = rows_for(nx, ny);
    B2S_CUDA(cudaEventRecord(h->ev0, st));
    B2S_CUDA(cudaMemsetAsync(q, 0, bytes, st));
    B2S_CHECK(b2s_matvec2d(u, hgrid, hgrid, c, q, nx, ny, B2S_POLICY_PARALLEL, st));
    mg_pcg_residual_kernel<<<sweep_grid(nx, ny, rows), kMGBX, 0, st>>>(f, q, r, nx, ny, rows);
    B2S_CUDA(cudaGetLastError());
    double ss = 0.0;
    B2S_CHECK(b2s_sumsq(r, n, &ss, st));
    h->kernel_launches += 3;
    double r_rms = sqrt(ss / N);
    double tolf = tol * r_rms;
    if (tol_mode == B2S_PCG_TOL_RHS) {  // MGsolve's criterion: r_rms < tol * f_rms, f_rms over all entries (multigrid.jl:53,70-75)
        double sf = 0.0;
        B2S_CHECK(b2s_sumsq(f, n, &sf, st));
        h->kernel_launches += 1;
        tolf = tol * sqrt(sf / N);
    }
    int it = 0;
    double rz = 0.0;
    for (int k = 1; k <= maxit && r_rms >= tolf && r_rms > 0.0; ++k) {
        it = k;
        B2S_CUDA(cudaMemsetAsync(z, 0, bytes, st));
        B2S_CHECK(b2s_mg_vcycle(h, z, r, hgrid, c, tol, 0, nullptr));  // z = M r
        double rz_new = 0.0;
        B2S_CHECK(b2s_dot(r, z, n, &rz_new, st));
        if (k == 1) B2S_CUDA(cudaMemcpyAsync(p, z, bytes, cudaMemcpyDeviceToDevice, st));
        else B2S_CHECK(b2s_xpby(z, rz_new / rz, p, n, st));  // p = z + beta p
        rz = rz_new;
        B2S_CUDA(cudaMemsetAsync(q, 0, bytes, st));
        B2S_CHECK(b2s_matvec2d(p, hgrid, hgrid, c, q, nx, ny, B2S_POLICY_PARALLEL, st));
        double pq = 0.0;
        B2S_CHECK(b2s_dot(p, q, n, &pq, st));
        const double alpha = rz / pq;
        B2S_CHECK(b2s_axpy(alpha, p, u, n, st));
        B2S_CHECK(b2s_axpy(-alpha, q, r, n, st));
        B2S_CHECK(b2s_sumsq(r, n, &ss, st));
        h->kernel_launches += 7;
        r_rms = sqrt(ss / N);
    }
    B2S_CUDA(cudaEventRecord(h->ev1, st));
    B2S_CUDA(cudaEventSynchronize(h->ev1));
    float ms = 0.f;
    B2S_CUDA(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    h->last_ms = ms;
    if (r_rms_out) *r_rms_out = r_rms;
    if (iters_out) *iters_out = it;
    return B2S_OK;
}

int b2s_mg_profile_kernels(b2s_mg *h, double *u, const double *f, double hgrid, double c, int reps, int *nlevels_out,
                           double *ms_down, double *ms_up, double *ms_tail, int *level_nx, int *level_ny)
{
    B2S_REQUIRE(h && u && f && reps >= 1 && nlevels_out && ms_down && ms_up && ms_tail, B2S_ERR_BAD_ARG, "bad argument");
    B2S_REQUIRE(fused_variant_a(h->cfg) || fused_variant_b(h->cfg), B2S_ERR_BAD_ARG, "profiling needs the fused level kernels");
    DeviceGuard guard;
    guard.set(h->cfg.device);
    const bool use_mid = !h->coarse_global && h->mid_nc > 0 && h->mid_first < h->first_smem;
    const int fs = h->coarse_global ? h->nlev - 1 : (use_mid ? h->mid_first : h->first_smem);
    B2S_CHECK(set_call(h, u, f, hgrid, c, 1e-6, 0, 0, 1 << 30, 0));
    {  // one whole cycle first: every level array holds representative data
        long long n = 0;
        B2S_CHECK(enqueue_vcycle(h, h->stream, &n, false));
    }
    auto time_one = [&](int level, int dir, double *out) -> int {
        long long n = 0;
        for (int i = 0; i < 3; ++i) B2S_CHECK(enqueue_vcycle(h, h->stream, &n, false, level, dir));  // warm-up
        B2S_CUDA(cudaEventRecord(h->ev0, h->stream));
        for (int i = 0; i < reps; ++i) B2S_CHECK(enqueue_vcycle(h, h->stream, &n, false, level, dir));
        B2S_CUDA(cudaEventRecord(h->ev1, h->stream));
        B2S_CUDA(cudaEventSynchronize(h->ev1));
        float ms = 0.f;
        B2S_CUDA(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
        *out = (double)ms / reps;
        return B2S_OK;
    };
    for (int l = 0; l < fs; ++l) {
        B2S_CHECK(time_one(l, 0, ms_down + l));
        B2S_CHECK(time_one(l, 1, ms_up + l));
        if (level_nx) level_nx[l] = h->nx[l];
        if (level_ny) level_ny[l] = h->ny[l];
    }
    B2S_CHECK(time_one(0, 2, ms_tail));
    *nlevels_out = fs;
    return B2S_OK;
}

int b2s_mg_last_coarse_sweeps(const b2s_mg *h, int *sweeps)
{
    B2S_REQUIRE(h && sweeps, B2S_ERR_BAD_ARG, "NULL argument");
    DeviceGuard guard;
    guard.set(h->cfg.device);
    B2S_CUDA(cudaStreamSynchronize(h->stream));
    if (h->coarse_global) { *sweeps = h->last_sweeps_host; return B2S_OK; }
    B2S_CUDA(cudaMemcpy(sweeps, h->sweeps_dev, sizeof(int), cudaMemcpyDeviceToHost));
    return B2S_OK;
}

int b2s_mg_stats(const b2s_mg *h, long long *kernel_launches, double *last_call_ms)
{
    B2S_REQUIRE(h, B2S_ERR_BAD_ARG, "NULL handle");
    if (kernel_launches) *kernel_launches = h->kernel_launches;
    if (last_call_ms) *last_call_ms = h->last_ms;
    return B2S_OK;
}

// ==================================================================================================================
// L0 wrappers
// ==================================================================================================================
static int check_policy(int policy)
{
    // execution_policy == serial -> error()   multigrid.jl:233-236, krylov.jl:46-49 (serial is a CPU-only debug path)
    B2S_REQUIRE(policy == B2S_POLICY_PARALLEL || policy == B2S_POLICY_PARALLEL_SHMEM, B2S_ERR_NOT_IMPLEMENTED,
                "execution policy %d is not available on the GPU", policy);
    return B2S_OK;
}

int b2s_residual2d(const double *u, const double *f, double h, double c, double *res, int nx, int ny, int policy, void *stream)
{
    B2S_REQUIRE(u && f && res && nx >= 3 && ny >= 3, B2S_ERR_BAD_ARG, "bad argument");
    B2S_CHECK(check_policy(policy));
    SweepArgs a = {};
    a.u = u; a.rhs = f; a.out = res; a.nx = nx; a.ny = ny; a.rows = rows_for(nx, ny); a.h = h; a.c = c; a.alpha = 1.0;
    a.mode = 0;
    mg_sweep_kernel<<<sweep_grid(nx, ny, a.rows), kMGBX, 0, (cudaStream_t)stream>>>(a);
    B2S_CUDA(cudaGetLastError());
    return B2S_OK;
}

int b2s_iteration2d(double *u, const double *f, double h, double c, double *res, int nx, int ny, double alpha, int policy,
                    double *r_rms_host, void *stream)
{
    B2S_REQUIRE(u && f && res && nx >= 3 && ny >= 3, B2S_ERR_BAD_ARG, "bad argument");
    B2S_CHECK(check_policy(policy));
    Scratch *sc = nullptr;
    B2S_CHECK(get_scratch(&sc));
    cudaStream_t st = (cudaStream_t)stream;
    SweepArgs a = {};
    a.u = u; a.rhs = f; a.out = res; a.nx = nx; a.ny = ny; a.rows = rows_for(nx, ny); a.h = h; a.c = c; a.alpha = alpha;
    a.mode = 0; a.want_norm = 1; a.partials = sc->partials; a.ticket = sc->ticket; a.sumsq_out = sc->result;
    mg_sweep_kernel<<<sweep_grid(nx, ny, a.rows), kMGBX, 0, st>>>(a);
    B2S_CUDA(cudaGetLastError());
    const double w = alpha * ((h * h) / (4.0 + c * (h * h)));
    mg_axpy_kernel<<<296, 256, 0, st>>>(w, res, u, (size_t)nx * ny, 0);  // u .+= w .* res (res frame untouched by the kernel)
    B2S_CUDA(cudaGetLastError());
    B2S_CUDA(cudaMemcpyAsync(sc->pinned, sc->result, sizeof(double), cudaMemcpyDeviceToHost, st));
    B2S_CUDA(cudaStreamSynchronize(st));
    if (r_rms_host) *r_rms_host = sqrt(sc->pinned[0] / ((double)nx * ny));
    return B2S_OK;
}

int b2s_rbgs2d(double *u, const double *f, double h, double c, int nx, int ny, double *r_rms_host, void *stream)
{
    B2S_REQUIRE(u && f && nx >= 3 && ny >= 3, B2S_ERR_BAD_ARG, "bad argument");
    Scratch *sc = nullptr;
    B2S_CHECK(get_scratch(&sc));
    cudaStream_t st = (cudaStream_t)stream;
    const int rows = rows_for(nx, ny);
    for (int colour = 0; colour < 2; ++colour) {
        RbgsArgs a = {};
        a.u = u; a.rhs = f; a.nx = nx; a.ny = ny; a.rows = rows; a.colour = colour; a.h = h; a.c = c; a.want_norm = 1;
        a.partials = sc->partials; a.ticket = sc->ticket; a.sumsq_out = sc->result + 2;
        dim3 g(((nx + 1) / 2 + kMGBX - 1) / kMGBX, (ny + rows - 1) / rows, 1);
        mg_rbgs_kernel<<<g, kMGBX, 0, st>>>(a);
        B2S_CUDA(cudaGetLastError());
    }
    B2S_CUDA(cudaMemcpyAsync(sc->pinned, sc->result + 2, 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
    B2S_CUDA(cudaStreamSynchronize(st));
    if (r_rms_host) *r_rms_host = sqrt((sc->pinned[0] + sc->pinned[1]) / ((double)nx * ny));
    return B2S_OK;
}

static int restrict_common(const double *fine, double *coarse, int nx, int ny, int apply_bcs, int fw, void *stream)
{
    B2S_REQUIRE(fine && coarse && nx >= 3 && ny >= 3, B2S_ERR_BAD_ARG, "bad argument");
    B2S_REQUIRE(((nx - 1) & 1) == 0 && ((ny - 1) & 1) == 0, B2S_ERR_BAD_SIZE, "ERROR:not a power of 2 (%dx%d)", nx, ny);
    RestrictArgs r = {};
    r.u = fine; r.coarse = coarse; r.nx = nx; r.ny = ny; r.nxc = 1 + (nx - 1) / 2; r.nyc = 1 + (ny - 1) / 2;
    r.h = 1.0; r.from_res = 0; r.full_weighting = fw; r.apply_bcs = apply_bcs;
    dim3 g((r.nxc + 63) / 64, (r.nyc + 3) / 4, 1);
    mg_restrict_kernel<<<g, dim3(64, 4, 1), 0, (cudaStream_t)stream>>>(r);
    B2S_CUDA(cudaGetLastError());
    return B2S_OK;
}
int b2s_restrict_inject2d(const double *fine, double *coarse, int nx, int ny, int apply_bcs, void *stream)
{
    return restrict_common(fine, coarse, nx, ny, apply_bcs, 0, stream);
}
int b2s_restrict_fw2d(const double *fine, double *coarse, int nx, int ny, int apply_bcs, void *stream)
{
    return restrict_common(fine, coarse, nx, ny, apply_bcs, 1, stream);
}

int b2s_prolongate2d(const double *coarse, double *fine, int nx, int ny, int apply_bcs, void *stream)
{
    B2S_REQUIRE(fine && coarse && nx >= 3 && ny >= 3, B2S_ERR_BAD_ARG, "bad argument");
    B2S_REQUIRE(((nx - 1) & 1) == 0 && ((ny - 1) & 1) == 0, B2S_ERR_BAD_SIZE, "ERROR:not a power of 2 (%dx%d)", nx, ny);
    ProlongArgs p = {};
    p.coarse = coarse; p.fine = fine; p.nx = nx; p.ny = ny; p.nxc = 1 + (nx - 1) / 2; p.nyc = 1 + (ny - 1) / 2;
    p.rows = rows_for(nx, ny); p.mode = 0; p.apply_bcs = apply_bcs;
    mg_prolong_kernel<<<sweep_grid(nx, ny, p.rows), kMGBX, 0, (cudaStream_t)stream>>>(p);
    B2S_CUDA(cudaGetLastError());
    return B2S_OK;
}

int b2s_matvec2d(const double *T, double hx, double hy, double c, double *out, int nx, int ny, int policy, void *stream)
{
    B2S_REQUIRE(T && out && nx >= 3 && ny >= 3, B2S_ERR_BAD_ARG, "bad argument");
    B2S_CHECK(check_policy(policy));
    const int rows = rows_for(nx, ny);
    mg_matvec_kernel<<<sweep_grid(nx, ny, rows), kMGBX, 0, (cudaStream_t)stream>>>(T, hx, hy, c, out, nx, ny, rows);
    B2S_CUDA(cudaGetLastError());
    return B2S_OK;
}

int b2s_apply_bc2d(double *T, int nx, int ny, int kind, void *stream)
{
    B2S_REQUIRE(T && nx >= 3 && ny >= 3 && kind >= 0 && kind <= 2, B2S_ERR_BAD_ARG, "bad argument");
    const int t = nx + ny;
    mg_bc_kernel<<<(t + 255) / 256, 256, 0, (cudaStream_t)stream>>>(nullptr, T, nx, ny, kind);
    B2S_CUDA(cudaGetLastError());
    return B2S_OK;
}

static int reduce_common(const double *x, const double *y, size_t n, double *out_host, void *stream)
{
    B2S_REQUIRE(x && y && out_host, B2S_ERR_BAD_ARG, "NULL argument");
    Scratch *sc = nullptr;
    B2S_CHECK(get_scratch(&sc));
    cudaStream_t st = (cudaStream_t)stream;
    mg_reduce_kernel<<<kReduceBlocks, kReduceThreads, 0, st>>>(x, y, n, sc->partials, sc->ticket, sc->result, nullptr, 0);
    B2S_CUDA(cudaGetLastError());
    B2S_CUDA(cudaMemcpyAsync(sc->pinned, sc->result, sizeof(double), cudaMemcpyDeviceToHost, st));
    B2S_CUDA(cudaStreamSynchronize(st));
    *out_host = sc->pinned[0];
    return B2S_OK;
}
int b2s_dot(const double *x, const double *y, size_t n, double *out_host, void *stream) { return reduce_common(x, y, n, out_host, stream); }
int b2s_sumsq(const double *x, size_t n, double *out_host, void *stream) { return reduce_common(x, x, n, out_host, stream); }

int b2s_axpy(double alpha, const double *x, double *y, size_t n, void *stream)
{
    B2S_REQUIRE(x && y, B2S_ERR_BAD_ARG, "NULL argument");
    mg_axpy_kernel<<<296, 256, 0, (cudaStream_t)stream>>>(alpha, x, y, n, 0);
    B2S_CUDA(cudaGetLastError());
    return B2S_OK;
}
int b2s_xpby(const double *x, double beta, double *y, size_t n, void *stream)
{
    B2S_REQUIRE(x && y, B2S_ERR_BAD_ARG, "NULL argument");
    mg_axpy_kernel<<<296, 256, 0, (cudaStream_t)stream>>>(beta, x, y, n, 1);
    B2S_CUDA(cudaGetLastError());
    return B2S_OK;
}

int b2s_cg_solve(double *x, const double *b, double hx, double hy, double c, double tol, int nmax, int nx, int ny, int policy,
                 double *res_rms, int *iters, void *stream)
{
    B2S_REQUIRE(x && b && nx >= 3 && ny >= 3, B2S_ERR_BAD_ARG, "bad argument");
    B2S_CHECK(check_policy(policy));
    const size_t n = (size_t)nx * ny;
    const size_t smem = 6 * n * sizeof(double);
    cudaStream_t st = (cudaStream_t)stream;
    if (smem > 220 * 1024) {  // larger than shared memory: host-driven global-memory CG (3 syncs per iteration, like the reference)
        double *work = nullptr;
        B2S_CUDA(cudaMalloc(&work, 4 * n * sizeof(double)));
        double ss = 0.0;
        int it = 0;
        const int rc = cg_global(x, b, work, hx, hy, c, tol, nmax, nx, ny, st, &ss, &it, nullptr);
        cudaFree(work);
        if (rc != B2S_OK) return rc;
        if (res_rms) *res_rms = sqrt(ss / ((double)nx * ny));
        if (iters) *iters = it;
        return B2S_OK;
    }
    Scratch *sc = nullptr;
    B2S_CHECK(get_scratch(&sc));
    static bool attr = false;
    if (!attr) {
        B2S_CUDA(cudaFuncSetAttribute(mg_cg_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        attr = true;
    }
    int *it_dev = (int *)(sc->ticket + 2);
    mg_cg_smem_kernel<<<1, 1024, smem, st>>>(x, b, hx, hy, c, tol, nmax, nx, ny, sc->result, it_dev);
    B2S_CUDA(cudaGetLastError());
    B2S_CUDA(cudaMemcpyAsync(sc->pinned, sc->result, sizeof(double), cudaMemcpyDeviceToHost, st));
    B2S_CUDA(cudaMemcpyAsync(sc->pinned + 1, it_dev, sizeof(int), cudaMemcpyDeviceToHost, st));
    B2S_CUDA(cudaStreamSynchronize(st));
    if (res_rms) *res_rms = sqrt(sc->pinned[0] / ((double)nx * ny));
    if (iters) *iters = *(int *)(sc->pinned + 1);
    return B2S_OK;
}

}  // extern "C"
