// diffusion3d.cu -- host side of hot path 1 (3-D pseudo-transient diffusion): kernel launchers, the L0 entry point
// and the L1 solver handle (device-resident PT loop, z-slab decomposition over GPUs with the halo exchange and the
// norm all-reduce fused into the step kernel).
//
// Reference entry point mirrored: diffusion_3D_kernel_programming, scripts-part1/part1_kernel_programming.jl:99-228.
#include "diffusion3d_kernels.cuh"

#include <math.h>
#include <stdlib.h>

#include <algorithm>
#include <vector>

namespace b2s {

// ---- kernel launchers ----------------------------------------------------------------------------------------
struct TmaChoice {
    int tx, ty, stages;
};
static const TmaChoice kTmaChoices[] = {{128, 8, 4}, {64, 8, 4}, {64, 16, 4}, {128, 4, 4}, {128, 8, 3}, {256, 4, 4}, {64, 8, 6}};
static const int kNumTmaChoices = (int)(sizeof(kTmaChoices) / sizeof(kTmaChoices[0]));

static int env_int(const char *name, int dflt)
{
    const char *s = getenv(name);
    return (s && *s) ? atoi(s) : dflt;
}

// Tile shape by problem size (measured on B200, profiles/r01_small_grid_configs.log): small, L2-resident grids are
// latency-bound and prefer many small blocks; 512^3 streams from HBM and prefers 128x8 tiles.
static int tma_choice_index(size_t cells = 0)
{
    const char *s = getenv("B2S_TMA_CFG");
    if (s && *s) {
        const int c = atoi(s);
        return (c >= 0 && c < kNumTmaChoices) ? c : 0;
    }
    if (cells > 0 && cells <= (size_t)192 * 192 * 192) return 1;  // 64 x 8
    if (cells > 0 && cells <= (size_t)384 * 384 * 384) return 3;  // 128 x 4
    return 0;                                                      // 128 x 8
}

template <int TX, int TY, int S>
static int launch_tma_t(const CUtensorMap &mA, const CUtensorMap &mH, const StepParams &p, cudaStream_t st)
{
    using C = TmaCfg<TX, TY>;
    static bool attr_set[64] = {};
    int dev = 0;
    B2S_CUDA(cudaGetDevice(&dev));
    const size_t smem = C::smem_bytes(S);
    if (!attr_set[dev & 63]) {
        B2S_CUDA(cudaFuncSetAttribute(step_tma_kernel<TX, TY, S, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        B2S_CUDA(cudaFuncSetAttribute(step_tma_kernel<TX, TY, S, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set[dev & 63] = true;
    }
    dim3 block(TX / 2, TY, 1);
    dim3 grid((p.nx + TX - 1) / TX, (p.ny + TY - 1) / TY, (p.nz - 2 + p.zchunk - 1) / p.zchunk);
    static const bool force_multi = env_int("B2S_FORCE_MULTI_KERNEL", 0) != 0;  // A/B: the exchange-capable instantiation on one slab
    const bool multi = force_multi || p.flagged || p.push_lo != nullptr || p.push_hi != nullptr || p.peer_slots != nullptr;
    if (multi) step_tma_kernel<TX, TY, S, true><<<grid, block, smem, st>>>(mA, mH, p);
    else step_tma_kernel<TX, TY, S, false><<<grid, block, smem, st>>>(mA, mH, p);
    B2S_CUDA(cudaGetLastError());
    return B2S_OK;
}

static int launch_tma(int choice, const CUtensorMap &mA, const CUtensorMap &mH, const StepParams &p, cudaStream_t st)
{
    switch (choice) {
    case 0: return launch_tma_t<128, 8, 4>(mA, mH, p, st);
    case 1: return launch_tma_t<64, 8, 4>(mA, mH, p, st);
    case 2: return launch_tma_t<64, 16, 4>(mA, mH, p, st);
    case 3: return launch_tma_t<128, 4, 4>(mA, mH, p, st);
    case 4: return launch_tma_t<128, 8, 3>(mA, mH, p, st);
    case 5: return launch_tma_t<256, 4, 4>(mA, mH, p, st);
    case 6: return launch_tma_t<64, 8, 6>(mA, mH, p, st);
    }
    set_error("bad TMA configuration index %d", choice);
    return B2S_ERR_BAD_ARG;
}

static int grid_blocks_tma(int choice, int nx, int ny, int nz, int zchunk)
{
    const TmaChoice &c = kTmaChoices[choice];
    return ((nx + c.tx - 1) / c.tx) * ((ny + c.ty - 1) / c.ty) * ((nz - 2 + zchunk - 1) / zchunk);
}

static int launch_direct(const StepParams &p, cudaStream_t st, const ArrayArith &aa = ArrayArith())
{
    dim3 block(kDirBX, kDirBY, 1);
    dim3 grid((p.nx + kDirBX - 1) / kDirBX, (p.ny + kDirBY - 1) / kDirBY, (p.nz - 2 + p.zchunk - 1) / p.zchunk);
    if (p.array_arith) step_direct_kernel<true><<<grid, block, 0, st>>>(p, aa);
    else step_direct_kernel<false><<<grid, block, 0, st>>>(p, aa);
    B2S_CUDA(cudaGetLastError());
    return B2S_OK;
}

static bool tma_eligible(int nx, int ny, int nz, const void *a, const void *b, const void *c)
{
    return (nx % 2 == 0) && nx >= 64 && ny >= 16 && nz >= 8 && (((uintptr_t)a | (uintptr_t)b | (uintptr_t)c) & 15) == 0;
}

// z-chunk: enough blocks for several waves over 148 SMs while keeping the two extra planes per chunk cheap.
// Returns 0 when even one chunk per xy tile exceeds max_blocks (the block partials would not fit).
static int pick_zchunk(int nxy_tiles, int nz, int max_blocks)
{
    if (nxy_tiles > max_blocks) return 0;
    int zc = env_int("B2S_ZCHUNK", 0);
    const int interior = nz - 2;
    if (zc <= 0 && (long long)nxy_tiles * interior <= 6000) {
        // L2-resident grids (up to ~128^3) are latency-bound: a block's time is its chain of dependent planes, so short
        // chunks win as long as there are not many more blocks than ~700 (measured on B200, profiles/r02_small_grid_zchunk.jsonl:
        // 32^3 10.9 -> 6.3 us, 64^3 8.5 -> 6.5 us, 128^3 13.3 -> 12.8 us per iteration)
        const int chunks = std::max(1, (700 + nxy_tiles - 1) / nxy_tiles);
        zc = std::max(2, (interior + chunks - 1) / chunks);
    }
    if (zc <= 0) {
        const int want_blocks = 148 * 8;
        int chunks = (want_blocks + nxy_tiles - 1) / nxy_tiles;
        chunks = std::max(chunks, (interior + 32) / 64);  // measured on B200: ~64 planes per chunk is the sweet spot at 512^3
        chunks = std::max(1, std::min(chunks, interior / 8 > 0 ? interior / 8 : 1));
        zc = (interior + chunks - 1) / chunks;
    }
    zc = std::max(1, std::min(zc, interior));
    while (zc < interior && (long long)nxy_tiles * ((interior + zc - 1) / zc) > max_blocks) ++zc;
    return zc;
}

static int make_maps(int choice, const double *A, const double *Ht, int nx, int ny, int nz, CUtensorMap *mA, CUtensorMap *mH)
{
    const TmaChoice &c = kTmaChoices[choice];
    if (mA) B2S_CHECK(make_tensor_map_3d(mA, A, nx, ny, nz, c.tx + 4, c.ty + 2, 1));
    if (mH) B2S_CHECK(make_tensor_map_3d(mH, Ht, nx, ny, nz, c.tx, c.ty, 1));
    return B2S_OK;
}

}  // namespace b2s

using namespace b2s;

// ==================================================================================================================
// L0
// ==================================================================================================================
extern "C" int b2s_diffusion3d_step_tau(const double *Ht, const double *Htau, double *Htau2, double *dHdtau, int nx, int ny,
                                        int nz, double dtau, double _dt, double _dx, double _dy, double _dz, double D_dx,
                                        double D_dy, double D_dz, double norm_scale, double *sumsq_dev, int kernel_variant,
                                        void *stream)
{
    B2S_REQUIRE(Ht && Htau && Htau2, B2S_ERR_BAD_ARG, "NULL field pointer");
    B2S_REQUIRE(nx >= 3 && ny >= 3 && nz >= 3, B2S_ERR_BAD_SIZE, "grid %dx%dx%d has no interior", nx, ny, nz);
    Scratch *sc = nullptr;
    B2S_CHECK(get_scratch(&sc));
    StepParams p = {};
    p.Ht = Ht; p.A = Htau; p.B = Htau2; p.R = dHdtau;
    p.nx = nx; p.ny = ny; p.nz = nz;
    p.dtau = dtau; p._dt = _dt; p._dx = _dx; p._dy = _dy; p._dz = _dz;
    p.mD_dx = -D_dx; p.mD_dy = -D_dy; p.mD_dz = -D_dz;
    p.norm_scale = norm_scale;
    p.partials = sc->partials; p.ticket = sc->ticket;
    p.sumsq_out = sumsq_dev;
    cudaStream_t st = (cudaStream_t)stream;
    bool use_tma = kernel_variant == B2S_KERNEL_TMA ||
                   (kernel_variant == B2S_KERNEL_AUTO && tma_eligible(nx, ny, nz, Ht, Htau, Htau2) &&
                    (size_t)nx * ny * nz >= (size_t)64 * 64 * 64);
    if (use_tma) {
        B2S_REQUIRE(nx % 2 == 0 && (((uintptr_t)Ht | (uintptr_t)Htau | (uintptr_t)Htau2 | (uintptr_t)dHdtau) & 15) == 0,
                    B2S_ERR_BAD_ARG, "TMA variant needs even nx and 16-byte aligned fields");
        const int ch = tma_choice_index((size_t)nx * ny * nz);
        const TmaChoice &c = kTmaChoices[ch];
        p.zchunk = pick_zchunk(((nx + c.tx - 1) / c.tx) * ((ny + c.ty - 1) / c.ty), nz, kMaxPartials);
        B2S_REQUIRE(p.zchunk > 0, B2S_ERR_BAD_SIZE, "grid %dx%dx%d has more xy tiles than the %d block partials", nx, ny, nz, kMaxPartials);
        CUtensorMap mA, mH;
        B2S_CHECK(make_maps(ch, Htau, Ht, nx, ny, nz, &mA, &mH));
        return launch_tma(ch, mA, mH, p, st);
    }
    B2S_REQUIRE(kernel_variant == B2S_KERNEL_AUTO || kernel_variant == B2S_KERNEL_DIRECT, B2S_ERR_BAD_ARG,
                "unknown kernel variant %d", kernel_variant);
    p.zchunk = pick_zchunk(((nx + kDirBX - 1) / kDirBX) * ((ny + kDirBY - 1) / kDirBY), nz, kMaxPartials);
    B2S_REQUIRE(p.zchunk > 0, B2S_ERR_BAD_SIZE, "grid %dx%dx%d has more xy tiles than the %d block partials", nx, ny, nz, kMaxPartials);
    return launch_direct(p, st);
}

// ==================================================================================================================
// L1 handle
// ==================================================================================================================
namespace {

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct Arena {  // layout of the per-slab device allocation (identical on every rank -> usable through IPC)
    size_t cells, off_ht, off_buf[2], off_slots, off_state, off_partials, off_ticket, off_peer_table, off_cart, off_cart_table,
        off_flags, flag_tiles, bytes;
    void layout(size_t ncells, size_t max_tiles)
    {
        cells = ncells;
        flag_tiles = max_tiles;
        size_t o = 0;
        const size_t fb = align_up(ncells * sizeof(double), 1024);
        off_buf[0] = o; o += fb;
        off_buf[1] = o; o += fb;
        off_ht = o; o += fb;
        off_slots = o; o += align_up(sizeof(RankSlots), 1024);
        off_state = o; o += 1024;
        off_partials = o; o += align_up(sizeof(double) * kMaxPartials, 1024);
        off_ticket = o; o += 1024;
        off_peer_table = o; o += align_up(sizeof(void *) * kMaxRanks, 1024);
        off_cart = o; o += align_up(sizeof(CartSync), 1024);
        off_cart_table = o; o += align_up(sizeof(void *) * kMaxRanks, 1024);
        off_flags = o; o += align_up(2 * max_tiles * sizeof(unsigned long long), 1024);  // halo flags [2][max_tiles]
        bytes = o;
    }
};

struct Slab {
    int rank = 0;  // global slab index
    int dev = 0;
    int devslot = 0;  // index into DeviceCtx
    char *arena = nullptr;
    double *Ht = nullptr, *buf[2] = {nullptr, nullptr};
    RankSlots *slots = nullptr;
    double *partials = nullptr;
    unsigned int *ticket = nullptr;
    CUtensorMap mapHt, mapBuf[2];
    // neighbours' buffers (local pointer, peer pointer or IPC mapping); nullptr at the ends of the slab stack
    double *lo_buf[2] = {nullptr, nullptr}, *hi_buf[2] = {nullptr, nullptr};
    PTState *state = nullptr;                      // this slab's own PT state (z-slab stacks; cart handles share one per device)
    unsigned long long *flags = nullptr;           // my halo flags [2][ntiles]
    unsigned long long *lo_flags = nullptr, *hi_flags = nullptr;  // the z neighbours' flag arrays (local, peer or IPC pointers)
    RankSlots **table = nullptr;                   // device array: every rank's mailbox (z-slab stacks)
    double *staging = nullptr;  // next job's state (b2s_diff3d_upload_state_async), allocated on first use
    double *stage_out = nullptr;  // device-side copy of a result on its way to the host (b2s_diff3d_download_state_async)
    bool staged = false, staging_used = false;
};

struct DeviceCtx {  // one per distinct device of the handle: stream, PT state, rank-slot mailbox
    int dev = 0;
    cudaStream_t stream = nullptr;
    PTState *state = nullptr;      // in the arena of the first slab on this device
    RankSlots *slots = nullptr;    // ditto
    RankSlots **peer_table = nullptr;  // device array: every RankSlots instance that must receive this device's partials
    double *err_hist = nullptr;
    int err_hist_cap = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaEvent_t ev_bar = nullptr;  // cross-device ordering of the plane copies (general decompositions)
    cudaStream_t copy_stream = nullptr;          // uploads (b2s_diff3d_upload_state_async)
    cudaStream_t copy_stream_down = nullptr;     // downloads (b2s_diff3d_download_state_async): PCIe is full duplex, so the two
                                                 // directions get a stream each and overlap
    cudaEvent_t ev_work = nullptr, ev_down = nullptr, ev_up = nullptr, ev_staged_free = nullptr;
    bool download_pending = false;
};

}  // namespace

struct b2s_diff3d {
    b2s_diff3d_config cfg;
    b2s_diff3d_params prm;
    double _dt, _dx, _dy, _dz, D_dx, D_dy, D_dz;
    Arena ar;
    std::vector<int> devices;
    std::vector<Slab> slabs;
    std::vector<DeviceCtx> devs;
    bool use_tma = false;
    int tma_choice = 0;
    int zchunk = 1;
    int nblocks = 0;
    int nslots_dst = 0;           // RankSlots instances that receive partials (devices in-process, ranks multi-process)
    bool multi = false;           // more than one slab in the global stack
    bool cart = false;            // decomposition in x or y as well: update_halo! as separate plane copies
    bool zstack = false;          // multi && !cart: fused halo push with neighbour flags, lagged norm evaluation
    bool cart_devbarrier = false; // in-process general decomposition with one rank per device: device-side rank barriers
    int ntiles = 0;               // xy tiles per launch (= halo flags per array)
    int skip_push = 0;            // host mirror of PTState::skip_push
    std::vector<char *> peer_base; // one process per GPU: every rank's arena (own or IPC-mapped), in rank order
    bool connected = false;       // multi-process: peers mapped
    std::vector<void *> ipc_mapped;
    long long launched = 0;       // PT iterations since create (host mirror of PTState::total_iters)
    unsigned long long seq = 1;   // host mirror of PTState::seq
    long long kernel_launches = 0;
    double last_ms = 0.0;
    PTState *pinned = nullptr;    // host staging of the state: [0] upload, [1], [2] polled snapshots
    cudaEvent_t ev_poll[2] = {nullptr, nullptr};
    // small (L2-resident) grids on one device: a whole batch of PT iterations is one CUDA graph per ping-pong parity (a
    // kernel boundary inside a graph costs ~1.4 us, a stream launch ~2.3 us of device-side gap: at 32^3..128^3 that is
    // 15-50 % of an iteration). The kernels' arguments never change, so the graphs are built once per handle.
    bool use_graph = false, warmed = false;
    int graph_batch = 0;
    cudaGraphExec_t graph[2] = {nullptr, nullptr};
    long long graph_launches_per_batch = 0;
    int it_step = 0;              // iter_inner of the time step in progress
};

namespace {

struct IpcBlob {
    cudaIpcMemHandle_t handle;
    int rank, dev;
    unsigned long long arena_bytes;
    int pid_lo;
    int pad;
};

int slab_index(const b2s_diff3d *h, int global_slab)
{
    const int i = global_slab - h->cfg.slab_begin;
    return (i >= 0 && i < (int)h->slabs.size()) ? i : -1;
}

// dims of the rank grid and the coordinates of a rank (MPI Cartesian order: z fastest)
void cart_dims(const b2s_diff3d_config &c, int dims[3])
{
    dims[0] = c.dimx > 1 ? c.dimx : 1;
    dims[1] = c.dimy > 1 ? c.dimy : 1;
    dims[2] = c.nslabs_total / (dims[0] * dims[1]);
}
void cart_coords(const b2s_diff3d_config &c, int rank, int coord[3])
{
    int dims[3];
    cart_dims(c, dims);
    coord[2] = rank % dims[2];
    coord[1] = (rank / dims[2]) % dims[1];
    coord[0] = rank / (dims[2] * dims[1]);
}
bool is_cart(const b2s_diff3d_config &c) { return (c.dimx > 1 ? c.dimx : 1) * (c.dimy > 1 ? c.dimy : 1) > 1; }

// every device stream waits for the work enqueued so far on all the others (no-op with one device)
int cross_device_barrier(b2s_diff3d *h)
{
    if (h->devs.size() < 2) return B2S_OK;
    for (DeviceCtx &d : h->devs) {
        B2S_CUDA(cudaSetDevice(d.dev));
        B2S_CUDA(cudaEventRecord(d.ev_bar, d.stream));
    }
    for (DeviceCtx &d : h->devs) {
        B2S_CUDA(cudaSetDevice(d.dev));
        for (DeviceCtx &o : h->devs)
            if (&o != &d) B2S_CUDA(cudaStreamWaitEvent(d.stream, o.ev_bar, 0));
    }
    return B2S_OK;
}

// update_halo!(buf[which]) of ImplicitGlobalGrid for a general decomposition: per axis x, y, z the low rank's plane n-2
// goes to the high rank's plane 0 and the high rank's plane 1 to the low rank's plane n-1 (whole planes, overlap 2).
// The reference exchanges the buffer the step kernel has just READ (part1_kernel_programming.jl:182,187: halos lag two
// iterations) -- `which` selects it; the consistent mode passes the buffer just written.
// The same for one process per GPU: this rank PULLS the two planes of every split axis from its neighbours' arenas
// (CUDA IPC mappings over NVLink); the phases are separated by cart_barrier_kernel instead of events.
int cart_update_halo_multiprocess(b2s_diff3d *h, int which)
{
    const b2s_diff3d_config &c = h->cfg;
    int dims[3], coord[3];
    cart_dims(c, dims);
    Slab &me = h->slabs[0];
    DeviceCtx &d = h->devs[me.devslot];
    cart_coords(c, me.rank, coord);
    const int nd[3] = {c.nx, c.ny, c.nz};
    CartSync *mine = (CartSync *)(me.arena + h->ar.off_cart);
    CartSync *const *table = (CartSync *const *)(me.arena + h->ar.off_cart_table);
    int k = 1;
    auto barrier = [&]() {
        cart_barrier_kernel<<<1, kMaxRanks, 0, d.stream>>>(d.state, mine, table, c.nslabs_total, me.rank, k++, (long long)20e9);
        h->kernel_launches += 1;
    };
    auto peer_buf = [&](const int cc[3]) {
        const int r = (cc[0] * dims[1] + cc[1]) * dims[2] + cc[2];
        return (double *)(h->peer_base[(size_t)r] + h->ar.off_buf[which]);
    };
    barrier();  // every rank's step kernel is done before anyone's halo cells change
    for (int axis = 0; axis < 3; ++axis) {
        if (dims[axis] == 1) continue;
        const size_t plane = (size_t)nd[(axis + 1) % 3] * nd[(axis + 2) % 3];
        const int blocks = (int)std::min<size_t>((plane + 255) / 256, 592);
        const int n = nd[axis];
        if (coord[axis] > 0) {  // low neighbour's plane n-2 -> my plane 0
            int cc[3] = {coord[0], coord[1], coord[2]};
            cc[axis] -= 1;
            halo_plane_copy_kernel<<<blocks, 256, 0, d.stream>>>(peer_buf(cc), me.buf[which], axis, n - 2, 0, c.nx, c.ny, c.nz, d.state);
            h->kernel_launches += 1;
        }
        if (coord[axis] + 1 < dims[axis]) {  // high neighbour's plane 1 -> my plane n-1
            int cc[3] = {coord[0], coord[1], coord[2]};
            cc[axis] += 1;
            halo_plane_copy_kernel<<<blocks, 256, 0, d.stream>>>(peer_buf(cc), me.buf[which], axis, 1, n - 1, c.nx, c.ny, c.nz, d.state);
            h->kernel_launches += 1;
        }
        barrier();  // the next axis (or the next step kernel) touches cells this one has read or written
    }
    B2S_CUDA(cudaGetLastError());
    return B2S_OK;
}

// In-process handle whose ranks all sit on DIFFERENT devices: the same pull + device-side rank barrier as with one process
// per GPU (every device runs its own stream; the barrier kernels of different devices run concurrently, so they may wait
// on one another). The event-based version below costs 4 x (N records + N(N-1) stream waits) host calls and as many
// cross-device event latencies per iteration (measured: 584 us per iteration for the 2x2x2 grid on 8 GPUs).
int cart_update_halo_devbarrier(b2s_diff3d *h, int which)
{
    const b2s_diff3d_config &c = h->cfg;
    int dims[3];
    cart_dims(c, dims);
    const int nd[3] = {c.nx, c.ny, c.nz};
    int k = 1;
    auto barrier_all = [&]() -> int {
        for (Slab &me : h->slabs) {
            DeviceCtx &d = h->devs[me.devslot];
            B2S_CUDA(cudaSetDevice(me.dev));
            cart_barrier_kernel<<<1, kMaxRanks, 0, d.stream>>>(d.state, (CartSync *)(me.arena + h->ar.off_cart),
                                                               (CartSync *const *)(me.arena + h->ar.off_cart_table), c.nslabs_total,
                                                               me.rank, k, (long long)20e9);
            h->kernel_launches += 1;
        }
        ++k;
        return B2S_OK;
    };
    B2S_CHECK(barrier_all());  // every rank's step kernel is done before anyone's halo cells change
    for (int axis = 0; axis < 3; ++axis) {
        if (dims[axis] == 1) continue;
        const size_t plane = (size_t)nd[(axis + 1) % 3] * nd[(axis + 2) % 3];
        const int blocks = (int)std::min<size_t>((plane + 255) / 256, 592);
        const int n = nd[axis];
        for (Slab &me : h->slabs) {
            DeviceCtx &d = h->devs[me.devslot];
            int coord[3];
            cart_coords(c, me.rank, coord);
            B2S_CUDA(cudaSetDevice(me.dev));
            auto nb_buf = [&](int delta) {
                int cc[3] = {coord[0], coord[1], coord[2]};
                cc[axis] += delta;
                return h->slabs[(size_t)((cc[0] * dims[1] + cc[1]) * dims[2] + cc[2])].buf[which];
            };
            if (coord[axis] > 0) {  // low neighbour's plane n-2 -> my plane 0
                halo_plane_copy_kernel<<<blocks, 256, 0, d.stream>>>(nb_buf(-1), me.buf[which], axis, n - 2, 0, c.nx, c.ny, c.nz, d.state);
                h->kernel_launches += 1;
            }
            if (coord[axis] + 1 < dims[axis]) {  // high neighbour's plane 1 -> my plane n-1
                halo_plane_copy_kernel<<<blocks, 256, 0, d.stream>>>(nb_buf(+1), me.buf[which], axis, 1, n - 1, c.nx, c.ny, c.nz, d.state);
                h->kernel_launches += 1;
            }
        }
        B2S_CHECK(barrier_all());  // the next axis (or the next step kernel) touches cells this one has read or written
    }
    B2S_CUDA(cudaGetLastError());
    return B2S_OK;
}

int cart_update_halo(b2s_diff3d *h, int which)
{
    if (h->slabs.size() == 1 && h->multi) return cart_update_halo_multiprocess(h, which);
    if (h->cart_devbarrier) return cart_update_halo_devbarrier(h, which);
    const b2s_diff3d_config &c = h->cfg;
    int dims[3];
    cart_dims(c, dims);
    const int nd[3] = {c.nx, c.ny, c.nz};
    B2S_CHECK(cross_device_barrier(h));  // all step kernels done before anyone's halo cells change
    for (int axis = 0; axis < 3; ++axis) {
        if (dims[axis] == 1) continue;
        const size_t plane = (size_t)nd[(axis + 1) % 3] * nd[(axis + 2) % 3];
        const int blocks = (int)std::min<size_t>((plane + 255) / 256, 592);
        for (size_t i = 0; i < h->slabs.size(); ++i) {
            Slab &lo = h->slabs[i];
            int coord[3];
            cart_coords(c, lo.rank, coord);
            if (coord[axis] + 1 >= dims[axis]) continue;
            int ch[3] = {coord[0], coord[1], coord[2]};
            ch[axis] += 1;
            Slab &hi = h->slabs[(size_t)((ch[0] * dims[1] + ch[1]) * dims[2] + ch[2])];
            const int n = nd[axis];
            {  // low rank's plane n-2 -> high rank's plane 0 (on the receiver's stream)
                DeviceCtx &d = h->devs[hi.devslot];
                B2S_CUDA(cudaSetDevice(hi.dev));
                halo_plane_copy_kernel<<<blocks, 256, 0, d.stream>>>(lo.buf[which], hi.buf[which], axis, n - 2, 0, c.nx, c.ny, c.nz, d.state);
            }
            {  // high rank's plane 1 -> low rank's plane n-1
                DeviceCtx &d = h->devs[lo.devslot];
                B2S_CUDA(cudaSetDevice(lo.dev));
                halo_plane_copy_kernel<<<blocks, 256, 0, d.stream>>>(hi.buf[which], lo.buf[which], axis, 1, n - 1, c.nx, c.ny, c.nz, d.state);
            }
            h->kernel_launches += 2;
        }
        B2S_CUDA(cudaGetLastError());
        B2S_CHECK(cross_device_barrier(h));  // the next axis forwards cells this one has just written
    }
    return B2S_OK;
}

int launch_iteration(b2s_diff3d *h)
{
    const int par = (int)(h->launched & 1);
    for (Slab &s : h->slabs) {
        DeviceCtx &d = h->devs[s.devslot];
        B2S_CUDA(cudaSetDevice(s.dev));
        StepParams p = {};
        p.Ht = s.Ht; p.A = s.buf[par]; p.B = s.buf[par ^ 1]; p.R = nullptr;
        p.nx = h->cfg.nx; p.ny = h->cfg.ny; p.nz = h->cfg.nz;
        p.dtau = h->prm.dtau; p._dt = h->_dt; p._dx = h->_dx; p._dy = h->_dy; p._dz = h->_dz;
        p.mD_dx = -h->D_dx; p.mD_dy = -h->D_dy; p.mD_dz = -h->D_dz;
        p.norm_scale = h->prm.dt;
        p.partials = s.partials; p.ticket = s.ticket;
        p.state = h->zstack ? s.state : d.state;
        p.err_hist = d.err_hist;
        p.zchunk = h->zchunk;
        p.consistent = h->cfg.halo_mode == B2S_HALO_CONSISTENT;
        const size_t plane = (size_t)p.nx * p.ny;
        if (s.lo_buf[par ^ 1]) p.push_lo = s.lo_buf[par ^ 1] + plane * (p.nz - 1);
        if (s.hi_buf[par ^ 1]) p.push_hi = s.hi_buf[par ^ 1];
        if (h->zstack) {
            p.peer_slots = s.table;
            p.nranks = h->cfg.nslabs_total;
            p.myrank = s.rank;
            p.flagged = 1;
            p.ntiles = h->ntiles;
            p.flags = s.flags; p.lo_flags = s.lo_flags; p.hi_flags = s.hi_flags;
            p.my_slots = s.slots;
            p.timeout_cycles = (long long)20e9;
        } else if (h->multi) {
            p.peer_slots = d.peer_table;
            p.nranks = h->nslots_dst;  // number of destinations
            p.myrank = s.rank;
        } else {
            p.fuse_finalize = 1;
        }
        p.array_arith = h->cfg.arithmetic == B2S_ARITH_ARRAY;
        if (h->use_tma) {
            B2S_CHECK(launch_tma(h->tma_choice, s.mapBuf[par], s.mapHt, p, d.stream));
        } else {
            ArrayArith aa = {1.0, h->prm.dx, h->prm.dy, h->prm.dz, h->prm.dt};
            B2S_CHECK(launch_direct(p, d.stream, aa));
        }
        h->kernel_launches += 1;
    }
    if (h->cart) B2S_CHECK(cart_update_halo(h, h->cfg.halo_mode == B2S_HALO_CONSISTENT ? (par ^ 1) : par));
    if (h->multi && !h->zstack) {
        for (DeviceCtx &d : h->devs) {
            B2S_CUDA(cudaSetDevice(d.dev));
            pt_finalize_kernel<<<1, kMaxRanks, 0, d.stream>>>(d.state, d.err_hist, nullptr, d.slots, h->cfg.nslabs_total,
                                                              (long long)20e9, 0);
            B2S_CUDA(cudaGetLastError());
            h->kernel_launches += 1;
        }
    }
    h->launched += 1;
    return B2S_OK;
}

int upload_state(b2s_diff3d *h, const PTState &st)
{
    *h->pinned = st;
    if (h->zstack) {  // every slab of a z stack has its own state
        for (Slab &s : h->slabs) {
            B2S_CUDA(cudaSetDevice(s.dev));
            B2S_CUDA(cudaMemcpyAsync(s.state, h->pinned, sizeof(PTState), cudaMemcpyHostToDevice, h->devs[s.devslot].stream));
        }
        return B2S_OK;
    }
    for (DeviceCtx &d : h->devs) {
        B2S_CUDA(cudaSetDevice(d.dev));
        B2S_CUDA(cudaMemcpyAsync(d.state, h->pinned, sizeof(PTState), cudaMemcpyHostToDevice, d.stream));
    }
    return B2S_OK;
}

// z-slab stacks: one tiny kernel per slab at the end of a host batch evaluates the iteration the step kernels left pending
int enqueue_lagged_finalize(b2s_diff3d *h)
{
    for (Slab &s : h->slabs) {
        DeviceCtx &d = h->devs[s.devslot];
        B2S_CUDA(cudaSetDevice(s.dev));
        pt_finalize_kernel<<<1, kMaxRanks, 0, d.stream>>>(s.state, d.err_hist, nullptr, s.slots, h->cfg.nslabs_total, (long long)20e9, 1);
        B2S_CUDA(cudaGetLastError());
        h->kernel_launches += 1;
    }
    return B2S_OK;
}

// Runs the device-resident loop until the state says done. Returns the final state of device 0 in *out.
// The host enqueues batches of iterations and polls one PTState per batch; the next batch is enqueued BEFORE the previous
// one is polled (two pinned snapshots, two events), so the GPU never idles on the host round trip. Kernels enqueued after
// the exit return at once.
int run_loop(b2s_diff3d *h, PTState st, PTState *out)
{
    B2S_REQUIRE(!h->multi || h->connected, B2S_ERR_STATE, "multi-process handle is not connected (b2s_diff3d_ipc_connect)");
    DeviceGuard guard;
    guard.set(h->devs[0].dev);
    st.total_iters = h->launched;
    st.seq = h->seq;
    st.error = 0;
    st.pending = 0;
    st.skip_push = h->skip_push;
    st.done = (st.it < st.iter_max) ? 0 : 1;
    B2S_CHECK(upload_state(h, st));
    for (DeviceCtx &d : h->devs) {
        B2S_CUDA(cudaSetDevice(d.dev));
        B2S_CUDA(cudaEventRecord(d.ev0, d.stream));
    }
    int batch = h->cfg.batch > 0 ? h->cfg.batch : 0;
    if (batch <= 0) {
        const double cells = (double)h->ar.cells;
        batch = cells >= 256.0 * 256 * 256 ? 16 : (cells >= 96.0 * 96 * 96 ? 64 : 128);
        if (!st.check) batch = 512;  // fixed count: nothing to poll for but errors
    }
    if (h->use_graph) batch = h->graph_batch;
    DeviceCtx &d0 = h->devs[0];
    PTState *src = h->zstack ? h->slabs[0].state : d0.state;
    PTState cur = st;
    int enqueued = 0, inflight = 0, slot = 0;
    auto submit = [&]() -> int {
        const int n = std::min(batch, st.iter_max - st.it - enqueued);
        if (n <= 0) return B2S_OK;
        if (h->use_graph && n == h->graph_batch && h->warmed) {  // (the very first batch runs as stream launches: it
                                                                   // sets the kernels' shared-memory attributes)
            const int par = (int)(h->launched & 1);
            if (!h->graph[par]) {  // capture n iterations starting at this parity (nothing executes during the capture)
                const long long launched0 = h->launched, kl0 = h->kernel_launches;
                cudaGraph_t g = nullptr;
                B2S_CUDA(cudaSetDevice(d0.dev));
                B2S_CUDA(cudaStreamBeginCapture(d0.stream, cudaStreamCaptureModeThreadLocal));
                int rc = B2S_OK;
                for (int i = 0; i < n && rc == B2S_OK; ++i) rc = launch_iteration(h);
                cudaError_t e = cudaStreamEndCapture(d0.stream, &g);
                h->graph_launches_per_batch = h->kernel_launches - kl0;
                h->launched = launched0; h->kernel_launches = kl0;
                if (rc != B2S_OK) { if (g) cudaGraphDestroy(g); return rc; }
                B2S_CUDA(e);
                B2S_CUDA(cudaGraphInstantiate(&h->graph[par], g, 0));
                B2S_CUDA(cudaGraphDestroy(g));
            }
            B2S_CUDA(cudaGraphLaunch(h->graph[par], d0.stream));
            h->launched += n;
            h->kernel_launches += h->graph_launches_per_batch;
        } else {
            for (int i = 0; i < n; ++i) B2S_CHECK(launch_iteration(h));
            h->warmed = true;
        }
        if (h->zstack) B2S_CHECK(enqueue_lagged_finalize(h));
        B2S_CUDA(cudaSetDevice(d0.dev));
        B2S_CUDA(cudaMemcpyAsync(h->pinned + 1 + slot, src, sizeof(PTState), cudaMemcpyDeviceToHost, d0.stream));
        B2S_CUDA(cudaEventRecord(h->ev_poll[slot], d0.stream));
        enqueued += n;
        inflight += 1;
        slot ^= 1;
        return B2S_OK;
    };
    if (!cur.done) B2S_CHECK(submit());
    while (inflight > 0) {
        if (inflight < 2) B2S_CHECK(submit());      // keep one batch queued behind the one being waited for
        const int oldest = inflight == 2 ? slot : slot ^ 1;
        B2S_CUDA(cudaEventSynchronize(h->ev_poll[oldest]));
        cur = h->pinned[1 + oldest];
        inflight -= 1;
        B2S_REQUIRE(!cur.error, B2S_ERR_CUDA, "timed out waiting for a peer GPU (halo flag or partial norm, iteration %d)", cur.it);
        if (cur.done) break;  // everything still queued is a no-op
    }
    h->launched = cur.total_iters;  // launches after the exit were no-ops (a speculative iteration is not counted)
    h->seq = cur.seq;
    h->skip_push = cur.skip_push;
    double ms_max = 0.0;
    for (DeviceCtx &d : h->devs) {
        B2S_CUDA(cudaSetDevice(d.dev));
        B2S_CUDA(cudaEventRecord(d.ev1, d.stream));
        B2S_CUDA(cudaEventSynchronize(d.ev1));
        float ms = 0.f;
        B2S_CUDA(cudaEventElapsedTime(&ms, d.ev0, d.ev1));
        ms_max = std::max(ms_max, (double)ms);
    }
    h->last_ms = ms_max;
    h->it_step = cur.it;
    if (out) *out = cur;
    return B2S_OK;
}

int ensure_hist(b2s_diff3d *h, int n)
{
    for (DeviceCtx &d : h->devs) {
        if (d.err_hist_cap >= n) continue;
        B2S_CUDA(cudaSetDevice(d.dev));
        if (d.err_hist) B2S_CUDA(cudaFree(d.err_hist));
        d.err_hist = nullptr;
        B2S_CUDA(cudaMalloc(&d.err_hist, sizeof(double) * (size_t)n));
        d.err_hist_cap = n;
        for (int i = 0; i < 2; ++i)  // the captured launches carry the old history pointer
            if (h->graph[i]) { cudaGraphExecDestroy(h->graph[i]); h->graph[i] = nullptr; }
    }
    return B2S_OK;
}

void compute_params_raw(const b2s_diff3d_config &c, b2s_diff3d_params &p)
{
    // part1_kernel_programming.jl:104-131 with dims = (dimx, dimy, nslabs_total / (dimx*dimy))
    const double D = 1.0;
    int dims[3];
    cart_dims(c, dims);
    p.lx = 10.0; p.ly = 10.0; p.lz = 10.0;
    if (c.scale_physical_size) { p.lx = dims[0] * 10.0; p.ly = dims[1] * 10.0; p.lz = dims[2] * 10.0; }
    p.nx_g = dims[0] * (c.nx - 2) + 2; p.ny_g = dims[1] * (c.ny - 2) + 2; p.nz_g = dims[2] * (c.nz - 2) + 2;
    p.dx = p.lx / p.nx_g; p.dy = p.ly / p.ny_g; p.dz = p.lz / p.nz_g;
    p.total_N = (double)c.nslabs_total * c.nx * c.ny * c.nz;
    p.dt = 0.2;
    const double m = fmin(p.dx, fmin(p.dy, p.dz));
    p.dtau = m * m / D / 8.1;
}

void compute_params(b2s_diff3d *h)
{
    compute_params_raw(h->cfg, h->prm);
    const b2s_diff3d_params &p = h->prm;
    const double D = 1.0;
    h->_dt = 1.0 / p.dt; h->_dx = 1.0 / p.dx; h->_dy = 1.0 / p.dy; h->_dz = 1.0 / p.dz;
    h->D_dx = D / p.dx; h->D_dy = D / p.dy; h->D_dz = D / p.dz;
}

int destroy_impl(b2s_diff3d *h)
{
    if (!h) return B2S_OK;
    for (void *m : h->ipc_mapped) cudaIpcCloseMemHandle(m);
    for (DeviceCtx &d : h->devs) {
        cudaSetDevice(d.dev);
        if (d.stream) cudaStreamSynchronize(d.stream);
        if (d.err_hist) cudaFree(d.err_hist);
        if (d.ev0) cudaEventDestroy(d.ev0);
        if (d.ev1) cudaEventDestroy(d.ev1);
        if (d.ev_bar) cudaEventDestroy(d.ev_bar);
        if (d.copy_stream) { cudaStreamSynchronize(d.copy_stream); cudaStreamDestroy(d.copy_stream); }
        if (d.copy_stream_down) { cudaStreamSynchronize(d.copy_stream_down); cudaStreamDestroy(d.copy_stream_down); }
        if (d.ev_work) cudaEventDestroy(d.ev_work);
        if (d.ev_down) cudaEventDestroy(d.ev_down);
        if (d.ev_up) cudaEventDestroy(d.ev_up);
        if (d.ev_staged_free) cudaEventDestroy(d.ev_staged_free);
        if (d.stream) cudaStreamDestroy(d.stream);
    }
    if (!h->devs.empty()) cudaSetDevice(h->devs[0].dev);
    for (int i = 0; i < 2; ++i) {
        if (h->ev_poll[i]) cudaEventDestroy(h->ev_poll[i]);
        if (h->graph[i]) cudaGraphExecDestroy(h->graph[i]);
    }
    for (Slab &s : h->slabs) {
        cudaSetDevice(s.dev);
        if (s.arena) cudaFree(s.arena);
        if (s.staging) cudaFree(s.staging);
        if (s.stage_out) cudaFree(s.stage_out);
    }
    if (h->pinned) cudaFreeHost(h->pinned);
    delete h;
    return B2S_OK;
}

}  // namespace

extern "C" {

int b2s_diff3d_create(b2s_diff3d **out, const b2s_diff3d_config *cfg)
{
    B2S_REQUIRE(out && cfg, B2S_ERR_BAD_ARG, "NULL argument");
    *out = nullptr;
    B2S_REQUIRE(cfg->nx >= 3 && cfg->ny >= 3 && cfg->nz >= 3, B2S_ERR_BAD_SIZE, "grid %dx%dx%d has no interior", cfg->nx,
                cfg->ny, cfg->nz);
    B2S_REQUIRE(cfg->nslabs_total >= 1 && cfg->nslabs_total <= kMaxRanks, B2S_ERR_BAD_ARG, "nslabs_total %d out of range",
                cfg->nslabs_total);
    B2S_REQUIRE(cfg->slab_count >= 1 && cfg->slab_begin >= 0 && cfg->slab_begin + cfg->slab_count <= cfg->nslabs_total,
                B2S_ERR_BAD_ARG, "bad slab range [%d,+%d) of %d", cfg->slab_begin, cfg->slab_count, cfg->nslabs_total);
    B2S_REQUIRE(cfg->slab_count == cfg->nslabs_total || cfg->slab_count == 1, B2S_ERR_BAD_ARG,
                "a handle hosts either all slabs (in-process) or exactly one (one process per GPU)");
    B2S_REQUIRE(cfg->halo_mode == B2S_HALO_REFERENCE_LAG2 || cfg->halo_mode == B2S_HALO_CONSISTENT, B2S_ERR_BAD_ARG,
                "bad halo_mode");
    B2S_REQUIRE(cfg->bc_mode == B2S_BC_LITERAL || cfg->bc_mode == B2S_BC_PROPER, B2S_ERR_BAD_ARG, "bad bc_mode");
    B2S_REQUIRE(cfg->arithmetic == B2S_ARITH_KERNEL || cfg->arithmetic == B2S_ARITH_ARRAY, B2S_ERR_BAD_ARG, "bad arithmetic");
    B2S_REQUIRE(cfg->arithmetic == B2S_ARITH_KERNEL || cfg->kernel_variant != B2S_KERNEL_TMA, B2S_ERR_BAD_ARG,
                "the array-programming arithmetic exists in the direct kernel only");
    {
        const int dx_ = cfg->dimx > 1 ? cfg->dimx : 1, dy_ = cfg->dimy > 1 ? cfg->dimy : 1;
        B2S_REQUIRE(cfg->dimx >= 0 && cfg->dimy >= 0 && cfg->nslabs_total % (dx_ * dy_) == 0, B2S_ERR_BAD_ARG,
                    "dims (%d, %d, .) do not divide %d ranks", cfg->dimx, cfg->dimy, cfg->nslabs_total);
    }
    int ndev = 0;
    B2S_CHECK(b2s_device_count(&ndev));
    B2S_REQUIRE(ndev > 0, B2S_ERR_NO_DEVICE, "no CUDA device (there is no CPU fallback)");
    DeviceGuard guard;
    int prev = 0;
    cudaGetDevice(&prev);
    guard.prev = prev; guard.active = true;

    b2s_diff3d *h = new b2s_diff3d();
    h->cfg = *cfg;
    h->cfg.devices = nullptr;
    for (int i = 0; i < cfg->slab_count; ++i) {
        const int d = cfg->devices ? cfg->devices[i] : 0;
        if (d < 0 || d >= ndev) {
            set_error("device ordinal %d out of range (have %d)", d, ndev);
            delete h;
            return B2S_ERR_BAD_ARG;
        }
        h->devices.push_back(d);
    }
    compute_params(h);
    h->multi = cfg->nslabs_total > 1;
    h->cart = is_cart(*cfg);
    h->zstack = h->multi && !h->cart;
    // halo flags are sized for the finest tiling any kernel variant uses (64 x 4), so the arena layout does not depend on it
    h->ar.layout((size_t)cfg->nx * cfg->ny * cfg->nz, (size_t)((cfg->nx + kDirBX - 1) / kDirBX) * ((cfg->ny + kDirBY - 1) / kDirBY));

#define FAIL_IF(call)                      \
    do {                                   \
        int rc_ = (call);                  \
        if (rc_ != B2S_OK) {               \
            destroy_impl(h);               \
            return rc_;                    \
        }                                  \
    } while (0)
#define CUDA_FAIL_IF(call)                                                                        \
    do {                                                                                          \
        cudaError_t e_ = (call);                                                                  \
        if (e_ != cudaSuccess) {                                                                  \
            set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_));      \
            destroy_impl(h);                                                                      \
            return B2S_ERR_CUDA;                                                                  \
        }                                                                                         \
    } while (0)

    CUDA_FAIL_IF(cudaMallocHost(&h->pinned, 3 * sizeof(PTState)));
    // device contexts
    for (int i = 0; i < cfg->slab_count; ++i) {
        const int d = h->devices[i];
        int slot = -1;
        for (size_t k = 0; k < h->devs.size(); ++k)
            if (h->devs[k].dev == d) slot = (int)k;
        if (slot < 0) {
            DeviceCtx dc;
            dc.dev = d;
            CUDA_FAIL_IF(cudaSetDevice(d));
            CUDA_FAIL_IF(cudaStreamCreateWithFlags(&dc.stream, cudaStreamNonBlocking));
            CUDA_FAIL_IF(cudaEventCreate(&dc.ev0));
            CUDA_FAIL_IF(cudaEventCreate(&dc.ev1));
            h->devs.push_back(dc);
            slot = (int)h->devs.size() - 1;
        }
        Slab s;
        s.rank = cfg->slab_begin + i;
        s.dev = d;
        s.devslot = slot;
        CUDA_FAIL_IF(cudaSetDevice(d));
        CUDA_FAIL_IF(cudaMalloc(&s.arena, h->ar.bytes));
        CUDA_FAIL_IF(cudaMemset(s.arena, 0, h->ar.bytes));
        s.Ht = (double *)(s.arena + h->ar.off_ht);
        s.buf[0] = (double *)(s.arena + h->ar.off_buf[0]);
        s.buf[1] = (double *)(s.arena + h->ar.off_buf[1]);
        s.slots = (RankSlots *)(s.arena + h->ar.off_slots);
        s.partials = (double *)(s.arena + h->ar.off_partials);
        s.ticket = (unsigned int *)(s.arena + h->ar.off_ticket);
        s.state = (PTState *)(s.arena + h->ar.off_state);
        s.flags = (unsigned long long *)(s.arena + h->ar.off_flags);
        s.table = (RankSlots **)(s.arena + h->ar.off_peer_table);
        h->slabs.push_back(s);
        DeviceCtx &dc = h->devs[slot];
        if (!dc.state) {
            dc.state = (PTState *)(s.arena + h->ar.off_state);
            dc.slots = s.slots;
            dc.peer_table = (RankSlots **)(s.arena + h->ar.off_peer_table);
        }
    }
    CUDA_FAIL_IF(cudaSetDevice(h->devs[0].dev));
    for (int i = 0; i < 2; ++i) CUDA_FAIL_IF(cudaEventCreateWithFlags(&h->ev_poll[i], cudaEventDisableTiming));
    // kernel variant + geometry
    {
        const Slab &s0 = h->slabs[0];
        const int kv = cfg->arithmetic == B2S_ARITH_ARRAY ? B2S_KERNEL_DIRECT : cfg->kernel_variant;
        const bool elig = tma_eligible(cfg->nx, cfg->ny, cfg->nz, s0.Ht, s0.buf[0], s0.buf[1]);
        if (kv == B2S_KERNEL_TMA && !elig) {
            set_error("TMA variant needs even nx >= 64, ny >= 16, nz >= 8 (got %dx%dx%d)", cfg->nx, cfg->ny, cfg->nz);
            destroy_impl(h);
            return B2S_ERR_BAD_ARG;
        }
        h->use_tma = kv == B2S_KERNEL_TMA || (kv == B2S_KERNEL_AUTO && elig && h->ar.cells >= (size_t)64 * 64 * 64);
        if (h->use_tma) {
            h->tma_choice = tma_choice_index(h->ar.cells);
            const TmaChoice &c = kTmaChoices[h->tma_choice];
            h->ntiles = ((cfg->nx + c.tx - 1) / c.tx) * ((cfg->ny + c.ty - 1) / c.ty);
            h->zchunk = pick_zchunk(h->ntiles, cfg->nz, kMaxPartials);
            if (h->zchunk <= 0) {
                set_error("grid %dx%dx%d has more xy tiles than the %d block partials", cfg->nx, cfg->ny, cfg->nz, kMaxPartials);
                destroy_impl(h);
                return B2S_ERR_BAD_SIZE;
            }
            h->nblocks = grid_blocks_tma(h->tma_choice, cfg->nx, cfg->ny, cfg->nz, h->zchunk);
            for (Slab &s : h->slabs) {
                CUDA_FAIL_IF(cudaSetDevice(s.dev));
                FAIL_IF(make_maps(h->tma_choice, s.buf[0], s.Ht, cfg->nx, cfg->ny, cfg->nz, &s.mapBuf[0], &s.mapHt));
                FAIL_IF(make_maps(h->tma_choice, s.buf[1], nullptr, cfg->nx, cfg->ny, cfg->nz, &s.mapBuf[1], nullptr));
            }
        } else {
            const int tiles = ((cfg->nx + kDirBX - 1) / kDirBX) * ((cfg->ny + kDirBY - 1) / kDirBY);
            h->ntiles = tiles;
            h->zchunk = pick_zchunk(tiles, cfg->nz, kMaxPartials);
            if (h->zchunk <= 0) {
                set_error("grid %dx%dx%d has more xy tiles than the %d block partials", cfg->nx, cfg->ny, cfg->nz, kMaxPartials);
                destroy_impl(h);
                return B2S_ERR_BAD_SIZE;
            }
            h->nblocks = tiles * ((cfg->nz - 2 + h->zchunk - 1) / h->zchunk);
        }
    }
    // batches of iterations as CUDA graphs: one device, no cross-device events in the iteration, L2-resident grid
    {
        const int eg = env_int("B2S_DIFF_GRAPH", -1);
        const double cells = (double)h->ar.cells * (double)h->slabs.size();
        h->use_graph = h->devs.size() == 1 && !h->cart && (eg < 0 ? cells <= 200.0 * 200 * 200 : eg != 0);
        h->graph_batch = cfg->batch > 0 ? cfg->batch : ((double)h->ar.cells >= 256.0 * 256 * 256 ? 16 : ((double)h->ar.cells >= 96.0 * 96 * 96 ? 64 : 128));
    }
    // in-process neighbours and partial-sum destinations
    if (cfg->slab_count == cfg->nslabs_total) {
        // peer access between the devices of the handle
        for (size_t a = 0; a < h->devs.size(); ++a)
            for (size_t b = 0; b < h->devs.size(); ++b) {
                if (a == b) continue;
                int can = 0;
                CUDA_FAIL_IF(cudaDeviceCanAccessPeer(&can, h->devs[a].dev, h->devs[b].dev));
                if (!can) {
                    set_error("devices %d and %d cannot access each other's memory", h->devs[a].dev, h->devs[b].dev);
                    destroy_impl(h);
                    return B2S_ERR_CUDA;
                }
                CUDA_FAIL_IF(cudaSetDevice(h->devs[a].dev));
                cudaError_t e = cudaDeviceEnablePeerAccess(h->devs[b].dev, 0);
                if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
                else CUDA_FAIL_IF(e);
            }
        if (!h->cart) {  // z-slabs: the halo push is fused into the step kernel
            for (size_t i = 0; i < h->slabs.size(); ++i) {
                Slab &s = h->slabs[i];
                if (i > 0) { s.lo_buf[0] = h->slabs[i - 1].buf[0]; s.lo_buf[1] = h->slabs[i - 1].buf[1]; s.lo_flags = h->slabs[i - 1].flags; }
                if (i + 1 < h->slabs.size()) {
                    s.hi_buf[0] = h->slabs[i + 1].buf[0]; s.hi_buf[1] = h->slabs[i + 1].buf[1]; s.hi_flags = h->slabs[i + 1].flags;
                }
            }
            // every slab publishes its partial norm into every slab's mailbox
            std::vector<RankSlots *> all;
            for (Slab &s : h->slabs) all.push_back(s.slots);
            for (Slab &s : h->slabs) {
                CUDA_FAIL_IF(cudaSetDevice(s.dev));
                CUDA_FAIL_IF(cudaMemcpy(s.table, all.data(), sizeof(RankSlots *) * all.size(), cudaMemcpyHostToDevice));
            }
        } else {
            for (DeviceCtx &d : h->devs) {
                CUDA_FAIL_IF(cudaSetDevice(d.dev));
                CUDA_FAIL_IF(cudaEventCreateWithFlags(&d.ev_bar, cudaEventDisableTiming));
            }
            // one rank per device: phase mailboxes of all ranks, reachable through peer access
            h->cart_devbarrier = h->slabs.size() > 1 && h->devs.size() == h->slabs.size() && env_int("B2S_CART_EVENTS", 0) == 0;
            if (h->cart_devbarrier) {
                std::vector<CartSync *> ct;
                for (Slab &s : h->slabs) ct.push_back((CartSync *)(s.arena + h->ar.off_cart));
                for (Slab &s : h->slabs) {
                    CUDA_FAIL_IF(cudaSetDevice(s.dev));
                    CUDA_FAIL_IF(cudaMemcpy(s.arena + h->ar.off_cart_table, ct.data(), sizeof(CartSync *) * ct.size(), cudaMemcpyHostToDevice));
                }
            }
        }
        if (h->cart) {
            std::vector<RankSlots *> tbl;
            for (DeviceCtx &d : h->devs) tbl.push_back(d.slots);
            h->nslots_dst = (int)tbl.size();
            for (DeviceCtx &d : h->devs) {
                CUDA_FAIL_IF(cudaSetDevice(d.dev));
                CUDA_FAIL_IF(cudaMemcpy(d.peer_table, tbl.data(), sizeof(RankSlots *) * tbl.size(), cudaMemcpyHostToDevice));
            }
        }
        h->connected = true;
    }
#undef FAIL_IF
#undef CUDA_FAIL_IF
    *out = h;
    return B2S_OK;
}

int b2s_diff3d_destroy(b2s_diff3d *h)
{
    int prev = -1;
    cudaGetDevice(&prev);
    destroy_impl(h);
    if (prev >= 0) cudaSetDevice(prev);
    return B2S_OK;
}

int b2s_diff3d_params_for(const b2s_diff3d_config *cfg, b2s_diff3d_params *out)
{
    B2S_REQUIRE(cfg && out, B2S_ERR_BAD_ARG, "NULL argument");
    B2S_REQUIRE(cfg->nx >= 3 && cfg->ny >= 3 && cfg->nz >= 3 && cfg->nslabs_total >= 1, B2S_ERR_BAD_SIZE, "bad grid");
    compute_params_raw(*cfg, *out);
    return B2S_OK;
}

int b2s_diff3d_get_params(const b2s_diff3d *h, b2s_diff3d_params *out)
{
    B2S_REQUIRE(h && out, B2S_ERR_BAD_ARG, "NULL argument");
    *out = h->prm;
    return B2S_OK;
}

int b2s_diff3d_set_initial(b2s_diff3d *h, const double *Ht_host)
{
    B2S_REQUIRE(h && Ht_host, B2S_ERR_BAD_ARG, "NULL argument");
    DeviceGuard guard;
    guard.set(h->devs[0].dev);
    const size_t n = h->ar.cells;
    // Mailboxes, halo flags and the sequence number are never reset: they are monotonic, so a re-initialisation of a
    // connected one-process-per-GPU handle cannot erase what a peer has already stored (the caller still has to put a
    // barrier across the ranks between set_initial / init_gaussian and the first iteration, see b200stencil.h).
    for (size_t i = 0; i < h->slabs.size(); ++i) {
        Slab &s = h->slabs[i];
        DeviceCtx &d = h->devs[s.devslot];
        B2S_CUDA(cudaSetDevice(s.dev));
        B2S_CUDA(cudaStreamSynchronize(d.stream));
        B2S_CUDA(cudaMemcpy(s.Ht, Ht_host + i * n, n * sizeof(double), cudaMemcpyHostToDevice));
        B2S_CUDA(cudaMemcpy(s.buf[0], s.Ht, n * sizeof(double), cudaMemcpyDeviceToDevice));  // Htau = copy(Ht)
        if (h->cfg.arithmetic == B2S_ARITH_ARRAY)  // in-place update of Htau (part1_array_programming.jl:17): the frame is Ht's forever
            B2S_CUDA(cudaMemcpy(s.buf[1], s.Ht, n * sizeof(double), cudaMemcpyDeviceToDevice));
        else
            B2S_CUDA(cudaMemset(s.buf[1], 0, n * sizeof(double)));                             // Htau2 = @zeros
        B2S_CUDA(cudaMemset(s.ticket, 0, 64));
    }
    h->launched = 0;
    h->skip_push = 0;
    h->it_step = 0;
    return B2S_OK;
}

int b2s_diff3d_init_gaussian(b2s_diff3d *h)
{
    B2S_REQUIRE(h, B2S_ERR_BAD_ARG, "NULL handle");
    // init_local_gaussian (part1_utils.jl:1-12) + apply_boundary_conditions! (:14-34), on the host like the reference.
    const b2s_diff3d_config &c = h->cfg;
    const b2s_diff3d_params &p = h->prm;
    const int nx = c.nx, ny = c.ny, nz = c.nz;
    const size_t n = h->ar.cells;
    std::vector<double> host(n * h->slabs.size());
    const double cx = p.lx / 2, cy = p.ly / 2, cz = p.lz / 2;
    for (size_t si = 0; si < h->slabs.size(); ++si) {
        double *Ht = host.data() + si * n;
        int coord[3], dims[3];
        cart_coords(c, h->slabs[si].rank, coord);
        cart_dims(c, dims);
#pragma omp parallel for schedule(static)
        for (int k = 0; k < nz; ++k)
            for (int j = 0; j < ny; ++j)
                for (int i = 0; i < nx; ++i) {
                    // x_g(ix,dx,H) = (coords*(n-2) + ix-1)*dx  [ImplicitGlobalGrid, overlap 2]
                    const double xg = (double)(coord[0] * (nx - 2) + i) * p.dx;
                    const double yg = (double)(coord[1] * (ny - 2) + j) * p.dy;
                    const double zg = (double)(coord[2] * (nz - 2) + k) * p.dz;
                    const double ax = xg + p.dx / 2 - cx, ay = yg + p.dy / 2 - cy, az = zg + p.dz / 2 - cz;
                    Ht[(size_t)i + (size_t)nx * (j + (size_t)ny * k)] = 2 * exp(-1.0 * ((ax * ax + ay * ay) + az * az));
                }
        const int nd[3] = {nx, ny, nz};
        for (int d = 0; d < 3; ++d) {
            bool zero_lo, zero_hi;
            if (c.bc_mode == B2S_BC_LITERAL) {  // the tests are written for 1-based coords, IGG's are 0-based
                zero_lo = coord[d] == 1;
                zero_hi = coord[d] == dims[d];
            } else {
                zero_lo = coord[d] == 0;
                zero_hi = coord[d] == dims[d] - 1;
            }
            for (int side = 0; side < 2; ++side) {
                if (!(side == 0 ? zero_lo : zero_hi)) continue;
                const int fixed = side == 0 ? 0 : nd[d] - 1;
                for (int k = 0; k < nz; ++k)
                    for (int j = 0; j < ny; ++j)
                        for (int i = 0; i < nx; ++i) {
                            const int idx[3] = {i, j, k};
                            if (idx[d] == fixed) Ht[(size_t)i + (size_t)nx * (j + (size_t)ny * k)] = 0.0;
                        }
            }
        }
    }
    return b2s_diff3d_set_initial(h, host.data());
}

size_t b2s_diff3d_ipc_blob_bytes(void) { return sizeof(IpcBlob); }

int b2s_diff3d_ipc_export(b2s_diff3d *h, void *blob_out)
{
    B2S_REQUIRE(h && blob_out, B2S_ERR_BAD_ARG, "NULL argument");
    B2S_REQUIRE(h->slabs.size() == 1, B2S_ERR_STATE, "IPC export is for one-slab-per-process handles");
    DeviceGuard guard;
    guard.set(h->slabs[0].dev);
    IpcBlob b;
    memset(&b, 0, sizeof(b));
    B2S_CUDA(cudaIpcGetMemHandle(&b.handle, h->slabs[0].arena));
    b.rank = h->slabs[0].rank;
    b.dev = h->slabs[0].dev;
    b.arena_bytes = h->ar.bytes;
    memcpy(blob_out, &b, sizeof(b));
    return B2S_OK;
}

int b2s_diff3d_ipc_connect(b2s_diff3d *h, const void *all_blobs, int nblobs)
{
    B2S_REQUIRE(h && all_blobs, B2S_ERR_BAD_ARG, "NULL argument");
    B2S_REQUIRE(h->slabs.size() == 1 && h->multi, B2S_ERR_STATE, "IPC connect is for one-slab-per-process handles");
    B2S_REQUIRE(nblobs == h->cfg.nslabs_total, B2S_ERR_BAD_ARG, "expected %d blobs, got %d", h->cfg.nslabs_total, nblobs);
    B2S_REQUIRE(!h->connected, B2S_ERR_STATE, "already connected");
    Slab &s = h->slabs[0];
    DeviceGuard guard;
    guard.set(s.dev);
    const IpcBlob *blobs = (const IpcBlob *)all_blobs;
    std::vector<RankSlots *> tbl(nblobs, nullptr);
    h->peer_base.assign(nblobs, nullptr);
    for (int r = 0; r < nblobs; ++r) {
        B2S_REQUIRE(blobs[r].rank == r && blobs[r].arena_bytes == h->ar.bytes, B2S_ERR_BAD_ARG,
                    "blob %d does not describe slab %d of a compatible handle", r, r);
        char *base = nullptr;
        if (r == s.rank) {
            base = s.arena;
        } else {
            void *m = nullptr;
            B2S_CUDA(cudaIpcOpenMemHandle(&m, blobs[r].handle, cudaIpcMemLazyEnablePeerAccess));
            h->ipc_mapped.push_back(m);
            base = (char *)m;
        }
        tbl[r] = (RankSlots *)(base + h->ar.off_slots);
        h->peer_base[r] = base;
        if (!h->cart) {  // z-slabs: the neighbours' buffers receive the fused halo push
            if (r == s.rank - 1) {
                s.lo_buf[0] = (double *)(base + h->ar.off_buf[0]); s.lo_buf[1] = (double *)(base + h->ar.off_buf[1]);
                s.lo_flags = (unsigned long long *)(base + h->ar.off_flags);
            }
            if (r == s.rank + 1) {
                s.hi_buf[0] = (double *)(base + h->ar.off_buf[0]); s.hi_buf[1] = (double *)(base + h->ar.off_buf[1]);
                s.hi_flags = (unsigned long long *)(base + h->ar.off_flags);
            }
        }
    }
    h->nslots_dst = nblobs;
    B2S_CUDA(cudaMemcpy(h->devs[0].peer_table, tbl.data(), sizeof(RankSlots *) * nblobs, cudaMemcpyHostToDevice));
    if (h->cart) {  // phase mailboxes of all ranks (cart_barrier_kernel)
        std::vector<CartSync *> ct(nblobs, nullptr);
        for (int r = 0; r < nblobs; ++r) ct[r] = (CartSync *)(h->peer_base[r] + h->ar.off_cart);
        B2S_CUDA(cudaMemcpy(s.arena + h->ar.off_cart_table, ct.data(), sizeof(CartSync *) * nblobs, cudaMemcpyHostToDevice));
    }
    h->connected = true;
    return B2S_OK;
}

int b2s_diff3d_solve_timestep(b2s_diff3d *h, double tol, int iter_max, int *iters, double *err)
{
    B2S_REQUIRE(h, B2S_ERR_BAD_ARG, "NULL handle");
    PTState st = {};
    st.it = 0;
    st.iter_max = iter_max;
    st.check = 1;
    st.tol = tol;
    st.err = 2 * tol;  // part1_kernel_programming.jl:178
    st.sqrt_total_N = sqrt(h->prm.total_N);
    st.hist_cap = 0;
    PTState fin;
    for (DeviceCtx &d : h->devs) d.err_hist_cap = d.err_hist_cap;  // history not recorded here
    B2S_CHECK(run_loop(h, st, &fin));
    if (iters) *iters = fin.it;
    if (err) *err = fin.err;
    return B2S_OK;
}

int b2s_diff3d_iterate(b2s_diff3d *h, int n, double *err_hist)
{
    B2S_REQUIRE(h && n >= 0, B2S_ERR_BAD_ARG, "bad argument");
    if (n == 0) { h->last_ms = 0.0; return B2S_OK; }
    DeviceGuard guard;
    guard.set(h->devs[0].dev);
    if (err_hist) B2S_CHECK(ensure_hist(h, n));
    PTState st = {};
    st.it = 0;
    st.iter_max = n;
    st.check = 0;
    st.tol = 0.0;
    st.err = 0.0;
    st.sqrt_total_N = sqrt(h->prm.total_N);
    st.hist_cap = err_hist ? n : 0;
    PTState fin;
    B2S_CHECK(run_loop(h, st, &fin));
    if (err_hist) {
        DeviceCtx &d0 = h->devs[0];
        B2S_CUDA(cudaSetDevice(d0.dev));
        B2S_CUDA(cudaMemcpy(err_hist, d0.err_hist, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost));
    }
    return B2S_OK;
}

int b2s_diff3d_advance_time(b2s_diff3d *h)
{
    B2S_REQUIRE(h, B2S_ERR_BAD_ARG, "NULL handle");
    DeviceGuard guard;
    guard.set(h->devs[0].dev);
    const int cur = (int)(h->launched & 1);
    for (Slab &s : h->slabs) {
        DeviceCtx &d = h->devs[s.devslot];
        B2S_CUDA(cudaSetDevice(s.dev));
        B2S_CUDA(cudaMemcpyAsync(s.Ht, s.buf[cur], h->ar.cells * sizeof(double), cudaMemcpyDeviceToDevice, d.stream));
    }
    return B2S_OK;
}

int b2s_diff3d_run(b2s_diff3d *h, double ttot, double tol, int iter_max, int *iters_per_step, int cap, int *nsteps)
{
    B2S_REQUIRE(h, B2S_ERR_BAD_ARG, "NULL handle");
    // length of the Julia range 0:dt:ttot-dt (part1_kernel_programming.jl:166)
    const double stop = ttot - h->prm.dt;
    const int nt = stop < 0 ? 0 : (int)floor(stop / h->prm.dt + 1e-9) + 1;
    double ms = 0.0;
    for (int t = 0; t < nt; ++t) {
        int it = 0;
        B2S_CHECK(b2s_diff3d_solve_timestep(h, tol, iter_max, &it, nullptr));
        ms += h->last_ms;
        if (iters_per_step && t < cap) iters_per_step[t] = it;
        B2S_CHECK(b2s_diff3d_advance_time(h));
    }
    h->last_ms = ms;
    if (nsteps) *nsteps = nt;
    return B2S_OK;
}

int b2s_diff3d_device_ptr(b2s_diff3d *h, int slab, int which, double **dev_out)
{
    B2S_REQUIRE(h && dev_out, B2S_ERR_BAD_ARG, "NULL argument");
    const int i = slab_index(h, slab);
    B2S_REQUIRE(i >= 0, B2S_ERR_BAD_ARG, "slab %d is not hosted by this handle", slab);
    B2S_REQUIRE(which >= 0 && which <= 2, B2S_ERR_BAD_ARG, "which must be 0 (Ht), 1 (Htau) or 2 (Htau2)");
    const int cur = (int)(h->launched & 1);
    Slab &s = h->slabs[i];
    *dev_out = which == 0 ? s.Ht : (which == 1 ? s.buf[cur] : s.buf[cur ^ 1]);
    return B2S_OK;
}

int b2s_diff3d_get_field(b2s_diff3d *h, int slab, int which, double *host_out)
{
    B2S_REQUIRE(host_out, B2S_ERR_BAD_ARG, "NULL argument");
    double *src = nullptr;
    B2S_CHECK(b2s_diff3d_device_ptr(h, slab, which, &src));
    Slab &s = h->slabs[slab_index(h, slab)];
    DeviceGuard guard;
    guard.set(s.dev);
    B2S_CUDA(cudaStreamSynchronize(h->devs[s.devslot].stream));
    B2S_CUDA(cudaMemcpy(host_out, src, h->ar.cells * sizeof(double), cudaMemcpyDeviceToHost));
    return B2S_OK;
}

int b2s_diff3d_gather(b2s_diff3d *h, double *H_g_host)
{
    B2S_REQUIRE(h && H_g_host, B2S_ERR_BAD_ARG, "NULL argument");
    // H_g has size (nx*dims[1], ny*dims[2], nz*dims[3]) and receives every rank's whole local Ht (:144,223)
    if (!h->cart) {  // z-slabs: rank-major blocks are the global layout
        for (size_t i = 0; i < h->slabs.size(); ++i)
            B2S_CHECK(b2s_diff3d_get_field(h, h->slabs[i].rank, 0, H_g_host + i * h->ar.cells));
        return B2S_OK;
    }
    const b2s_diff3d_config &c = h->cfg;
    int dims[3];
    cart_dims(c, dims);
    std::vector<double> loc(h->ar.cells);
    const size_t gx = (size_t)c.nx * dims[0], gy = (size_t)c.ny * dims[1];
    for (size_t i = 0; i < h->slabs.size(); ++i) {
        B2S_CHECK(b2s_diff3d_get_field(h, h->slabs[i].rank, 0, loc.data()));
        int coord[3];
        cart_coords(c, h->slabs[i].rank, coord);
        for (int k = 0; k < c.nz; ++k)
            for (int j = 0; j < c.ny; ++j)
                memcpy(H_g_host + ((size_t)coord[0] * c.nx + gx * (((size_t)coord[1] * c.ny + j) + gy * ((size_t)coord[2] * c.nz + k))),
                       loc.data() + (size_t)c.nx * (j + (size_t)c.ny * k), (size_t)c.nx * sizeof(double));
    }
    return B2S_OK;
}

// A new job on an existing handle starts exactly like a fresh handle after set_initial (the reference allocates per run:
// Htau = copy(Ht), Htau2 = @zeros, part1_kernel_programming.jl:141-142): Htau goes to buffer 0, the other buffer is
// cleared, the ping-pong parity restarts. `src` (device) holds the new Ht. In a z-slab stack the halo planes of buffer 1
// that a neighbour stores into are left alone: the neighbour's first kernel of the new job overwrites them before they
// are read (with zeros in lag-2 mode -- the old content of its own cleared buffer --, with new values otherwise), and
// clearing them here could race with that store.
static int start_job_on_stream(b2s_diff3d *h, Slab &s, DeviceCtx &d, const double *src)
{
    const size_t bytes = h->ar.cells * sizeof(double);
    if (src != s.Ht) B2S_CUDA(cudaMemcpyAsync(s.Ht, src, bytes, cudaMemcpyDeviceToDevice, d.stream));
    B2S_CUDA(cudaMemcpyAsync(s.buf[0], src, bytes, cudaMemcpyDeviceToDevice, d.stream));
    if (h->cfg.arithmetic == B2S_ARITH_ARRAY) {
        B2S_CUDA(cudaMemcpyAsync(s.buf[1], src, bytes, cudaMemcpyDeviceToDevice, d.stream));
    } else {
        const size_t plane = (size_t)h->cfg.nx * h->cfg.ny;
        const size_t first = (h->zstack && s.lo_buf[1]) ? 1 : 0;
        const size_t last = (h->zstack && s.hi_buf[1]) ? (size_t)h->cfg.nz - 2 : (size_t)h->cfg.nz - 1;
        B2S_CUDA(cudaMemsetAsync(s.buf[1] + plane * first, 0, plane * (last - first + 1) * sizeof(double), d.stream));
    }
    h->launched = 0;
    h->skip_push = 0;
    h->it_step = 0;
    return B2S_OK;
}

int b2s_diff3d_upload_state(b2s_diff3d *h, int slab, const double *Ht_host)
{
    B2S_REQUIRE(h && Ht_host, B2S_ERR_BAD_ARG, "NULL argument");
    const int i = slab_index(h, slab);
    B2S_REQUIRE(i >= 0, B2S_ERR_BAD_ARG, "slab %d is not hosted by this handle", slab);
    Slab &s = h->slabs[i];
    DeviceCtx &d = h->devs[s.devslot];
    DeviceGuard guard;
    guard.set(s.dev);
    const size_t bytes = h->ar.cells * sizeof(double);
    B2S_CUDA(cudaMemcpyAsync(s.Ht, Ht_host, bytes, cudaMemcpyHostToDevice, d.stream));
    return start_job_on_stream(h, s, d, s.Ht);
}

// copy stream + events of the pipelined transfers (created on first use)
static int ensure_copy_stream(DeviceCtx &d)
{
    if (!d.copy_stream) {
        B2S_CUDA(cudaStreamCreateWithFlags(&d.copy_stream, cudaStreamNonBlocking));
        B2S_CUDA(cudaStreamCreateWithFlags(&d.copy_stream_down, cudaStreamNonBlocking));
        B2S_CUDA(cudaEventCreateWithFlags(&d.ev_work, cudaEventDisableTiming));
        B2S_CUDA(cudaEventCreateWithFlags(&d.ev_down, cudaEventDisableTiming));
        B2S_CUDA(cudaEventCreateWithFlags(&d.ev_up, cudaEventDisableTiming));
        B2S_CUDA(cudaEventCreateWithFlags(&d.ev_staged_free, cudaEventDisableTiming));
    }
    return B2S_OK;
}

int b2s_diff3d_download_state_async(b2s_diff3d *h, int slab, double *Htau_host)
{
    B2S_REQUIRE(h && Htau_host, B2S_ERR_BAD_ARG, "NULL argument");
    const int i = slab_index(h, slab);
    B2S_REQUIRE(i >= 0, B2S_ERR_BAD_ARG, "slab %d is not hosted by this handle", slab);
    Slab &s = h->slabs[i];
    DeviceCtx &d = h->devs[s.devslot];
    DeviceGuard guard;
    guard.set(s.dev);
    B2S_CHECK(ensure_copy_stream(d));
    const int cur = (int)(h->launched & 1);
    const size_t bytes = h->ar.cells * sizeof(double);
    // The result is first copied aside on the device (0.3 ms at 512^3), so the solver's buffers are free for the next
    // job at once; the PCIe transfer then runs from that copy on the copy stream, overlapped with whatever comes next.
    if (!s.stage_out) B2S_CUDA(cudaMalloc(&s.stage_out, bytes));
    if (d.download_pending) B2S_CUDA(cudaStreamWaitEvent(d.stream, d.ev_down, 0));  // previous transfer out of stage_out
    B2S_CUDA(cudaMemcpyAsync(s.stage_out, s.buf[cur], bytes, cudaMemcpyDeviceToDevice, d.stream));
    B2S_CUDA(cudaEventRecord(d.ev_work, d.stream));
    B2S_CUDA(cudaStreamWaitEvent(d.copy_stream_down, d.ev_work, 0));
    B2S_CUDA(cudaMemcpyAsync(Htau_host, s.stage_out, bytes, cudaMemcpyDeviceToHost, d.copy_stream_down));
    B2S_CUDA(cudaEventRecord(d.ev_down, d.copy_stream_down));
    d.download_pending = true;
    return B2S_OK;
}

int b2s_diff3d_upload_state_async(b2s_diff3d *h, int slab, const double *Ht_host)
{
    B2S_REQUIRE(h && Ht_host, B2S_ERR_BAD_ARG, "NULL argument");
    const int i = slab_index(h, slab);
    B2S_REQUIRE(i >= 0, B2S_ERR_BAD_ARG, "slab %d is not hosted by this handle", slab);
    Slab &s = h->slabs[i];
    DeviceCtx &d = h->devs[s.devslot];
    DeviceGuard guard;
    guard.set(s.dev);
    B2S_CHECK(ensure_copy_stream(d));
    const size_t bytes = h->ar.cells * sizeof(double);
    if (!s.staging) B2S_CUDA(cudaMalloc(&s.staging, bytes));
    // the previous commit (compute stream) must have consumed the staging array before it is overwritten
    if (s.staging_used) B2S_CUDA(cudaStreamWaitEvent(d.copy_stream, d.ev_staged_free, 0));
    B2S_CUDA(cudaMemcpyAsync(s.staging, Ht_host, bytes, cudaMemcpyHostToDevice, d.copy_stream));
    B2S_CUDA(cudaEventRecord(d.ev_up, d.copy_stream));
    s.staged = true;
    return B2S_OK;
}

int b2s_diff3d_commit_upload(b2s_diff3d *h, int slab)
{
    B2S_REQUIRE(h, B2S_ERR_BAD_ARG, "NULL handle");
    const int i = slab_index(h, slab);
    B2S_REQUIRE(i >= 0, B2S_ERR_BAD_ARG, "slab %d is not hosted by this handle", slab);
    Slab &s = h->slabs[i];
    DeviceCtx &d = h->devs[s.devslot];
    B2S_REQUIRE(s.staged, B2S_ERR_STATE, "no staged upload (call b2s_diff3d_upload_state_async first)");
    DeviceGuard guard;
    guard.set(s.dev);
    B2S_CUDA(cudaStreamWaitEvent(d.stream, d.ev_up, 0));
    B2S_CHECK(start_job_on_stream(h, s, d, s.staging));
    B2S_CUDA(cudaEventRecord(d.ev_staged_free, d.stream));
    s.staged = false;
    s.staging_used = true;
    return B2S_OK;
}

int b2s_diff3d_sync(b2s_diff3d *h)
{
    B2S_REQUIRE(h, B2S_ERR_BAD_ARG, "NULL handle");
    DeviceGuard guard;
    guard.set(h->devs[0].dev);
    for (DeviceCtx &d : h->devs) {
        B2S_CUDA(cudaSetDevice(d.dev));
        B2S_CUDA(cudaStreamSynchronize(d.stream));
        if (d.copy_stream) B2S_CUDA(cudaStreamSynchronize(d.copy_stream));
        if (d.copy_stream_down) B2S_CUDA(cudaStreamSynchronize(d.copy_stream_down));
    }
    return B2S_OK;
}

int b2s_diff3d_download_state(b2s_diff3d *h, int slab, double *Htau_host)
{
    B2S_REQUIRE(h && Htau_host, B2S_ERR_BAD_ARG, "NULL argument");
    const int i = slab_index(h, slab);
    B2S_REQUIRE(i >= 0, B2S_ERR_BAD_ARG, "slab %d is not hosted by this handle", slab);
    Slab &s = h->slabs[i];
    DeviceCtx &d = h->devs[s.devslot];
    DeviceGuard guard;
    guard.set(s.dev);
    const int cur = (int)(h->launched & 1);
    B2S_CUDA(cudaMemcpyAsync(Htau_host, s.buf[cur], h->ar.cells * sizeof(double), cudaMemcpyDeviceToHost, d.stream));
    B2S_CUDA(cudaStreamSynchronize(d.stream));
    return B2S_OK;
}

int b2s_diff3d_stats(const b2s_diff3d *h, long long *kernel_launches, double *last_call_ms)
{
    B2S_REQUIRE(h, B2S_ERR_BAD_ARG, "NULL handle");
    if (kernel_launches) *kernel_launches = h->kernel_launches;
    if (last_call_ms) *last_call_ms = h->last_ms;
    return B2S_OK;
}

}  // extern "C"
