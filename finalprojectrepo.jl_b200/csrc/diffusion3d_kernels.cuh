// diffusion3d_kernels.cuh -- device code of hot path 1: the fused flux/residual/update step of the 3-D dual-time
// (pseudo-transient) diffusion solver, with the residual norm, the PT-loop exit test and the z-slab halo exchange
// fused in.
//
// Reference semantics (file:line relative to the reference repository):
//   scripts-part1/part1_kernel_programming.jl:12-20   @qx/@qy/@qz      q(i) = -D_d * (H[i] - H[i-1])
//   scripts-part1/part1_kernel_programming.jl:46-58   diffusion_3D_step_tau (and the shared-memory twin :75-97)
//   scripts-part1/part1_kernel_programming.jl:177-193 while err > tol && iter < iter_max ... update_halo!(Htau); swap
//   scripts-part1/part1_utils.jl:36-40                dist_norm_L2
// Arithmetic is evaluated in exactly the order written there and this translation unit is compiled with
// -fmad=false, so single-GPU fields are bit-identical to a non-contracting CPU evaluation.
#pragma once
#include "common.cuh"

namespace b2s {

// Device-resident state of the PT loop of one time step (one copy per device that hosts slabs of a handle).
struct PTState {
    int it;          // iter_inner of the current time step
    int done;        // exit flag: kernels launched after it is set return immediately
    int iter_max;    // part1_kernel_programming.jl:130
    int check;       // 1: evaluate "err > tol" (solve_timestep); 0: fixed count (iterate)
    int error;       // 1: cross-GPU wait timed out
    int hist_pos;    // next entry of err_hist
    int hist_cap;
    int pending;     // z-slab stacks: 1 = the iteration of the last executed step kernel has not been evaluated yet
    double tol;
    double err;
    double sumsq;          // global sum over ranks and cells of (R*dt)^2
    double sqrt_total_N;   // sqrt(prod(dims)*nx*ny*nz), part1_kernel_programming.jl:124,191
    long long total_iters; // PT iterations since create (selects the ping-pong parity)
    unsigned long long seq; // cross-GPU sequence number of the next partial to publish / consume
    int skip_push;   // z-slab stacks, lag-2 halos: the next step kernel signals its halo flags without storing planes
    int pad;
};

constexpr int kMaxRanks = 64;
constexpr int kSlotGens = 4;

// Cross-GPU reduction mailbox living in every rank's arena: peers store their partial and then the sequence number.
// Four generations (seq mod 4): with the lagged evaluation of z-slab stacks a rank publishes partial s+4 only after every
// rank has finished kernel s+2, i.e. after partial s has been consumed everywhere (two generations suffice for the
// per-iteration handshake of general decompositions).
// A partial travels as two self-validating 8-byte words (value bits 31..0 | seq32 << 32, value bits 63..32 | seq32 << 32):
// naturally aligned 8-byte stores are single transactions, so the sender needs NO fence and no separate flag store (a
// release per destination used to cost one NVLink round trip each, in the tail of every step kernel) and the receiver
// polls until both words carry the sequence number it waits for.
struct RankSlots {
    unsigned long long w[kSlotGens][kMaxRanks][2];
};

__device__ __forceinline__ void slot_publish(RankSlots *dst, int gen, int rank, double value, unsigned long long seq)
{
    const unsigned long long bits = (unsigned long long)__double_as_longlong(value);
    const unsigned long long tag = (seq & 0xffffffffull) << 32;
    st_relaxed_sys_u64(&dst->w[gen][rank][0], (bits & 0xffffffffull) | tag);
    st_relaxed_sys_u64(&dst->w[gen][rank][1], (bits >> 32) | tag);
}
// false on timeout
__device__ __forceinline__ bool slot_consume(const RankSlots *mine, int gen, int rank, unsigned long long seq, long long timeout,
                                             double *value)
{
    const unsigned long long tag = seq & 0xffffffffull;
    const long long t0 = clock64();
    for (;;) {
        const unsigned long long a = ld_relaxed_sys_u64(&mine->w[gen][rank][0]);
        const unsigned long long b = ld_relaxed_sys_u64(&mine->w[gen][rank][1]);
        if ((a >> 32) == tag && (b >> 32) == tag) {
            *value = __longlong_as_double((long long)((a & 0xffffffffull) | (b << 32)));
            return true;
        }
        if (clock64() - t0 > timeout) return false;
        __nanosleep(32);
    }
}

// z-slab stacks: per-tile halo flags in every rank's arena, written by the two z neighbours (peer stores), read locally.
// A flag holds the sequence number of the neighbour's step kernel that set it (monotonic, never reset).
//   kFlagLo / kFlagHi: the block of the low / high neighbour that owns this tile of the adjacent z chunk has finished its
//   kernel `value`: it has stored its boundary plane into my halo plane (data ready) AND it has read the plane I stored
//   into its halo plane one kernel earlier (free to overwrite) -- one flag serves both directions of the dependency.
enum { kFlagLo = 0, kFlagHi = 1 };

// General decompositions with one process per GPU: phase mailbox in every rank's arena. phase[r] is the last phase rank r
// has completed (4*seq + k: k-th barrier of the PT iteration with sequence number seq), written by rank r itself.
struct CartSync {
    unsigned long long phase[kMaxRanks];
};

struct StepParams {
    const double *Ht;
    const double *A;  // Htau  (read)
    double *B;        // Htau2 (written, interior only)
    double *R;        // dHdtau, nullable
    int nx, ny, nz;
    double dtau, _dt, _dx, _dy, _dz, mD_dx, mD_dy, mD_dz;  // mD_* = -D_d*
    double norm_scale;                                     // dt in "residual_H * dt"
    double *partials;                                      // one per block
    unsigned int *ticket;
    double *sumsq_out;   // nullable: receives the local sum
    PTState *state;      // nullable (L0 call)
    double *err_hist;    // nullable
    int fuse_finalize;   // 1: single slab -> the last block runs the exit test itself
    // z-slab halo exchange fused into the kernel. Pointers to the first element of the neighbour's
    // halo plane of the buffer with the same parity as B (peer-mapped when the neighbour is on another GPU).
    double *push_lo;     // low neighbour's plane nz-1, nullable
    double *push_hi;     // high neighbour's plane 0, nullable
    int consistent;      // 0: reference lag-2 semantics (forward the OLD content of B's boundary planes);
                         // 1: forward the values just computed
    // one-process-per-GPU mode: publish the local sum to every rank's RankSlots
    RankSlots *const *peer_slots;  // device array [nranks] of (peer-mapped) pointers, nullable
    int nranks, myrank;
    int zchunk;          // interior planes per block
    // z-slab stacks (flagged protocol, see "Halo flags" below)
    int flagged;                      // 1: neighbour flags + lagged evaluation of the norm
    int ntiles;                       // xy tiles of this launch = flags per array
    unsigned long long *flags;        // my 4 flag arrays [4][ntiles]
    unsigned long long *lo_flags;     // low / high neighbour's flag arrays (peer-mapped), nullable
    unsigned long long *hi_flags;
    RankSlots *my_slots;              // my own mailbox (the last block consumes the previous iteration's partials)
    long long timeout_cycles;
    int array_arith;                  // 1: arithmetic of part1_array_programming.jl:9-18 (direct kernel only)
};

__device__ __forceinline__ void pt_finalize(PTState *s, double total, double *err_hist)
{
    // err = dist_norm_L2(residual_H*dt)/sqrt(total_N)   part1_kernel_programming.jl:191
    const double err = sqrt(total) / s->sqrt_total_N;
    const int it = s->it + 1;
    s->it = it;
    s->err = err;
    s->sumsq = total;
    s->total_iters += 1;
    if (err_hist != nullptr && s->hist_pos < s->hist_cap) err_hist[s->hist_pos++] = err;
    // while err > tol && iter_inner < iter_max        part1_kernel_programming.jl:179
    const bool go_on = s->check ? (err > s->tol && it < s->iter_max) : (it < s->iter_max);
    if (!go_on) s->done = 1;
}

// Residual of one cell, operation order of part1_kernel_programming.jl:48-53.
__device__ __forceinline__ double cell_residual(double c, double xl, double xr, double ys, double yn, double zp, double zn,
                                                double ht, const StepParams &p)
{
    return ((p.mD_dx * (xr - c)) - (p.mD_dx * (c - xl))) * p._dx + ((p.mD_dy * (yn - c)) - (p.mD_dy * (c - ys))) * p._dy +
           ((p.mD_dz * (zn - c)) - (p.mD_dz * (c - zp))) * p._dz + (c - ht) * p._dt;
}

// The same residual from ready-made fluxes q(i) = -D_d*(H[i] - H[i-1]) (the macros of part1_kernel_programming.jl:12-20):
// R = (qx(i+1) - qx(i))*_dx + (qy(j+1) - qy(j))*_dy + (qz(k+1) - qz(k))*_dz + (H - Ht)*_dt. A flux between two cells is the
// same number for both of them, so the z-marching kernel computes every z flux once (carried to the next plane in a
// register) and the x flux between the two cells of a thread once: 3 of 28 FP64 operations per cell less, same bits.
__device__ __forceinline__ double cell_residual_flux(double qx_lo, double qx_hi, double qy_lo, double qy_hi, double qz_lo, double qz_hi,
                                                     double c, double ht, const StepParams &p)
{
    return (qx_hi - qx_lo) * p._dx + (qy_hi - qy_lo) * p._dy + (qz_hi - qz_lo) * p._dz + (c - ht) * p._dt;
}

// ---- Halo flags: the z-slab protocol -----------------------------------------------------------------------------
// No rank ever waits for ALL ranks on the critical path. Kernel number s (PTState::seq, identical on all ranks) of a rank
//   * reads its halo planes 0 / nz-1 of Htau, which the neighbour's kernel s-1 has stored   (RAW)
//   * stores planes into the neighbour's Htau2 halo, which the neighbour's kernel s-1 read  (WAR)
// Both hold once the neighbour's block that owns the same xy tile t of the adjacent z chunk has finished kernel s-1, so
// the blocks that own the first / last z chunk wait for flag[t] >= s-1 (one acquire load by one thread) and store
// flag[t] = s into the neighbour's arena when they are done (one __syncthreads, one system fence, one 8-byte store by
// one thread; no other block fences). They are scheduled right after the first interior blocks, so their planes reach
// the neighbour within the first quarter of the kernel and a neighbour that is up to ~3/4 of a kernel behind never stalls anybody.
// All waits refer to strictly earlier kernels, so every rank makes progress independently of co-residency.
// The global norm is evaluated one kernel late: the last block of kernel s publishes its partial (as before) and then
// consumes the partials of kernel s-1, which every rank published a whole kernel ago. If that test ends the PT loop,
// kernel s was speculative: it is not counted, the ping-pong still holds the accepted state, and in lag-2 mode the next
// executed kernel skips its plane stores (the speculative kernel has already forwarded exactly those values:
// the old content of Htau2's boundary planes) -- bit-identical to the per-iteration handshake it replaces.
__device__ __forceinline__ bool flag_wait(const unsigned long long *f, unsigned long long want, long long timeout)
{
    if (ld_acquire_sys_u64(f) >= want) return true;
    const long long t0 = clock64();
    while (ld_acquire_sys_u64(f) < want) {
        if (clock64() - t0 > timeout) return false;
        __nanosleep(32);
    }
    return true;
}

// thread 0 of a boundary block, before any halo plane is read or any plane is stored into a neighbour
__device__ __forceinline__ void halo_flags_wait(const StepParams &p, int tile, bool first_chunk, bool last_chunk,
                                                unsigned long long seq)
{
    bool ok = true;
    const unsigned long long want = seq - 1;
    if (first_chunk && p.push_lo != nullptr) ok &= flag_wait(p.flags + (size_t)kFlagLo * p.ntiles + tile, want, p.timeout_cycles);
    if (last_chunk && p.push_hi != nullptr) ok &= flag_wait(p.flags + (size_t)kFlagHi * p.ntiles + tile, want, p.timeout_cycles);
    if (!ok) { p.state->error = 1; }
    fence_proxy_async_all();  // the planes are fetched by TMA (async proxy) after this generic-proxy acquire
}

// all threads of a boundary block, after its last plane (the stores into the neighbours precede the flags)
__device__ __forceinline__ void halo_flags_signal(const StepParams &p, int tile, bool first_chunk, bool last_chunk,
                                                  unsigned long long seq, int tid)
{
    __syncthreads();
    if (tid != 0) return;
    __threadfence_system();
    // to the low neighbour I am its HIGH neighbour: its plane nz-1 holds my tile, and I have read what it stored into my plane 0
    if (first_chunk && p.push_lo != nullptr) st_relaxed_sys_u64(p.lo_flags + (size_t)kFlagHi * p.ntiles + tile, seq);
    if (last_chunk && p.push_hi != nullptr) st_relaxed_sys_u64(p.hi_flags + (size_t)kFlagLo * p.ntiles + tile, seq);
}

// grid z index -> z chunk: one interior chunk first, then the two boundary chunks, then the rest. The boundary blocks
// start as soon as the first blocks retire (their planes reach the neighbour within the first quarter of the kernel), and
// their flag polls and system fences always overlap with an interior block resident on the same SM.
__device__ __forceinline__ int chunk_of_block(const StepParams &p, int bz, int nchunks)
{
    if (!p.flagged || nchunks < 3) return bz;
    return bz == 0 ? 1 : (bz == 1 ? 0 : (bz == 2 ? nchunks - 1 : bz - 1));
}

// Shared tail of both kernel variants: block partial -> deterministic grid sum -> (optionally) exit test / publish.
template <bool MULTI = true>
__device__ __forceinline__ void step_epilogue(const StepParams &p, double acc, double *red, unsigned long long seq)
{
    const int tid = threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
    const int nblocks = gridDim.x * gridDim.y * gridDim.z;
    const int bl = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
    double bsum = block_sum(acc, red);
    double total = 0.0;
    if (!grid_sum_last_block_all(bsum, p.partials, p.ticket, nblocks, bl, red, &total)) return;
    // ---- last block of the grid (all its threads; `total` is valid in thread 0) ----
    if (tid == 0) {
        if (p.sumsq_out != nullptr) *p.sumsq_out = total;
        if (MULTI && p.peer_slots != nullptr) {
            if (!p.flagged) {  // general decompositions: per-iteration handshake, consumed by pt_finalize_kernel
                const unsigned long long sq = p.state->seq;
                const int g = (int)(sq & (unsigned long long)(kSlotGens - 1));
                for (int r = 0; r < p.nranks; ++r) slot_publish(p.peer_slots[r], g, p.myrank, total, sq);
            }
        } else if (p.fuse_finalize) {
            pt_finalize(p.state, total, p.err_hist);
        }
    }
    if (!MULTI || !p.flagged) return;
    // publish this kernel's partial: thread r stores the two words into rank r's mailbox (N posted stores in parallel)
    __shared__ double total_bc;
    if (tid == 0) total_bc = total;
    __syncthreads();
    if (tid < p.nranks) slot_publish(p.peer_slots[tid], (int)(seq & (unsigned long long)(kSlotGens - 1)), p.myrank, total_bc, seq);
    // lagged evaluation: consume the partials of kernel seq-1 (published a whole kernel ago by every rank)
    __shared__ double vals[kMaxRanks];
    __shared__ int failed;
    PTState *s = p.state;
    const int pending = s->pending;  // uniform: only this block writes it, below
    if (tid == 0) failed = 0;
    __syncthreads();
    if (pending) {
        const unsigned long long want = seq - 1;
        const int g1 = (int)(want & (unsigned long long)(kSlotGens - 1));
        if (tid < p.nranks) {
            double v = 0.0;
            if (!slot_consume(p.my_slots, g1, tid, want, p.timeout_cycles, &v)) failed = 1;
            vals[tid] = v;
        }
        __syncthreads();
    }
    if (tid == 0) {
        if (pending) {
            if (failed) {
                s->error = 1;
                s->done = 1;
            } else {
                double t = 0.0;
                for (int r = 0; r < p.nranks; ++r) t += vals[r];  // MPI.Allreduce!(+) in rank order
                pt_finalize(s, t, p.err_hist);
                if (s->done) {  // the loop ended with kernel seq-1: this kernel was speculative and is discarded
                    s->pending = 0;
                    s->skip_push = p.consistent ? 0 : 1;
                }
            }
        } else {
            s->pending = 1;
            s->skip_push = 0;
        }
        s->seq = seq + 1;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Variant DIRECT: one thread per (x,y) column, z-marching with a register queue for the z neighbours; x/y neighbours
// come through L1/L2. Works for every size and alignment; the correctness anchor and the small-grid path.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kDirBX = 64, kDirBY = 4;

// Scalars of the array-programming arithmetic (part1_array_programming.jl:9-18), which divides by dx, dy, dz, dt where
// the kernel version multiplies by reciprocals: q = D*d(Htau)/dx, dHdtau = -(Htau - Ht)/dt + (d(qx)/dx + d(qy)/dy + d(qz)/dz),
// Htau += dHdtau*dtau. Evaluated in the reference's order; only the direct kernel has this mode (StepParams::array_arith).
struct ArrayArith {
    double D, dx, dy, dz, dt;
};

template <bool ARRAY>
__global__ void __launch_bounds__(kDirBX *kDirBY) step_direct_kernel(const StepParams p, const ArrayArith aa)
{
    __shared__ double red[32];
    if (p.state != nullptr && p.state->done) return;
    const unsigned long long seq = p.flagged ? p.state->seq : 0ull;
    const bool skip = p.flagged && p.state->skip_push;
    const int x = blockIdx.x * kDirBX + threadIdx.x;
    const int y = blockIdx.y * kDirBY + threadIdx.y;
    const int nchunks = gridDim.z;
    const int zs = 1 + chunk_of_block(p, blockIdx.z, nchunks) * p.zchunk;
    const int ze = min(zs + p.zchunk, p.nz - 1);
    const int tid = threadIdx.x + kDirBX * threadIdx.y;
    const int tile = blockIdx.x + gridDim.x * blockIdx.y;
    const bool first_chunk = zs == 1, last_chunk = ze == p.nz - 1;
    const bool boundary = p.flagged && ((first_chunk && p.push_lo != nullptr) || (last_chunk && p.push_hi != nullptr));
    if (boundary) {  // block-uniform
        if (tid == 0) halo_flags_wait(p, tile, first_chunk, last_chunk, seq);
        __syncthreads();
    }
    const bool inb = x < p.nx && y < p.ny;
    const bool valid = x >= 1 && x <= p.nx - 2 && y >= 1 && y <= p.ny - 2;
    const size_t sy = (size_t)p.nx, sz = (size_t)p.nx * p.ny;
    double acc = 0.0;
    if (inb) {
        size_t q = (size_t)x + sy * y + sz * zs;
        double zp = 0.0, c = 0.0;
        // halo planes are written by the neighbour GPUs: read them past L1 (__ldcg)
        if (valid) { zp = zs == 1 ? __ldcg(p.A + (q - sz)) : p.A[q - sz]; c = p.A[q]; }
        for (int z = zs; z < ze; ++z, q += sz) {
            const bool plo = p.push_lo != nullptr && z == 1;
            const bool phi = p.push_hi != nullptr && z == p.nz - 2;
            double outv = 0.0;
            if (plo || phi) outv = p.B[q];  // old content (also the value forwarded for boundary cells)
            if (valid) {
                const double zn = z == p.nz - 2 ? __ldcg(p.A + (q + sz)) : p.A[q + sz];
                double r;
                if (ARRAY) {
                    const double xl = p.A[q - 1], xr = p.A[q + 1], ys = p.A[q - sy], yn = p.A[q + sy];
                    const double dH = -(c - p.Ht[q]) / aa.dt + ((aa.D * (xr - c) / aa.dx - aa.D * (c - xl) / aa.dx) / aa.dx +
                                                                (aa.D * (yn - c) / aa.dy - aa.D * (c - ys) / aa.dy) / aa.dy +
                                                                (aa.D * (zn - c) / aa.dz - aa.D * (c - zp) / aa.dz) / aa.dz);
                    r = -dH;  // Htau + dHdtau*dtau == Htau - dtau*(-dHdtau) bit for bit; (-dH*dt)^2 == (dH*dt)^2
                } else {
                    r = cell_residual(c, p.A[q - 1], p.A[q + 1], p.A[q - sy], p.A[q + sy], zp, zn, p.Ht[q], p);
                }
                const double b = c - p.dtau * r;
                p.B[q] = b;
                if (p.R != nullptr) p.R[q] = r;
                const double v = r * p.norm_scale;
                acc += v * v;
                if (p.consistent) outv = b;
                zp = c;
                c = zn;
            }
            if (plo && !skip) p.push_lo[(size_t)x + sy * y] = outv;
            if (phi && !skip) p.push_hi[(size_t)x + sy * y] = outv;
        }
    }
    if (boundary) halo_flags_signal(p, tile, first_chunk, last_chunk, seq, tid);
    step_epilogue(p, acc, red, seq);
}

// ---------------------------------------------------------------------------------------------------------------
// Variant TMA: 2.5-D z-marching. A block owns an (TX x TY) xy-tile and a chunk of z planes. Planes of Htau (with a
// one-cell y halo and a two-cell, 16-byte aligned x halo) and of Ht are staged into a ring of S shared-memory stages
// by TMA (cp.async.bulk.tensor.3d, completion on an mbarrier per stage; out-of-range parts of a box are zero-filled
// by the hardware and masked in the arithmetic). Every thread owns two x-adjacent cells: the z neighbours live in a
// register queue, x/y neighbours are read from the staged plane, the result goes out as one coalesced 16-byte store.
// ---------------------------------------------------------------------------------------------------------------
template <int TX, int TY>
struct TmaCfg {
    static constexpr int BW = TX + 4;  // box width: cells X0-2 .. X0+TX+1
    static constexpr int BH = TY + 2;  // rows Y0-1 .. Y0+TY
    static constexpr int A_BYTES = BW * BH * 8;
    static constexpr int H_BYTES = TX * TY * 8;
    static constexpr int A_STRIDE = ((A_BYTES + 127) / 128) * 128 / 8;  // doubles, 128-byte aligned stages
    static constexpr int H_STRIDE = ((H_BYTES + 127) / 128) * 128 / 8;
    static constexpr int THREADS = (TX / 2) * TY;
    static constexpr size_t smem_bytes(int S) { return (size_t)S * (A_STRIDE + H_STRIDE) * 8 + (size_t)S * 8 + 128; }
};

// MULTI = false: the single-slab instantiation (no neighbours, no mailbox): all of the exchange code is compiled out.
template <int TX, int TY, int S, bool MULTI>
__global__ void __launch_bounds__((TX / 2) * TY)
    step_tma_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapHt, const StepParams p)
{
    using C = TmaCfg<TX, TY>;
    extern __shared__ __align__(128) unsigned char smem_dyn[];  // TMA destinations need 128-byte alignment
    __shared__ double red[32];
    if (p.state != nullptr && p.state->done) return;
    const bool flagged = MULTI && p.flagged;
    double *const push_lo = MULTI ? p.push_lo : nullptr, *const push_hi = MULTI ? p.push_hi : nullptr;
    const unsigned long long seq = flagged ? p.state->seq : 0ull;
    const bool skip = flagged && p.state->skip_push;

    // carve-up of the dynamic shared memory (kept in the shared state space: plain LDS/STS, no generic loads)
    double *sA = reinterpret_cast<double *>(smem_dyn);
    double *sH = sA + (size_t)S * C::A_STRIDE;
    uint64_t *full = (uint64_t *)(sH + (size_t)S * C::H_STRIDE);

    const int tid = threadIdx.x + (TX / 2) * threadIdx.y;
    const int X0 = blockIdx.x * TX, Y0 = blockIdx.y * TY;
    const int zs = 1 + (MULTI ? chunk_of_block(p, blockIdx.z, gridDim.z) : (int)blockIdx.z) * p.zchunk;
    const int ze = min(zs + p.zchunk, p.nz - 1);
    const int nplanes = (ze - zs) + 2;  // planes zs-1 .. ze
    const int tile = blockIdx.x + gridDim.x * blockIdx.y;
    const bool first_chunk = zs == 1, last_chunk = ze == p.nz - 1;
    const bool boundary = flagged && ((first_chunk && push_lo != nullptr) || (last_chunk && push_hi != nullptr));

    if (tid == 0) {
        for (int s = 0; s < S; ++s) mbar_init(&full[s], 1);
        fence_mbar_init();
        fence_proxy_async();
        if (boundary) halo_flags_wait(p, tile, first_chunk, last_chunk, seq);
    }
    __syncthreads();

    // Halo push, kept OUT of the plane loop (the loop is the single-slab kernel's, instruction for instruction).
    // lag-2 (reference) semantics forward the OLD content of Htau2's boundary planes: it is read and stored into the
    // neighbours right here, before the loop overwrites it -- the planes are on their way within microseconds of the
    // block's start. Consistent semantics forward the new values: after the loop every thread re-reads the two cells it
    // has just written (frame cells keep their old content, exactly what the in-loop version forwarded).
    auto push_planes = [&]() {
        const int x_ = X0 + 2 * (int)threadIdx.x, y_ = Y0 + (int)threadIdx.y;
        const bool i0 = x_ < p.nx && y_ < p.ny, i1 = x_ + 1 < p.nx && y_ < p.ny;
        const size_t pl = (size_t)p.nx * p.ny, o = (size_t)x_ + (size_t)p.nx * y_;
        if (first_chunk && push_lo != nullptr) {
            const double *src = p.B + pl;  // plane 1
            if (i0) push_lo[o] = src[o];
            if (i1) push_lo[o + 1] = src[o + 1];
        }
        if (last_chunk && push_hi != nullptr) {
            const double *src = p.B + pl * (size_t)(p.nz - 2);
            if (i0) push_hi[o] = src[o];
            if (i1) push_hi[o + 1] = src[o + 1];
        }
    };

    auto issue = [&](int q) {
        const int st = q % S;
        const int z = zs - 1 + q;
        const bool needH = (q >= 1 && q <= nplanes - 2);
        mbar_arrive_expect_tx(&full[st], (uint32_t)(C::A_BYTES + (needH ? C::H_BYTES : 0)));
        tma_load_3d(sA + (size_t)st * C::A_STRIDE, &mapA, &full[st], X0 - 2, Y0 - 1, z);
        if (needH) tma_load_3d(sH + (size_t)st * C::H_STRIDE, &mapHt, &full[st], X0, Y0, z);
    };
    if (tid == 0) {
        const int n0 = nplanes < S ? nplanes : S;
        for (int q = 0; q < n0; ++q) issue(q);
    }
    if (MULTI && boundary && !p.consistent && !skip) push_planes();  // while the first planes are in flight

    const int x = X0 + 2 * (int)threadIdx.x, y = Y0 + (int)threadIdx.y;
    const bool yok = y >= 1 && y <= p.ny - 2;
    const bool valid0 = yok && x >= 1 && x <= p.nx - 2;
    const bool valid1 = yok && x + 1 <= p.nx - 2;  // x+1 >= 1 always
    const bool inb0 = x < p.nx && y < p.ny, inb1 = x + 1 < p.nx && y < p.ny;
    const int ci = ((int)threadIdx.y + 1) * C::BW + 2 * (int)threadIdx.x + 2;  // centre pair inside a staged plane
    const int hi = (int)threadIdx.y * TX + 2 * (int)threadIdx.x;
    const size_t sy = (size_t)p.nx, sz = (size_t)p.nx * p.ny;
    const size_t pxy = (size_t)x + sy * y;

    double acc = 0.0;
    // plane zs-1: only its centre values are needed
    mbar_wait(&full[0], 0);
    double2 aprev = *reinterpret_cast<const double2 *>(sA + ci);
    __syncthreads();
    if (tid == 0 && S < nplanes) issue(S);
    mbar_wait(&full[1 % S], (uint32_t)((1 / S) & 1));
    double2 acur = *reinterpret_cast<const double2 *>(sA + (size_t)(1 % S) * C::A_STRIDE + ci);
    double qz_lo0 = p.mD_dz * (acur.x - aprev.x), qz_lo1 = p.mD_dz * (acur.y - aprev.y);

    for (int q = 2; q < nplanes; ++q) {
        const int z = zs + q - 2;
        const int stn = q % S, stc = (q - 1) % S;
        mbar_wait(&full[stn], (uint32_t)((q / S) & 1));
        const double2 anext = *reinterpret_cast<const double2 *>(sA + (size_t)stn * C::A_STRIDE + ci);
        const double *pc = sA + (size_t)stc * C::A_STRIDE + ci;
        const double xl = pc[-1], xr = pc[2];
        const double2 ysv = *reinterpret_cast<const double2 *>(pc - C::BW);
        const double2 ynv = *reinterpret_cast<const double2 *>(pc + C::BW);
        const double2 ht = *reinterpret_cast<const double2 *>(sH + (size_t)stc * C::H_STRIDE + hi);

        const double qx_mid = p.mD_dx * (acur.y - acur.x);  // between the two cells of this thread
        const double qz_hi0 = p.mD_dz * (anext.x - acur.x), qz_hi1 = p.mD_dz * (anext.y - acur.y);
        const double r0 = cell_residual_flux(p.mD_dx * (acur.x - xl), qx_mid, p.mD_dy * (acur.x - ysv.x), p.mD_dy * (ynv.x - acur.x),
                                             qz_lo0, qz_hi0, acur.x, ht.x, p);
        const double r1 = cell_residual_flux(qx_mid, p.mD_dx * (xr - acur.y), p.mD_dy * (acur.y - ysv.y), p.mD_dy * (ynv.y - acur.y),
                                             qz_lo1, qz_hi1, acur.y, ht.y, p);
        qz_lo0 = qz_hi0; qz_lo1 = qz_hi1;
        const double b0 = acur.x - p.dtau * r0;
        const double b1 = acur.y - p.dtau * r1;
        const size_t g = pxy + sz * z;
        if (valid0 && valid1) {
            *reinterpret_cast<double2 *>(p.B + g) = make_double2(b0, b1);
            if (p.R != nullptr) *reinterpret_cast<double2 *>(p.R + g) = make_double2(r0, r1);
        } else {
            if (valid0) { p.B[g] = b0; if (p.R != nullptr) p.R[g] = r0; }
            if (valid1) { p.B[g + 1] = b1; if (p.R != nullptr) p.R[g + 1] = r1; }
        }
        if (valid0) { const double v = r0 * p.norm_scale; acc += v * v; }
        if (valid1) { const double v = r1 * p.norm_scale; acc += v * v; }

        __syncthreads();  // every thread is done with stage stc -> refill it
        if (tid == 0 && q - 1 + S < nplanes) issue(q - 1 + S);
        aprev = acur;
        acur = anext;
    }
    if (MULTI && boundary && p.consistent) push_planes();
    if (boundary) halo_flags_signal(p, tile, first_chunk, last_chunk, seq, tid);
    step_epilogue<MULTI>(p, acc, red, seq);
}

// update_halo! of a general Cartesian decomposition (ImplicitGlobalGrid): one plane of `src` (index sp along `axis`) is
// copied to plane dp of `dst` -- whole planes including their edges, so values travel across corners through the
// x -> y -> z sequence exactly as in the reference. src and dst may live on different GPUs (peer access).
__global__ void __launch_bounds__(256) halo_plane_copy_kernel(const double *__restrict__ src, double *__restrict__ dst, int axis,
                                                              int sp, int dp, int nx, int ny, int nz, const PTState *state)
{
    if (state != nullptr && state->done) return;
    const int n1 = axis == 0 ? ny : nx;            // fast index of the plane
    const int n2 = axis == 2 ? ny : nz;            // slow index of the plane
    const size_t total = (size_t)n1 * n2;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        const int a = (int)(t % n1), b = (int)(t / n1);
        size_t ps, pd;
        if (axis == 0) { ps = (size_t)sp + (size_t)nx * (a + (size_t)ny * b); pd = (size_t)dp + (size_t)nx * (a + (size_t)ny * b); }
        else if (axis == 1) { ps = (size_t)a + (size_t)nx * (sp + (size_t)ny * b); pd = (size_t)a + (size_t)nx * (dp + (size_t)ny * b); }
        else { ps = (size_t)a + (size_t)nx * (b + (size_t)ny * sp); pd = (size_t)a + (size_t)nx * (b + (size_t)ny * dp); }
        dst[pd] = src[ps];
    }
}

// Barrier between the ranks of a general decomposition hosted by different processes: announces "this rank has finished
// everything before barrier k of the current PT iteration" in every rank's mailbox and waits until all ranks have done
// the same. The phase number is derived from the device-resident sequence number, which advances identically on all
// ranks (the host-side launch counts may differ after convergence). One thread per rank.
__global__ void cart_barrier_kernel(PTState *state, CartSync *mine, CartSync *const *peers, int nranks, int myrank, int k,
                                    long long timeout_cycles)
{
    if (state->done) return;
    const int t = threadIdx.x;
    const unsigned long long phase = 4ull * state->seq + (unsigned long long)k;
    __shared__ int failed;
    if (t == 0) failed = 0;
    __syncthreads();
    if (t < nranks) {
        __threadfence_system();  // the plane copies / the step kernel of this rank are visible before the flag
        st_release_sys_u64(&peers[t]->phase[myrank], phase);
        const long long t0 = clock64();
        while (ld_acquire_sys_u64(&mine->phase[t]) < phase) {
            if (clock64() - t0 > timeout_cycles) { failed = 1; break; }
            __nanosleep(64);
        }
    }
    __syncthreads();
    if (t == 0 && failed) { state->error = 1; state->done = 1; }
}

// One block: consume the partial sums of all ranks in rank order (deterministic, identical on every GPU).
// In-process handles pass `local` (nranks contiguous doubles on this device); otherwise `slots` is waited on (bounded).
// lagged = 0 (general decompositions): once per PT iteration, consumes the partials of kernel `seq` and advances seq.
// lagged = 1 (z-slab stacks): once per host batch, evaluates the one iteration the step kernels have left pending.
__global__ void pt_finalize_kernel(PTState *state, double *err_hist, const double *local, RankSlots *slots, int nranks,
                                   long long timeout_cycles, int lagged)
{
    if (state->done) return;
    if (lagged && !state->pending) return;
    __shared__ double vals[kMaxRanks];
    __shared__ int failed;
    const int t = threadIdx.x;
    if (t == 0) failed = 0;
    __syncthreads();
    if (slots != nullptr) {
        const unsigned long long seq = lagged ? state->seq - 1 : state->seq;
        const int g = (int)(seq & (unsigned long long)(kSlotGens - 1));
        if (t < nranks) {
            double v = 0.0;
            if (!slot_consume(slots, g, t, seq, timeout_cycles, &v)) failed = 1;
            vals[t] = v;
        }
    } else if (t < nranks) {
        vals[t] = local[t];
    }
    __syncthreads();
    if (t == 0) {
        if (failed) {
            state->error = 1;
            state->done = 1;
        } else {
            double total = 0.0;
            for (int r = 0; r < nranks; ++r) total += vals[r];  // MPI.Allreduce!(+) in rank order
            if (lagged) state->pending = 0;
            else state->seq += 1;
            pt_finalize(state, total, err_hist);
        }
    }
}

}  // namespace b2s
