// common.cuh -- error plumbing, deterministic reductions, mbarrier/TMA PTX wrappers (sm_100a).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <string.h>

#include "../../include/b200stencil.h"

namespace b2s {

// ---- host-side error plumbing -------------------------------------------------------------------------
void set_error(const char *fmt, ...);
const char *get_error();

#define B2S_CUDA(call)                                                                                      \
    do {                                                                                                    \
        cudaError_t e__ = (call);                                                                           \
        if (e__ != cudaSuccess) {                                                                           \
            b2s::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__));          \
            return (e__ == cudaErrorNoDevice || e__ == cudaErrorInsufficientDriver) ? B2S_ERR_NO_DEVICE      \
                                                                                      : B2S_ERR_CUDA;       \
        }                                                                                                   \
    } while (0)

#define B2S_CHECK(call)                  \
    do {                                 \
        int rc__ = (call);               \
        if (rc__ != B2S_OK) return rc__; \
    } while (0)

#define B2S_REQUIRE(cond, code, ...)     \
    do {                                 \
        if (!(cond)) {                   \
            b2s::set_error(__VA_ARGS__); \
            return (code);               \
        }                                \
    } while (0)

// Library-owned per-device scratch for L0 reductions (lazily allocated, freed by b2s_shutdown()).
struct Scratch {
    double *partials = nullptr;   // kMaxPartials doubles
    unsigned int *ticket = nullptr;
    double *result = nullptr;     // device scalar(s)
    double *pinned = nullptr;     // host pinned scalar(s)
};
constexpr int kMaxPartials = 1 << 16;
int get_scratch(Scratch **out);  // for the current device
int free_all_scratch();

struct DeviceGuard {
    int prev = -1;
    bool active = false;
    int set(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; }
        if (prev != dev) {
            cudaError_t e = cudaSetDevice(dev);
            if (e != cudaSuccess) { set_error("cudaSetDevice(%d): %s", dev, cudaGetErrorString(e)); return B2S_ERR_CUDA; }
            active = true;
        }
        return B2S_OK;
    }
    ~DeviceGuard() { if (active && prev >= 0) cudaSetDevice(prev); }
};

// ---- device-side helpers ------------------------------------------------------------------------------
#ifdef __CUDACC__

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block-wide sum with a fixed reduction tree (deterministic for a given block size). Result valid in thread 0.
// `sm` must hold >= 32 doubles. All threads of the block must call.
__device__ __forceinline__ double block_sum(double v, double *sm)
{
    const int tid = threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
    const int nthreads = blockDim.x * blockDim.y * blockDim.z;
    const int lane = tid & 31, warp = tid >> 5;
    v = warp_sum(v);
    __syncthreads();  // protect sm reuse across calls
    if (lane == 0) sm[warp] = v;
    __syncthreads();
    double r = 0.0;
    if (warp == 0) {
        const int nw = (nthreads + 31) >> 5;
        r = lane < nw ? sm[lane] : 0.0;
        r = warp_sum(r);
    }
    return r;
}

// Two-stage deterministic grid reduction: every block deposits its partial; the last block to arrive sums all
// partials in a fixed order (strided per thread, then the fixed block tree) independent of arrival order.
// Returns true in thread 0 of the last block, with the total in *total.
__device__ __forceinline__ bool grid_sum_last_block(double block_partial_thread0, double *partials, unsigned int *ticket,
                                                    int nblocks, int block_linear, double *sm, double *total)
{
    __shared__ bool is_last;
    const int tid = threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
    const int nthreads = blockDim.x * blockDim.y * blockDim.z;
    if (tid == 0) {
        partials[block_linear] = block_partial_thread0;
        __threadfence();
        unsigned int t = atomicAdd(ticket, 1u);
        is_last = (t == (unsigned int)(nblocks - 1));
    }
    __syncthreads();
    if (!is_last) return false;
    __threadfence();
    double a = 0.0;
    for (int i = tid; i < nblocks; i += nthreads) a += __ldcg(partials + i);
    double t = block_sum(a, sm);
    if (tid == 0) {
        *total = t;
        *ticket = 0u;  // ready for the next launch
    }
    return tid == 0;
}

// Same reduction, but the return value (is this the last block?) is valid in EVERY thread of the block, so the whole last
// block can take part in what follows; *total is written by thread 0 only.
__device__ __forceinline__ bool grid_sum_last_block_all(double block_partial_thread0, double *partials, unsigned int *ticket,
                                                        int nblocks, int block_linear, double *sm, double *total)
{
    __shared__ bool is_last_all;
    const int tid = threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
    const int nthreads = blockDim.x * blockDim.y * blockDim.z;
    if (tid == 0) {
        partials[block_linear] = block_partial_thread0;
        __threadfence();
        unsigned int t = atomicAdd(ticket, 1u);
        is_last_all = (t == (unsigned int)(nblocks - 1));
    }
    __syncthreads();
    if (!is_last_all) return false;
    __threadfence();
    double a = 0.0;
    for (int i = tid; i < nblocks; i += nthreads) a += __ldcg(partials + i);
    double t = block_sum(a, sm);
    if (tid == 0) {
        *total = t;
        *ticket = 0u;  // ready for the next launch
    }
    return true;
}

// ---- mbarrier + TMA (cp.async.bulk.tensor) -------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
// generic-proxy accesses (e.g. an acquire of a flag another GPU has released) before later async-proxy (TMA) accesses,
// all state spaces
__device__ __forceinline__ void fence_proxy_async_all()
{
    asm volatile("fence.proxy.async;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    const uint32_t addr = smem_u32(bar);
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t"
            ".reg .pred P1;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, P1;\n\t"
            "}"
            : "=r"(ok)
            : "r"(addr), "r"(parity)
            : "memory");
    } while (!ok);
}
// 3-D tiled TMA load: box -> shared memory, completion counted on the mbarrier.
__device__ __forceinline__ void tma_load_3d(void *smem_dst, const CUtensorMap *tmap, uint64_t *bar, int c0, int c1, int c2)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"((uint64_t)tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap *tmap)
{
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)tmap) : "memory");
}

// system-scope release/acquire for cross-GPU flags
__device__ __forceinline__ void st_release_sys_u64(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_sys_u64(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys_u64(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ double ld_relaxed_sys_f64(const double *p)
{
    double v;
    asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}
#endif  // __CUDACC__

// Host: encode a 3-D Float64 tiled tensor map (driver entry point resolved at run time; no libcuda link).
int make_tensor_map_3d(CUtensorMap *out, const double *base, uint64_t d0, uint64_t d1, uint64_t d2, uint32_t b0,
                       uint32_t b1, uint32_t b2);

}  // namespace b2s
