// common.cu -- error plumbing, per-device scratch, tensor-map encoding, library-level entry points.
#include "common.cuh"

#include <mutex>

namespace b2s {

static thread_local char g_err[1024] = "";

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
const char *get_error() { return g_err; }

constexpr int kMaxDevices = 64;
static Scratch g_scratch[kMaxDevices];
static bool g_scratch_ok[kMaxDevices];
static std::mutex g_scratch_mu;

int get_scratch(Scratch **out)
{
    int dev = 0;
    B2S_CUDA(cudaGetDevice(&dev));
    B2S_REQUIRE(dev >= 0 && dev < kMaxDevices, B2S_ERR_BAD_ARG, "device ordinal %d out of range", dev);
    std::lock_guard<std::mutex> lk(g_scratch_mu);
    Scratch &s = g_scratch[dev];
    if (!g_scratch_ok[dev]) {
        B2S_CUDA(cudaMalloc(&s.partials, sizeof(double) * kMaxPartials));
        B2S_CUDA(cudaMalloc(&s.ticket, sizeof(unsigned int) * 4));
        B2S_CUDA(cudaMemset(s.ticket, 0, sizeof(unsigned int) * 4));
        B2S_CUDA(cudaMalloc(&s.result, sizeof(double) * 16));
        B2S_CUDA(cudaMemset(s.result, 0, sizeof(double) * 16));
        B2S_CUDA(cudaMallocHost(&s.pinned, sizeof(double) * 16));
        g_scratch_ok[dev] = true;
    }
    *out = &s;
    return B2S_OK;
}

int free_all_scratch()
{
    std::lock_guard<std::mutex> lk(g_scratch_mu);
    int prev = -1;
    cudaGetDevice(&prev);
    for (int d = 0; d < kMaxDevices; ++d) {
        if (!g_scratch_ok[d]) continue;
        cudaSetDevice(d);
        cudaFree(g_scratch[d].partials);
        cudaFree(g_scratch[d].ticket);
        cudaFree(g_scratch[d].result);
        cudaFreeHost(g_scratch[d].pinned);
        g_scratch[d] = Scratch();
        g_scratch_ok[d] = false;
    }
    if (prev >= 0) cudaSetDevice(prev);
    return B2S_OK;
}

// cuTensorMapEncodeTiled resolved through the runtime so the library needs no link-time libcuda.
typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode()
{
    static PFN_encodeTiled fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (PFN_encodeTiled)p;
    }
    return fn;
}

int make_tensor_map_3d(CUtensorMap *out, const double *base, uint64_t d0, uint64_t d1, uint64_t d2, uint32_t b0,
                       uint32_t b1, uint32_t b2)
{
    PFN_encodeTiled enc = get_encode();
    B2S_REQUIRE(enc != nullptr, B2S_ERR_CUDA, "cuTensorMapEncodeTiled driver entry point unavailable");
    B2S_REQUIRE(((uintptr_t)base & 15) == 0 && (d0 * 8) % 16 == 0 && (b0 * 8) % 16 == 0 && b0 <= 256 && b1 <= 256 &&
                    b2 <= 256,
                B2S_ERR_BAD_ARG, "tensor map constraints violated (base %p, d0 %llu, box %u %u %u)", (const void *)base,
                (unsigned long long)d0, b0, b1, b2);
    cuuint64_t dims[3] = {d0, d1, d2};
    cuuint64_t strides[2] = {d0 * 8, d0 * d1 * 8};
    cuuint32_t box[3] = {b0, b1, b2};
    cuuint32_t estr[3] = {1, 1, 1};
    // FLOAT32 pairs would also work, but FLOAT64 keeps the coordinates in elements of the field.
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, (void *)base, dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    B2S_REQUIRE(r == CUDA_SUCCESS, B2S_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return B2S_OK;
}

}  // namespace b2s

extern "C" {

const char *b2s_last_error(void) { return b2s::get_error(); }
int b2s_version(void) { return B2S_VERSION; }

int b2s_device_count(int *count)
{
    B2S_REQUIRE(count != nullptr, B2S_ERR_BAD_ARG, "count is NULL");
    *count = 0;
    cudaError_t e = cudaGetDeviceCount(count);
    if (e != cudaSuccess) {
        b2s::set_error("cudaGetDeviceCount: %s", cudaGetErrorString(e));
        *count = 0;
        return B2S_ERR_NO_DEVICE;
    }
    return B2S_OK;
}

int b2s_shutdown(void) { return b2s::free_all_scratch(); }

}  // extern "C"
