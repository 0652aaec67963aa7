// Error string, SM count, TMA descriptor encoding.
#include "mtn_host.h"
#include <stdarg.h>
#include <string.h>

namespace mtn {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int num_sms() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

int ensure_dyn_smem(const void* kern, int bytes, std::atomic<unsigned long long>& done_mask, const char* what) {
    int dev = -1;
    if (cudaGetDevice(&dev) != cudaSuccess) dev = -1;
    const bool tracked = dev >= 0 && dev < 64;
    if (tracked && ((done_mask.load(std::memory_order_acquire) >> dev) & 1ull)) return MTN_OK;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) {
        set_error("%s: cudaFuncSetAttribute(%d B dynamic smem) failed on device %d: %s", what, bytes, dev,
                  cudaGetErrorString(e));
        return MTN_ECUDA;
    }
    if (tracked) done_mask.fetch_or(1ull << dev, std::memory_order_release);
    return MTN_OK;
}

typedef CUresult (*encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static encode_tiled_fn get_encode() {
    static encode_tiled_fn fn = nullptr;
    if (fn) return fn;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess || !p) {
        set_error("cuTensorMapEncodeTiled entry point not available");
        return nullptr;
    }
    fn = reinterpret_cast<encode_tiled_fn>(p);
    return fn;
}

bool encode_tmap(CUtensorMap* map, CUtensorMapDataType dtype, int rank, const void* base, const uint64_t* dims,
                 const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle swizzle,
                 CUtensorMapL2promotion l2_promotion) {
    encode_tiled_fn fn = get_encode();
    if (!fn) return false;
    cuuint64_t gdim[5];
    cuuint64_t gstr[4];
    cuuint32_t bx[5];
    cuuint32_t es[5];
    for (int i = 0; i < rank; ++i) {
        gdim[i] = dims[i];
        bx[i] = box[i];
        es[i] = 1;
    }
    for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
    CUresult r = fn(map, dtype, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bx, es,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, l2_promotion,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (CUresult %d): rank %d dims [%llu,%llu,%llu] box [%u,%u,%u]", (int)r,
                  rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
                  (unsigned long long)(rank > 2 ? dims[2] : 0), box[0], rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0);
        return false;
    }
    return true;
}

}  // namespace mtn

extern "C" const char* mtn_last_error_string(void) { return mtn::g_err; }
#ifdef MTN_SCAN_DEV
extern "C" int mtn_abi_version(void) { return MTN_ABI_VERSION + 1000; }   // tools/devbuild.sh experiment build
#else
extern "C" int mtn_abi_version(void) { return MTN_ABI_VERSION; }
#endif
extern "C" size_t mtn_sizeof_gemm_args(void) { return sizeof(mtn_gemm_args); }
extern "C" size_t mtn_sizeof_scan_args(void) { return sizeof(mtn_scan_args); }
extern "C" size_t mtn_sizeof_gn_apply_args(void) { return sizeof(mtn_gn_apply_args); }
