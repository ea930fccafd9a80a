// hd_common.cu -- host-side helpers: error string, driver entry point for TMA descriptor encoding.
#include "hd_common.cuh"

#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace hd {

static thread_local char g_err[512] = "";

bool pdl_enabled() {
    static int on = -1;
    if (on < 0) {
        const char* e = getenv("HD_PDL");
        on = (e == nullptr || e[0] != '0') ? 1 : 0;
    }
    return on == 1;
}

void set_last_error(const char* file, int line, const char* msg) {
    const char* base = strrchr(file, '/');
    snprintf(g_err, sizeof(g_err), "%s:%d: %s", base ? base + 1 : file, line, msg);
}

typedef CUresult (*encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static encode_tiled_fn get_encode() {
    // resolved through the runtime so that the library itself does not link libcuda (loadable on a CPU-only host)
    static encode_tiled_fn fn = nullptr;
    if (fn == nullptr) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<encode_tiled_fn>(p);
    }
    return fn;
}

int make_tensor_map(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                    const uint32_t* box, int swizzle_bytes) {
    encode_tiled_fn enc = get_encode();
    if (enc == nullptr) {
        set_last_error(__FILE__, __LINE__, "cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
        return HD_ERR_CUDA;
    }
    cuuint64_t gdim[5], gstr[4];
    cuuint32_t bdim[5], estr[5];
    for (int i = 0; i < rank; ++i) {
        gdim[i] = dims[i];
        bdim[i] = box[i];
        estr[i] = 1;
    }
    for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
    CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                            : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                            : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                                                  : CU_TENSOR_MAP_SWIZZLE_NONE;
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, static_cast<cuuint32_t>(rank), const_cast<void*>(base), gdim,
                     gstr, bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        char msg[256];
        snprintf(msg, sizeof(msg),
                 "cuTensorMapEncodeTiled failed (%d): rank %d dims [%llu %llu %llu %llu %llu] box [%u %u %u %u %u] swz %d",
                 static_cast<int>(r), rank, (unsigned long long)gdim[0], (unsigned long long)(rank > 1 ? gdim[1] : 0),
                 (unsigned long long)(rank > 2 ? gdim[2] : 0), (unsigned long long)(rank > 3 ? gdim[3] : 0),
                 (unsigned long long)(rank > 4 ? gdim[4] : 0), bdim[0], rank > 1 ? bdim[1] : 0, rank > 2 ? bdim[2] : 0,
                 rank > 3 ? bdim[3] : 0, rank > 4 ? bdim[4] : 0, swizzle_bytes);
        set_last_error(__FILE__, __LINE__, msg);
        return HD_ERR_CUDA;
    }
    return HD_OK;
}

}  // namespace hd

extern "C" int hd_version(void) { return 1; }

extern "C" const char* hd_last_error(void) { return hd::g_err; }

extern "C" int hd_device_ok(void) {
    int dev = 0;
    cudaDeviceProp prop;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&prop, dev) != cudaSuccess) {
        hd::set_last_error(__FILE__, __LINE__, "no usable CUDA device");
        return HD_ERR_CUDA;
    }
    if (prop.major != 10) {
        hd::set_last_error(__FILE__, __LINE__, "device is not sm_100 (B200)");
        return HD_ERR_CUDA;
    }
    return HD_OK;
}
