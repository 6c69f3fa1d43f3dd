// Library-wide host glue: error string, device validation, tensor-map encoding.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <cudaTypedefs.h>

#include "common.h"

namespace mfb {

static thread_local char g_err[512] = "";
static int g_device = -1;
static int g_sms = 0;
static PFN_cuTensorMapEncodeTiled_v12000 g_encode = nullptr;

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int device_sm_count() { return g_sms; }

static thread_local Program* g_recording = nullptr;
Program* recording() { return g_recording; }

__global__ void copy32_kernel(float* __restrict__ dst, const float* __restrict__ src, long long n) {
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * blockDim.x)
        dst[i] = src[i];
}

bool pdl_enabled() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("MFB_PDL");
        v = (e && e[0] == '1') ? 1 : 0;   // opt-in: measured neutral-to-slightly-negative inside the CUDA graph (profiles/r01f)
    }
    return v == 1;
}

int encode_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                     const uint32_t* box, int swizzle_bytes) {
    if (!g_encode) {
        set_error("mfb_init() has not been called (cuTensorMapEncodeTiled unresolved)");
        return MFB_EINVAL;
    }
    cuuint64_t gdim[5];
    cuuint64_t gstr[4];
    cuuint32_t bdim[5];
    cuuint32_t estr[5];
    for (int i = 0; i < rank; ++i) {
        gdim[i] = dims[i];
        bdim[i] = box[i];
        estr[i] = 1;
        if (i > 0) gstr[i - 1] = strides_bytes[i - 1];
    }
    if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) {
        set_error("tensor base %p is not 16-byte aligned", base);
        return MFB_EINVAL;
    }
    CUresult r = g_encode(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(base), gdim, gstr, bdim, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE,
                          swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                          : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (CUresult %d) rank %d dims [%llu,%llu,%llu,%llu] box [%u,%u,%u,%u]", int(r), rank,
                  (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
                  (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0), box[0],
                  rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0);
        return MFB_ECUDA;
    }
    return MFB_OK;
}

}  // namespace mfb

using namespace mfb;

extern "C" int mfb_abi_version(void) { return MFB_ABI_VERSION; }

extern "C" const char* mfb_last_error(void) { return g_err; }

extern "C" int mfb_init(int device) {
    if (g_device == device && g_encode) return MFB_OK;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        set_error("no CUDA device visible (%s); mirrorfusion_b200 has no CPU fallback", cudaGetErrorString(e));
        return MFB_ECUDA;
    }
    MFB_REQUIRE(device >= 0 && device < n, "device %d out of range (%d visible)", device, n);
    cudaDeviceProp prop;
    MFB_CUDA_OK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        set_error("device %d is sm_%d%d; libmfb200 is built for sm_100a (B200) only", device, prop.major, prop.minor);
        return MFB_EUNSUPPORTED;
    }
    MFB_CUDA_OK(cudaSetDevice(device));
    g_sms = prop.multiProcessorCount;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    MFB_CUDA_OK(cudaGetDriverEntryPointByVersion("cuTensorMapEncodeTiled", &fn, 12000, cudaEnableDefault, &q));
    if (q != cudaDriverEntryPointSuccess || !fn) {
        set_error("cuTensorMapEncodeTiled not available from the driver");
        return MFB_ECUDA;
    }
    g_encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
    g_device = device;
    return MFB_OK;
}

// ---------------------------------------------------------------------------------------------- recorded launch programs
extern "C" int mfb_program_begin(mfb_program** out) {
    MFB_REQUIRE(out, "null argument");
    MFB_REQUIRE(g_recording == nullptr, "a program is already being recorded on this thread");
    g_recording = new Program();
    *out = reinterpret_cast<mfb_program*>(g_recording);
    return MFB_OK;
}

extern "C" int mfb_program_end(void) {
    MFB_REQUIRE(g_recording != nullptr, "no program is being recorded on this thread");
    g_recording = nullptr;
    return MFB_OK;
}

extern "C" int mfb_program_size(const mfb_program* prog) {
    return prog ? static_cast<int>(reinterpret_cast<const Program*>(prog)->ops.size()) : 0;
}

extern "C" int mfb_program_run(mfb_program* prog, void* stream) {
    MFB_REQUIRE(prog, "null program");
    Program* p = reinterpret_cast<Program*>(prog);
    MFB_REQUIRE(p != g_recording, "the program is still being recorded (call mfb_program_end first)");
    for (auto& op : p->ops) {
        const int rc = op(stream);
        if (rc) return rc;
    }
    return MFB_OK;
}

extern "C" int mfb_program_destroy(mfb_program* prog) {
    Program* p = reinterpret_cast<Program*>(prog);
    if (p && p == g_recording) g_recording = nullptr;
    delete p;
    return MFB_OK;
}

extern "C" int mfb_copy_f32(float* dst, const float* src, long long n, void* stream) {
    MFB_RECORD(mfb_copy_f32(dst, src, n, stream));
    MFB_REQUIRE(dst && src && n > 0, "bad arguments");
    long long blocks = (n + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    copy32_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(dst, src, n);
    MFB_CUDA_OK(cudaGetLastError());
    return MFB_OK;
}
