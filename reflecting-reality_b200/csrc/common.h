// Host-side helpers shared by the translation units of libmfb200.so.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/mfb200.h"

#include <functional>
#include <vector>

namespace mfb {

// A recorded launch program (mfb_program_* in mfb200.h): while a recorder is active on the calling thread, every step-level entry
// point appends a closure of itself — its arguments by value, the stream left open — before executing; mfb_program_run replays the
// closures in order on the stream it is given.  Buffers and plans referenced by the recorded calls must outlive the program (the
// engines allocate everything once per geometry, which is what CUDA-graph capture needs as well).
struct Program {
    std::vector<std::function<int(void*)>> ops;
};
Program* recording();
#define MFB_RECORD(...)                                                                                   \
    do {                                                                                                  \
        if (::mfb::Program* _rec = ::mfb::recording())                                                    \
            _rec->ops.emplace_back([=](void* stream) -> int { return __VA_ARGS__; });                    \
    } while (0)

void set_error(const char* fmt, ...);
int device_sm_count();

#define MFB_CUDA_OK(expr)                                                                           \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess) {                                                                    \
            ::mfb::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return MFB_ECUDA;                                                                       \
        }                                                                                           \
    } while (0)

#define MFB_REQUIRE(cond, ...)                \
    do {                                      \
        if (!(cond)) {                        \
            ::mfb::set_error(__VA_ARGS__);    \
            return MFB_EINVAL;                \
        }                                     \
    } while (0)

// Encode a tiled bf16 tensor map (rank <= 5).  dims/box are innermost-first; strides_bytes has rank-1 entries
// (stride of dims 1..rank-1).  swizzle_bytes: 0 (none), 32, 64 or 128 — the inner box extent must span exactly
// that many bytes (16 / 32 / 64 bf16 elements).
int encode_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                     const uint32_t* box, int swizzle_bytes);

// Launch helper: programmatic stream serialization (PDL) when MFB_PDL=1; cluster_x > 1 adds a cluster dimension.
bool pdl_enabled();

template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, int cluster_x,
                            Args&&... args) {
    cudaLaunchConfig_t cfg;
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    int n = 0;
    if (cluster_x > 1) {
        attr[n].id = cudaLaunchAttributeClusterDimension;
        attr[n].val.clusterDim.x = cluster_x;
        attr[n].val.clusterDim.y = 1;
        attr[n].val.clusterDim.z = 1;
        ++n;
    }
    if (pdl_enabled()) {
        attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[n].val.programmaticStreamSerializationAllowed = 1;
        ++n;
    }
    cfg.attrs = attr;
    cfg.numAttrs = n;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

}  // namespace mfb
