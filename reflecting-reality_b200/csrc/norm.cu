// GroupNorm(+SiLU) and LayerNorm over channels-last bf16 activations, fp32 statistics.  HBM-bound kernels:
// every access is a 16-byte vector, a thread keeps a fixed 8-channel slice so per-channel scale/shift live in
// registers, and the up-block skip concat is read from its two source tensors directly.
#include <stdlib.h>
#include "common.h"
#include "ptx.cuh"

namespace mfb {

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
    float2 t;
    t = unpack_bf16x2(u.x); f[0] = t.x; f[1] = t.y;
    t = unpack_bf16x2(u.y); f[2] = t.x; f[3] = t.y;
    t = unpack_bf16x2(u.z); f[4] = t.x; f[5] = t.y;
    t = unpack_bf16x2(u.w); f[6] = t.x; f[7] = t.y;
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
    uint4 o;
    o.x = pack_bf16x2(f[0], f[1]);
    o.y = pack_bf16x2(f[2], f[3]);
    o.z = pack_bf16x2(f[4], f[5]);
    o.w = pack_bf16x2(f[6], f[7]);
    return o;
}

// grid (chunks, B), block (CV = C/8, PY).  Deterministic: every CTA writes its per-group partial {sum, sumsq} to
// part[b][chunk][g], the last CTA of a batch image to finish (ticket counter) adds the partials in chunk order into
// stats[b][g] and re-arms the counter — no floating-point atomics on global memory, no memset between calls.
__global__ void gn_stats_kernel(const __nv_bfloat16* __restrict__ x1, int C1, const __nv_bfloat16* __restrict__ x2, int C2,
                                int HW, int groups, int pix_per_cta, float* __restrict__ stats, float* __restrict__ part,
                                unsigned int* __restrict__ counters) {
    __shared__ float gs[64], gq[64];
    __shared__ float4 ps[1024];
    __shared__ unsigned int s_ticket;
    const int C = C1 + C2;
    const int cpg = C / groups;
    const int b = blockIdx.y;
    const int c0 = threadIdx.x * 8;
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    pdl_trigger();
    pdl_wait();
    const __nv_bfloat16* src;
    int ld, cc;
    if (c0 < C1) { src = x1; ld = C1; cc = c0; } else { src = x2; ld = C2; cc = c0 - C1; }
    src += static_cast<size_t>(b) * HW * ld + cc;
    const int p0 = blockIdx.x * pix_per_cta;
    const int p1 = min(HW, p0 + pix_per_cta);
    float s[8], q[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) { s[e] = 0.f; q[e] = 0.f; }
    // 4 independent 16-byte loads in flight per thread (HBM-bound: memory-level parallelism is what matters)
    const int step = blockDim.y;
    int p = p0 + threadIdx.y;
    for (; p + 7 * step < p1; p += 8 * step) {
        uint4 u[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) u[k] = __ldg(reinterpret_cast<const uint4*>(src + static_cast<size_t>(p + k * step) * ld));
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            float f[8];
            unpack8(u[k], f);
#pragma unroll
            for (int e = 0; e < 8; ++e) { s[e] += f[e]; q[e] = fmaf(f[e], f[e], q[e]); }
        }
    }
    for (; p + 3 * step < p1; p += 4 * step) {
        uint4 u[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) u[k] = __ldg(reinterpret_cast<const uint4*>(src + static_cast<size_t>(p + k * step) * ld));
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float f[8];
            unpack8(u[k], f);
#pragma unroll
            for (int e = 0; e < 8; ++e) { s[e] += f[e]; q[e] = fmaf(f[e], f[e], q[e]); }
        }
    }
    for (; p < p1; p += step) {
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(src + static_cast<size_t>(p) * ld));
        float f[8];
        unpack8(u, f);
#pragma unroll
        for (int e = 0; e < 8; ++e) { s[e] += f[e]; q[e] = fmaf(f[e], f[e], q[e]); }
    }
    // 8 consecutive channels touch at most two groups (cpg >= 8).  Fixed-order (deterministic) CTA reduction:
    // every thread parks its two partial pairs in shared memory, then thread g sums group g over all threads.
    const int g0 = c0 / cpg;
    float sa = 0.f, qa = 0.f, sb = 0.f, qb = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        if ((c0 + e) / cpg == g0) { sa += s[e]; qa += q[e]; } else { sb += s[e]; qb += q[e]; }
    }
    // level 1: over the pixel lanes (ty) of each channel vector, in ty order
    ps[tid] = make_float4(sa, qa, sb, qb);
    __syncthreads();
    const int CV = blockDim.x;
    if (tid < CV) {
        float4 acc = ps[tid];
        for (int y = 1; y < blockDim.y; ++y) {
            const float4 v = ps[y * CV + tid];
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        ps[tid] = acc;
    }
    __syncthreads();
    // level 2: thread g sums the (few) channel vectors that touch group g, in channel order
    if (tid < groups) {
        float a = 0.f, q2 = 0.f;
        int v0 = (tid * cpg) / 8 - 1;
        if (v0 < 0) v0 = 0;
        int v1 = ((tid + 1) * cpg - 1) / 8;
        if (v1 > CV - 1) v1 = CV - 1;
        for (int v = v0; v <= v1; ++v) {
            const int gv = (v * 8) / cpg;
            const float4 pv = ps[v];
            if (gv == tid) { a += pv.x; q2 += pv.y; }
            else if (gv + 1 == tid) { a += pv.z; q2 += pv.w; }
        }
        gs[tid] = a;
        gq[tid] = q2;
    }
    __syncthreads();
    const int chunks = gridDim.x;
    float* mypart = part + (static_cast<size_t>(b) * chunks + blockIdx.x) * groups * 2;
    if (tid < groups) {
        mypart[tid * 2 + 0] = gs[tid];
        mypart[tid * 2 + 1] = gq[tid];
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) s_ticket = atomicAdd(&counters[b], 1u);
    __syncthreads();
    if (s_ticket == static_cast<unsigned int>(chunks - 1)) {   // last CTA of this image: fixed-order final reduction
        __threadfence();
        if (tid < groups) {
            // four independent partial chains (loads in flight), combined in a fixed order
            float a[4] = {0.f, 0.f, 0.f, 0.f}, q2[4] = {0.f, 0.f, 0.f, 0.f};
            const float2* pb = reinterpret_cast<const float2*>(part + static_cast<size_t>(b) * chunks * groups * 2);
            int c = 0;
            for (; c + 3 < chunks; c += 4) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float2 v = __ldcg(&pb[static_cast<size_t>(c + k) * groups + tid]);
                    a[k] += v.x; q2[k] += v.y;
                }
            }
            for (; c < chunks; ++c) {
                const float2 v = __ldcg(&pb[static_cast<size_t>(c) * groups + tid]);
                a[0] += v.x; q2[0] += v.y;
            }
            stats[(static_cast<size_t>(b) * groups + tid) * 2 + 0] = (a[0] + a[1]) + (a[2] + a[3]);
            stats[(static_cast<size_t>(b) * groups + tid) * 2 + 1] = (q2[0] + q2[1]) + (q2[2] + q2[3]);
        }
        if (tid == 0) counters[b] = 0u;
    }
}

// Statistics already produced by the igemm epilogue as per-(image, slot, channel) partials (a slot = the 32 rows of one epilogue
// warp): reduce them to stats[b][g].  grid (groups, B), block 128: thread t takes the (slot, channel) pairs t, t + 128, ... of the
// group (channels of the first source, then of the second: a group of a concat may straddle both), then the 128 partials are
// added by a fixed binary tree -> deterministic.
__global__ void __launch_bounds__(128) gn_finalize_kernel(const float2* __restrict__ part1, int C1, int tiles1,
                                                          const float2* __restrict__ part2, int C2, int tiles2, int groups,
                                                          float* __restrict__ stats) {
    __shared__ float2 red[128];
    pdl_trigger();
    pdl_wait();
    const int g = blockIdx.x, b = blockIdx.y, t = threadIdx.x;
    const int C = C1 + C2, cpg = C / groups;
    const int lo = g * cpg, hi = lo + cpg;
    float sm = 0.f, sq = 0.f;
    {   // channels [lo, min(hi, C1)) of source 1
        const int n = min(hi, C1) - lo;
        if (n > 0) {
            const float2* pb = part1 + static_cast<size_t>(b) * tiles1 * C1 + lo;
            for (int i = t; i < n * tiles1; i += 128) {
                const float2 v = __ldcg(pb + static_cast<size_t>(i / n) * C1 + (i % n));
                sm += v.x;
                sq += v.y;
            }
        }
    }
    {   // channels [max(lo, C1), hi) of source 2
        const int l2 = max(lo, C1), n = hi - l2;
        if (n > 0 && C2 > 0) {
            const float2* pb = part2 + static_cast<size_t>(b) * tiles2 * C2 + (l2 - C1);
            for (int i = t; i < n * tiles2; i += 128) {
                const float2 v = __ldcg(pb + static_cast<size_t>(i / n) * C2 + (i % n));
                sm += v.x;
                sq += v.y;
            }
        }
    }
    red[t] = make_float2(sm, sq);
    __syncthreads();
#pragma unroll
    for (int o = 64; o > 0; o >>= 1) {
        if (t < o) {
            red[t].x += red[t + o].x;
            red[t].y += red[t + o].y;
        }
        __syncthreads();
    }
    if (t == 0) {
        stats[(static_cast<size_t>(b) * groups + g) * 2 + 0] = red[0].x;
        stats[(static_cast<size_t>(b) * groups + g) * 2 + 1] = red[0].y;
    }
}

__global__ void gn_apply_kernel(const __nv_bfloat16* __restrict__ x1, int C1, const __nv_bfloat16* __restrict__ x2, int C2,
                                int HW, int groups, int pix_per_cta, const float* __restrict__ stats, float eps,
                                const float* __restrict__ gamma, const float* __restrict__ beta, int silu,
                                __nv_bfloat16* __restrict__ out) {
    const int C = C1 + C2;
    const int cpg = C / groups;
    const int b = blockIdx.y;
    const int c0 = threadIdx.x * 8;
    const float inv_cnt = 1.0f / (static_cast<float>(HW) * cpg);
    pdl_trigger();
    pdl_wait();
    float sc[8], sh[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const int c = c0 + e;
        const int g = c / cpg;
        const float sum = __ldg(&stats[(static_cast<size_t>(b) * groups + g) * 2 + 0]);
        const float sq = __ldg(&stats[(static_cast<size_t>(b) * groups + g) * 2 + 1]);
        const float mean = sum * inv_cnt;
        const float var = fmaxf(sq * inv_cnt - mean * mean, 0.f);
        const float rstd = rsqrtf(var + eps);
        sc[e] = rstd * __ldg(&gamma[c]);
        sh[e] = __ldg(&beta[c]) - mean * sc[e];
    }
    const __nv_bfloat16* src;
    int ld, cc;
    if (c0 < C1) { src = x1; ld = C1; cc = c0; } else { src = x2; ld = C2; cc = c0 - C1; }
    src += static_cast<size_t>(b) * HW * ld + cc;
    __nv_bfloat16* dst = out + static_cast<size_t>(b) * HW * C + c0;
    const int p0 = blockIdx.x * pix_per_cta;
    const int p1 = min(HW, p0 + pix_per_cta);
    const int step = blockDim.y;
    int p = p0 + threadIdx.y;
    for (; p + 7 * step < p1; p += 8 * step) {   // 8 x 16 B in flight per thread
        uint4 u[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) u[k] = __ldg(reinterpret_cast<const uint4*>(src + static_cast<size_t>(p + k * step) * ld));
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            float f[8];
            unpack8(u[k], f);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const float y = fmaf(f[e], sc[e], sh[e]);
                f[e] = silu ? silu_f(y) : y;
            }
            *reinterpret_cast<uint4*>(dst + static_cast<size_t>(p + k * step) * C) = pack8(f);
        }
    }
    for (; p + 3 * step < p1; p += 4 * step) {
        uint4 u[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) u[k] = __ldg(reinterpret_cast<const uint4*>(src + static_cast<size_t>(p + k * step) * ld));
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float f[8];
            unpack8(u[k], f);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const float y = fmaf(f[e], sc[e], sh[e]);
                f[e] = silu ? silu_f(y) : y;
            }
            *reinterpret_cast<uint4*>(dst + static_cast<size_t>(p + k * step) * C) = pack8(f);
        }
    }
    for (; p < p1; p += step) {
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(src + static_cast<size_t>(p) * ld));
        float f[8];
        unpack8(u, f);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const float y = fmaf(f[e], sc[e], sh[e]);
            f[e] = silu ? silu_f(y) : y;
        }
        *reinterpret_cast<uint4*>(dst + static_cast<size_t>(p) * C) = pack8(f);
    }
}

// Small feature maps (8x8 level: a few MB per tensor; measured slower than the two-kernel path from 16x16 up): statistics and normalisation in ONE launch, the
// data held in registers in between.  A CTA owns (image b, slab of G whole groups) with G*cpg a multiple of 8
// channels; blockDim = (vectors per slab, pixel lanes); thread (v, ty) keeps pixels ty, ty+PY, ... of channel vector v
// (<= MAXP of them; up to 1024 threads so that every thread has only a few loads, all in flight at once).  The
// two-kernel path costs two dependent launches on these sizes.  Reduction order is fixed (tree over pixel lanes, then
// vectors in channel order), so results are run-to-run deterministic.
template <int MAXP>
__global__ void __launch_bounds__(1024) gn_small_kernel(const __nv_bfloat16* __restrict__ x1, int C1, const __nv_bfloat16* __restrict__ x2,
                                                        int C2, int HW, int groups, int G, float eps,
                                                        const float* __restrict__ gamma, const float* __restrict__ beta,
                                                        int silu, __nv_bfloat16* __restrict__ out) {
    __shared__ float4 ps[1024];
    __shared__ float g_mean[8], g_rstd[8];
    const int C = C1 + C2;
    const int cpg = C / groups;
    const int b = blockIdx.y;
    const int nvec = blockDim.x, PY = blockDim.y;
    const int tid = threadIdx.y * nvec + threadIdx.x;
    const int cs = blockIdx.x * G * cpg;            // first channel of the slab
    const int c0 = cs + threadIdx.x * 8;
    pdl_trigger();
    pdl_wait();
    const __nv_bfloat16* src;
    int ld, cc;
    if (c0 < C1) { src = x1; ld = C1; cc = c0; } else { src = x2; ld = C2; cc = c0 - C1; }
    src += static_cast<size_t>(b) * HW * ld + cc;
    uint4 u[MAXP];
#pragma unroll
    for (int k = 0; k < MAXP; ++k) {
        const int p = threadIdx.y + k * PY;
        u[k] = p < HW ? __ldg(reinterpret_cast<const uint4*>(src + static_cast<size_t>(p) * ld)) : make_uint4(0, 0, 0, 0);
    }
    float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, q[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < MAXP; ++k) {
        float f[8];
        unpack8(u[k], f);
#pragma unroll
        for (int e = 0; e < 8; ++e) { s[e] += f[e]; q[e] = fmaf(f[e], f[e], q[e]); }
    }
    // 8 consecutive channels touch at most two groups (cpg >= 8)
    const int g0 = (c0 - cs) / cpg;                 // slab-local group of the first channel
    float sa = 0.f, qa = 0.f, sb = 0.f, qb = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        if ((c0 - cs + e) / cpg == g0) { sa += s[e]; qa += q[e]; } else { sb += s[e]; qb += q[e]; }
    }
    ps[tid] = make_float4(sa, qa, sb, qb);
    __syncthreads();
    // level 1: binary tree over the pixel lanes of each channel vector (fixed shape -> deterministic)
    int stride = 1;
    while (stride < PY) stride <<= 1;
    for (stride >>= 1; stride > 0; stride >>= 1) {
        if (int(threadIdx.y) < stride && int(threadIdx.y) + stride < PY) {
            float4 a = ps[tid];
            const float4 v = ps[tid + stride * nvec];
            a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
            ps[tid] = a;
        }
        __syncthreads();
    }
    if (tid < G) {                                  // level 2: over the vectors that touch group tid, in channel order
        float a = 0.f, q2 = 0.f;
        for (int v = 0; v < nvec; ++v) {
            const int gv = (v * 8) / cpg;
            const float4 pv = ps[v];
            if (gv == tid) { a += pv.x; q2 += pv.y; }
            else if (gv + 1 == tid) { a += pv.z; q2 += pv.w; }
        }
        const float inv_cnt = 1.0f / (static_cast<float>(HW) * cpg);
        const float mean = a * inv_cnt;
        const float var = fmaxf(q2 * inv_cnt - mean * mean, 0.f);
        g_mean[tid] = mean;
        g_rstd[tid] = rsqrtf(var + eps);
    }
    __syncthreads();
    float sc[8], sh[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const int c = c0 + e;
        const int g = (c - cs) / cpg;
        sc[e] = g_rstd[g] * __ldg(&gamma[c]);
        sh[e] = __ldg(&beta[c]) - g_mean[g] * sc[e];
    }
    __nv_bfloat16* dst = out + static_cast<size_t>(b) * HW * C + c0;
#pragma unroll
    for (int k = 0; k < MAXP; ++k) {
        const int p = threadIdx.y + k * PY;
        if (p < HW) {
            float f[8];
            unpack8(u[k], f);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const float y = fmaf(f[e], sc[e], sh[e]);
                f[e] = silu ? silu_f(y) : y;
            }
            *reinterpret_cast<uint4*>(dst + static_cast<size_t>(p) * C) = pack8(f);
        }
    }
}

// one warp per row, two-pass (mean, then centred variance) on register-resident data
template <int NV>
__global__ void layernorm_kernel(const __nv_bfloat16* __restrict__ x, int rows, int C, float eps,
                                 const float* __restrict__ gamma, const float* __restrict__ beta,
                                 __nv_bfloat16* __restrict__ out) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    pdl_trigger();
    pdl_wait();
    if (row >= rows) return;
    const int nvec = C >> 3;
    const __nv_bfloat16* src = x + static_cast<size_t>(row) * C;
    float f[NV][8];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int v = lane + i * 32;
        if (v < nvec) {
            unpack8(__ldg(reinterpret_cast<const uint4*>(src + v * 8)), f[i]);
#pragma unroll
            for (int e = 0; e < 8; ++e) sum += f[i][e];
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum / C;
    float var = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int v = lane + i * 32;
        if (v < nvec) {
#pragma unroll
            for (int e = 0; e < 8; ++e) { const float d = f[i][e] - mean; var = fmaf(d, d, var); }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) var += __shfl_xor_sync(0xffffffffu, var, o);
    const float rstd = rsqrtf(var / C + eps);
    __nv_bfloat16* dst = out + static_cast<size_t>(row) * C;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int v = lane + i * 32;
        if (v < nvec) {
            const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + v * 8));
            const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + v * 8 + 4));
            const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + v * 8));
            const float4 b1 = __ldg(reinterpret_cast<const float4*>(beta + v * 8 + 4));
            float y[8];
            y[0] = (f[i][0] - mean) * rstd * g0.x + b0.x;
            y[1] = (f[i][1] - mean) * rstd * g0.y + b0.y;
            y[2] = (f[i][2] - mean) * rstd * g0.z + b0.z;
            y[3] = (f[i][3] - mean) * rstd * g0.w + b0.w;
            y[4] = (f[i][4] - mean) * rstd * g1.x + b1.x;
            y[5] = (f[i][5] - mean) * rstd * g1.y + b1.y;
            y[6] = (f[i][6] - mean) * rstd * g1.z + b1.z;
            y[7] = (f[i][7] - mean) * rstd * g1.w + b1.w;
            *reinterpret_cast<uint4*>(dst + v * 8) = pack8(y);
        }
    }
}

// Rows of exactly 5 * L sixteen-byte vectors (C = 320 / 640 for L = 8 / 16 — the LayerNorms of the two large SD1.5
// transformer levels): L lanes share a row, each lane owns 5 vectors, so a warp normalises 32 / L rows at once with every
// lane busy and all 5 loads of a lane in flight together (C = 320: 2.5 KB in flight per warp instead of 640 B with 62 % of
// the lanes loading, and a quarter of the CTAs).  Same two-pass arithmetic as layernorm_kernel.
template <int L>
__global__ void __launch_bounds__(256) layernorm5_kernel(const __nv_bfloat16* __restrict__ x, int rows, float eps,
                                                         const float* __restrict__ gamma, const float* __restrict__ beta,
                                                         __nv_bfloat16* __restrict__ out) {
    constexpr int RPW = 32 / L, C = L * 40;
    const int lane = threadIdx.x & 31, sub = lane % L;
    const int row = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * RPW + lane / L;
    pdl_trigger();
    pdl_wait();
    const bool ok = row < rows;
    const __nv_bfloat16* src = x + static_cast<size_t>(ok ? row : rows - 1) * C;
    uint4 u[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) u[i] = __ldg(reinterpret_cast<const uint4*>(src + (sub + L * i) * 8));
    float f[5][8];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        unpack8(u[i], f[i]);
#pragma unroll
        for (int e = 0; e < 8; ++e) sum += f[i][e];
    }
#pragma unroll
    for (int o = L / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum / C;
    float var = 0.f;
#pragma unroll
    for (int i = 0; i < 5; ++i)
#pragma unroll
        for (int e = 0; e < 8; ++e) { const float d = f[i][e] - mean; var = fmaf(d, d, var); }
#pragma unroll
    for (int o = L / 2; o > 0; o >>= 1) var += __shfl_xor_sync(0xffffffffu, var, o);
    const float rstd = rsqrtf(var / C + eps);
    if (!ok) return;
    __nv_bfloat16* dst = out + static_cast<size_t>(row) * C;
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        const int v = sub + L * i;
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + v * 8));
        const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + v * 8 + 4));
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + v * 8));
        const float4 b1 = __ldg(reinterpret_cast<const float4*>(beta + v * 8 + 4));
        float y[8];
        y[0] = (f[i][0] - mean) * rstd * g0.x + b0.x;
        y[1] = (f[i][1] - mean) * rstd * g0.y + b0.y;
        y[2] = (f[i][2] - mean) * rstd * g0.z + b0.z;
        y[3] = (f[i][3] - mean) * rstd * g0.w + b0.w;
        y[4] = (f[i][4] - mean) * rstd * g1.x + b1.x;
        y[5] = (f[i][5] - mean) * rstd * g1.y + b1.y;
        y[6] = (f[i][6] - mean) * rstd * g1.z + b1.z;
        y[7] = (f[i][7] - mean) * rstd * g1.w + b1.w;
        *reinterpret_cast<uint4*>(dst + v * 8) = pack8(y);
    }
}

}  // namespace mfb

using namespace mfb;

static int groupnorm_impl(const void* x1, int C1, const void* x2, int C2, int B, int HW, int groups, float eps, const float* gamma,
                          const float* beta, int silu, float* stats_ws, void* out, void* stream, const float* part1, int tiles1,
                          const float* part2, int tiles2) {
    MFB_REQUIRE(x1 && out && gamma && beta && stats_ws, "null pointer");
    if (!x2) C2 = 0;
    const int C = C1 + C2;
    MFB_REQUIRE(C1 % 8 == 0 && C2 % 8 == 0, "channel counts must be multiples of 8");
    // an 8-channel vector may straddle at most two groups: >= 8 channels per group, or exactly 4 (the VAE's 128-channel level)
    MFB_REQUIRE(groups > 0 && groups <= 64 && C % groups == 0 && (C / groups >= 8 || C / groups == 4),
                "unsupported group size (C=%d groups=%d)", C, groups);
    const int CV = C / 8;
    MFB_REQUIRE(CV <= 1024, "C too large");
    const int PY = CV >= 256 ? 1 : 256 / CV;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    static const int small_max_hw = [] { const char* e = getenv("MFB_GN_SMALL_HW"); return e ? atoi(e) : 64; }();
    if (part1 == nullptr && HW <= small_max_hw) {
        // single-launch path for the small feature maps
        const int cpg = C / groups;
        int G = 1;
        while (G <= 4 && (G * cpg) % 8 != 0) G *= 2;
        const int nvec = G * cpg / 8;
        if (G <= 4 && groups % G == 0 && nvec <= 64) {
            const int py = 256 / nvec < HW ? 256 / nvec : HW;   // ~256 threads: several CTAs per SM, one wave
            const int ppt = (HW + py - 1) / py;
            dim3 g2(groups / G, B), b2(nvec, py);
            auto X1 = static_cast<const __nv_bfloat16*>(x1);
            auto X2 = static_cast<const __nv_bfloat16*>(x2);
            auto O = static_cast<__nv_bfloat16*>(out);
            if (ppt <= 2) {
                MFB_CUDA_OK(launch_k(gn_small_kernel<2>, g2, b2, 0, st, 1, X1, C1, X2, C2, HW, groups, G, eps, gamma, beta, silu, O));
                return MFB_OK;
            }
            if (ppt <= 4) {
                MFB_CUDA_OK(launch_k(gn_small_kernel<4>, g2, b2, 0, st, 1, X1, C1, X2, C2, HW, groups, G, eps, gamma, beta, silu, O));
                return MFB_OK;
            }
        }
    }
    // Grid = ONE full wave of resident CTAs (occupancy API x SM count), every CTA an equal pixel chunk: a grid a
    // third larger than the resident set (the earlier fixed 4 CTAs/SM guess against 3 resident at 80 registers)
    // costs a whole second wave.  At least 4*PY pixels per CTA.
    auto pick_chunks = [&](const void* kernel) {
        int resident = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, kernel, CV * PY, 0) != cudaSuccess || resident < 1) resident = 2;
        int chunks = (resident * device_sm_count()) / B;
        const int max_chunks = (HW + 4 * PY - 1) / (4 * PY);
        if (chunks > max_chunks) chunks = max_chunks;
        if (chunks > MFB_GN_MAX_CHUNKS) chunks = MFB_GN_MAX_CHUNKS;
        if (chunks < 1) chunks = 1;
        const int ppc = (HW + chunks - 1) / chunks;
        return ppc;
    };
    const int ppc = pick_chunks(reinterpret_cast<const void*>(gn_stats_kernel));          // statistics pass
    const int chunks = (HW + ppc - 1) / ppc;
    const int ppc_apply = pick_chunks(reinterpret_cast<const void*>(gn_apply_kernel));    // apply pass
    const int chunks_apply = (HW + ppc_apply - 1) / ppc_apply;
    // workspace layout: stats [B*groups*2] | partials [B*MAX_CHUNKS*groups*2] | ticket counters [B] (zero on first use)
    float* part = stats_ws + static_cast<size_t>(2) * B * groups;
    unsigned int* counters = reinterpret_cast<unsigned int*>(part + static_cast<size_t>(2) * B * groups * MFB_GN_MAX_CHUNKS);
    dim3 grid(chunks, B), block(CV, PY);
    if (part1 != nullptr && (C2 == 0 || part2 != nullptr)) {
        // statistics came with the producing GEMM(s): only the tiny fixed-order reduction is left
        MFB_CUDA_OK(launch_k(gn_finalize_kernel, dim3(groups, B), dim3(128), 0, st, 1, reinterpret_cast<const float2*>(part1), C1, tiles1,
                             reinterpret_cast<const float2*>(part2), C2, tiles2, groups, stats_ws));
    } else {
        MFB_CUDA_OK(launch_k(gn_stats_kernel, grid, block, 0, st, 1, static_cast<const __nv_bfloat16*>(x1), C1,
                             static_cast<const __nv_bfloat16*>(x2), C2, HW, groups, ppc, stats_ws, part, counters));
    }
    MFB_CUDA_OK(launch_k(gn_apply_kernel, dim3(chunks_apply, B), block, 0, st, 1, static_cast<const __nv_bfloat16*>(x1), C1,
                         static_cast<const __nv_bfloat16*>(x2), C2, HW, groups, ppc_apply, static_cast<const float*>(stats_ws), eps, gamma,
                         beta, silu, static_cast<__nv_bfloat16*>(out)));
    return MFB_OK;
}

// Statistics only: stats_ws[0 : 2*B*groups] = per-(image, group) {sum, sum of squares} (the layout gn_apply_kernel and the
// GroupNorm backward read); same workspace contract as mfb_groupnorm.  Used by the training path, whose backward needs them.
extern "C" int mfb_groupnorm_stats(const void* x1, int C1, const void* x2, int C2, int B, int HW, int groups, float* stats_ws,
                                   void* stream) {
    MFB_REQUIRE(x1 && stats_ws, "null pointer");
    if (!x2) C2 = 0;
    const int C = C1 + C2;
    MFB_REQUIRE(C1 % 8 == 0 && C2 % 8 == 0 && groups > 0 && groups <= 64 && C % groups == 0 && (C / groups >= 8 || C / groups == 4),
                "unsupported group size (C=%d groups=%d)", C, groups);
    const int CV = C / 8;
    MFB_REQUIRE(CV <= 1024, "C too large");
    const int PY = CV >= 256 ? 1 : 256 / CV;
    int resident = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, gn_stats_kernel, CV * PY, 0) != cudaSuccess || resident < 1) resident = 2;
    int chunks = (resident * device_sm_count()) / B;
    const int max_chunks = (HW + 4 * PY - 1) / (4 * PY);
    if (chunks > max_chunks) chunks = max_chunks;
    if (chunks > MFB_GN_MAX_CHUNKS) chunks = MFB_GN_MAX_CHUNKS;
    if (chunks < 1) chunks = 1;
    const int ppc = (HW + chunks - 1) / chunks;
    chunks = (HW + ppc - 1) / ppc;
    float* part = stats_ws + static_cast<size_t>(2) * B * groups;
    unsigned int* counters = reinterpret_cast<unsigned int*>(part + static_cast<size_t>(2) * B * groups * MFB_GN_MAX_CHUNKS);
    MFB_CUDA_OK(launch_k(gn_stats_kernel, dim3(chunks, B), dim3(CV, PY), 0, static_cast<cudaStream_t>(stream), 1,
                         static_cast<const __nv_bfloat16*>(x1), C1, static_cast<const __nv_bfloat16*>(x2), C2, HW, groups, ppc, stats_ws,
                         part, counters));
    return MFB_OK;
}

extern "C" int mfb_groupnorm(const void* x1, int C1, const void* x2, int C2, int B, int HW, int groups, float eps,
                             const float* gamma, const float* beta, int silu, float* stats_ws, void* out, void* stream) {
    MFB_RECORD(mfb_groupnorm(x1, C1, x2, C2, B, HW, groups, eps, gamma, beta, silu, stats_ws, out, stream));
    return groupnorm_impl(x1, C1, x2, C2, B, HW, groups, eps, gamma, beta, silu, stats_ws, out, stream, nullptr, 0, nullptr, 0);
}

extern "C" int mfb_groupnorm_prestat(const void* x1, int C1, const float* part1, int tiles1, const void* x2, int C2,
                                     const float* part2, int tiles2, int B, int HW, int groups, float eps, const float* gamma,
                                     const float* beta, int silu, float* stats_ws, void* out, void* stream) {
    MFB_RECORD(mfb_groupnorm_prestat(x1, C1, part1, tiles1, x2, C2, part2, tiles2, B, HW, groups, eps, gamma, beta, silu, stats_ws, out, stream));
    MFB_REQUIRE(part1 && tiles1 > 0 && (!x2 || (part2 && tiles2 > 0)), "mfb_groupnorm_prestat needs the partial statistics of every source");
    MFB_REQUIRE(groups <= 64, "at most 64 groups");
    return groupnorm_impl(x1, C1, x2, C2, B, HW, groups, eps, gamma, beta, silu, stats_ws, out, stream, part1, tiles1, part2, tiles2);
}

// Row softmax over [rows, cols] bf16 (fp32 math), one CTA of 256 threads per row, 16-byte accesses; in place allowed.
// Used by the single-head (d = 512) attention of the VAE mid block, whose scores come out of a GEMM already scaled
// (S/models/attention_processor.py:1266-1268 — softmax(Q K^T * d^-0.5) as two GEMMs around this kernel).
namespace mfb {
__global__ void __launch_bounds__(256) softmax_rows_kernel(const __nv_bfloat16* __restrict__ x, int cols, __nv_bfloat16* __restrict__ out) {
    __shared__ float red[8];
    pdl_trigger();
    pdl_wait();
    const __nv_bfloat16* xr = x + static_cast<size_t>(blockIdx.x) * cols;
    __nv_bfloat16* orow = out + static_cast<size_t>(blockIdx.x) * cols;
    const int nv = cols / 8;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    auto block_reduce = [&](float v, bool is_max) {
        for (int o = 16; o > 0; o >>= 1) {
            const float t = __shfl_xor_sync(0xffffffffu, v, o);
            v = is_max ? fmaxf(v, t) : v + t;
        }
        __syncthreads();
        if (lane == 0) red[wid] = v;
        __syncthreads();
        float r = red[0];
        for (int i = 1; i < 8; ++i) r = is_max ? fmaxf(r, red[i]) : r + red[i];      // fixed order: deterministic
        return r;
    };
    float mx = -INFINITY;
    for (int v = threadIdx.x; v < nv; v += 256) {
        float f[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(xr) + v), f);
#pragma unroll
        for (int e = 0; e < 8; ++e) mx = fmaxf(mx, f[e]);
    }
    mx = block_reduce(mx, true);
    float sum = 0.f;
    for (int v = threadIdx.x; v < nv; v += 256) {
        float f[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(xr) + v), f);
#pragma unroll
        for (int e = 0; e < 8; ++e) sum += __expf(f[e] - mx);
    }
    sum = block_reduce(sum, false);
    const float inv = 1.0f / sum;
    for (int v = threadIdx.x; v < nv; v += 256) {
        float f[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(xr) + v), f);
#pragma unroll
        for (int e = 0; e < 8; ++e) f[e] = __expf(f[e] - mx) * inv;
        reinterpret_cast<uint4*>(orow)[v] = pack8(f);
    }
}
}  // namespace mfb

extern "C" int mfb_softmax_rows(const void* x, int rows, int cols, void* out, void* stream) {
    MFB_RECORD(mfb_softmax_rows(x, rows, cols, out, stream));
    MFB_REQUIRE(x && out && rows > 0, "null pointer");
    MFB_REQUIRE(cols > 0 && cols % 8 == 0, "cols must be a positive multiple of 8 (got %d)", cols);
    MFB_CUDA_OK(launch_k(softmax_rows_kernel, dim3(rows), dim3(256), 0, static_cast<cudaStream_t>(stream), 1,
                         static_cast<const __nv_bfloat16*>(x), cols, static_cast<__nv_bfloat16*>(out)));
    return MFB_OK;
}

extern "C" int mfb_layernorm(const void* x, int rows, int C, float eps, const float* gamma, const float* beta, void* out,
                             void* stream) {
    MFB_RECORD(mfb_layernorm(x, rows, C, eps, gamma, beta, out, stream));
    MFB_REQUIRE(x && out && gamma && beta, "null pointer");
    MFB_REQUIRE(C % 8 == 0 && C <= 2048, "C must be a multiple of 8 and <= 2048 (got %d)", C);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int warps = 8;
    dim3 grid((rows + warps - 1) / warps), block(warps * 32);
    const int nv = (C / 8 + 31) / 32;
    auto X = static_cast<const __nv_bfloat16*>(x);
    auto O = static_cast<__nv_bfloat16*>(out);
    static const bool use5 = [] { const char* e = getenv("MFB_LN5"); return !e || atoi(e) != 0; }();
    // several rows per warp for the narrow rows (measured, tools/bench_ln.py, B200: C = 320 20.9 -> 17.6 us per 42 MB tensor,
    // C = 640 11.3 -> 10.5 us; at C = 1280 a warp already owns one full row and the old kernel is faster: 6.4 vs 7.3 us)
    if (use5 && (C == 320 || C == 640)) {
        const int L = C / 40, rpw = 32 / L;
        dim3 g5((rows + warps * rpw - 1) / (warps * rpw));
        if (L == 8) MFB_CUDA_OK(launch_k(layernorm5_kernel<8>, g5, block, 0, st, 1, X, rows, eps, gamma, beta, O));
        else MFB_CUDA_OK(launch_k(layernorm5_kernel<16>, g5, block, 0, st, 1, X, rows, eps, gamma, beta, O));
        return MFB_OK;
    }
    if (nv <= 2) MFB_CUDA_OK(launch_k(layernorm_kernel<2>, grid, block, 0, st, 1, X, rows, C, eps, gamma, beta, O));
    else if (nv <= 5) MFB_CUDA_OK(launch_k(layernorm_kernel<5>, grid, block, 0, st, 1, X, rows, C, eps, gamma, beta, O));
    else MFB_CUDA_OK(launch_k(layernorm_kernel<8>, grid, block, 0, st, 1, X, rows, C, eps, gamma, beta, O));
    return MFB_OK;
}
