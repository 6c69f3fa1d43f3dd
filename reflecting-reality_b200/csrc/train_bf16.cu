// bf16 backward kernels of the bandwidth-bound ops of the fine-tune step (BASELINE config 4): what autograd runs under
// `accelerator.backward(loss)` (E/train_brushnet_mirror.py:1459) for F.group_norm + F.silu (S/models/resnet.py:337-338,381,393;
// transformer_2d.py:338), F.layer_norm (S/models/attention.py:313,360,386), GEGLU (S/models/activations.py:100-103) and conv_out
// (S/models/unets/unet_2d_condition.py:1339).  All HBM-bound: 16-byte vectors, a thread keeps a fixed 8-channel slice, fp32 math,
// deterministic fixed-order reductions (per-CTA partials + ticketed last-CTA finalize; no float atomics).
#include <stdlib.h>

#include "common.h"
#include "ptx.cuh"

namespace mfb {

__device__ __forceinline__ void unpack8b(const uint4& u, float (&f)[8]) {
    float2 t;
    t = unpack_bf16x2(u.x); f[0] = t.x; f[1] = t.y;
    t = unpack_bf16x2(u.y); f[2] = t.x; f[3] = t.y;
    t = unpack_bf16x2(u.z); f[4] = t.x; f[5] = t.y;
    t = unpack_bf16x2(u.w); f[6] = t.x; f[7] = t.y;
}
__device__ __forceinline__ uint4 pack8b(const float (&f)[8]) {
    uint4 o;
    o.x = pack_bf16x2(f[0], f[1]);
    o.y = pack_bf16x2(f[2], f[3]);
    o.z = pack_bf16x2(f[4], f[5]);
    o.w = pack_bf16x2(f[6], f[7]);
    return o;
}

// derivative of silu at z times the incoming gradient
__device__ __forceinline__ float silu_grad(float z, float dy) {
    const float sg = __frcp_rn(1.0f + __expf(-z));
    return dy * sg * fmaf(z, 1.0f - sg, 1.0f);
}

constexpr int GNB_MAX_CHUNKS = 32;

// ------------------------------------------------------------------------------------------------ GroupNorm(+SiLU) backward
// y = silu?(gamma * xh + beta), xh = (x - mean_g) * rstd_g over the (HW x cpg) slab of group g of image b.  With dz = dL/d(gamma xh + beta):
//   dbeta_c = sum_p dz,  dgamma_c = sum_p dz xh,  S1_g = mean(dz gamma),  S2_g = mean(dz gamma xh),
//   dx = rstd (dz gamma - S1 - xh S2)  (+ the gradients arriving over the residual paths).
// Pass 1 (this kernel): per-channel sums a_c = sum_p dz, b_c = sum_p dz xh for one pixel chunk; the last CTA of an image adds the
// chunk partials in chunk order, writes them to ws_db / ws_dg [B][C] and derives S1 / S2 per group.  grid (chunks, B), block (CV, PY).
__global__ void gn_bwd_sums_kernel(const __nv_bfloat16* __restrict__ x1, int C1, const __nv_bfloat16* __restrict__ x2, int C2,
                                   const __nv_bfloat16* __restrict__ dy, int HW, int groups, int pix_per_cta,
                                   const float* __restrict__ stats, float eps, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, int silu, float* __restrict__ part, float* __restrict__ ws_dg,
                                   float* __restrict__ ws_db, float* __restrict__ gsum, unsigned int* __restrict__ counters) {
    extern __shared__ float sm[];           // [2][PY][C]
    __shared__ unsigned int s_ticket;
    const int C = C1 + C2, cpg = C / groups;
    const int b = blockIdx.y, c0 = threadIdx.x * 8;
    const int PY = blockDim.y, nthr = blockDim.x * blockDim.y, tid = threadIdx.y * blockDim.x + threadIdx.x;
    const float inv_cnt = 1.0f / (static_cast<float>(HW) * cpg);
    pdl_trigger();
    pdl_wait();
    float mean[8], rstd[8], gm[8], bt[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const int c = c0 + e, g = c / cpg;
        const float sum = __ldg(&stats[(static_cast<size_t>(b) * groups + g) * 2 + 0]);
        const float sq = __ldg(&stats[(static_cast<size_t>(b) * groups + g) * 2 + 1]);
        mean[e] = sum * inv_cnt;
        rstd[e] = rsqrtf(fmaxf(sq * inv_cnt - mean[e] * mean[e], 0.f) + eps);
        gm[e] = __ldg(&gamma[c]);
        bt[e] = __ldg(&beta[c]);
    }
    const __nv_bfloat16* src;
    int ld, cc;
    if (c0 < C1) { src = x1; ld = C1; cc = c0; } else { src = x2; ld = C2; cc = c0 - C1; }
    src += static_cast<size_t>(b) * HW * ld + cc;
    const __nv_bfloat16* dsrc = dy + static_cast<size_t>(b) * HW * C + c0;
    const int p0 = blockIdx.x * pix_per_cta, p1 = min(HW, p0 + pix_per_cta);
    float a[8], bb[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) { a[e] = 0.f; bb[e] = 0.f; }
    auto accum = [&](const uint4& ux, const uint4& ud) {
        float fx[8], fd[8];
        unpack8b(ux, fx);
        unpack8b(ud, fd);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const float xh = (fx[e] - mean[e]) * rstd[e];
            const float dz = silu ? silu_grad(fmaf(gm[e], xh, bt[e]), fd[e]) : fd[e];
            a[e] += dz;
            bb[e] = fmaf(dz, xh, bb[e]);
        }
    };
    int p = p0 + threadIdx.y;
    for (; p + 3 * PY < p1; p += 4 * PY) {            // 8 x 16 B in flight per thread
        uint4 ux[4], ud[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            ux[k] = __ldg(reinterpret_cast<const uint4*>(src + static_cast<size_t>(p + k * PY) * ld));
            ud[k] = __ldg(reinterpret_cast<const uint4*>(dsrc + static_cast<size_t>(p + k * PY) * C));
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) accum(ux[k], ud[k]);
    }
    for (; p < p1; p += PY)
        accum(__ldg(reinterpret_cast<const uint4*>(src + static_cast<size_t>(p) * ld)),
              __ldg(reinterpret_cast<const uint4*>(dsrc + static_cast<size_t>(p) * C)));
    // CTA reduction over the pixel lanes, in lane order
    float* sa = sm;
    float* sb = sm + static_cast<size_t>(PY) * C;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        sa[threadIdx.y * C + c0 + e] = a[e];
        sb[threadIdx.y * C + c0 + e] = bb[e];
    }
    __syncthreads();
    const int chunks = gridDim.x;
    float* mypart = part + (static_cast<size_t>(b) * chunks + blockIdx.x) * C * 2;
    for (int c = tid; c < C; c += nthr) {
        float va = 0.f, vb = 0.f;
        for (int y = 0; y < PY; ++y) { va += sa[y * C + c]; vb += sb[y * C + c]; }
        mypart[c] = va;
        mypart[C + c] = vb;
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) s_ticket = atomicAdd(&counters[b], 1u);
    __syncthreads();
    if (s_ticket != static_cast<unsigned int>(chunks - 1)) return;
    // last CTA of this image: chunk partials in chunk order -> per-channel sums, then the two group means
    __threadfence();
    const float* pb = part + static_cast<size_t>(b) * chunks * C * 2;
    for (int c = tid; c < C; c += nthr) {
        float va = 0.f, vb = 0.f;
        for (int k = 0; k < chunks; ++k) {
            va += __ldcg(pb + static_cast<size_t>(k) * C * 2 + c);
            vb += __ldcg(pb + static_cast<size_t>(k) * C * 2 + C + c);
        }
        ws_db[static_cast<size_t>(b) * C + c] = va;
        ws_dg[static_cast<size_t>(b) * C + c] = vb;
        const float g_ = __ldg(&gamma[c]);
        sa[c] = va * g_;
        sb[c] = vb * g_;
    }
    __syncthreads();
    for (int g = tid; g < groups; g += nthr) {
        float s1 = 0.f, s2 = 0.f;
        for (int c = g * cpg; c < (g + 1) * cpg; ++c) { s1 += sa[c]; s2 += sb[c]; }
        gsum[(static_cast<size_t>(b) * groups + g) * 2 + 0] = s1 * inv_cnt;
        gsum[(static_cast<size_t>(b) * groups + g) * 2 + 1] = s2 * inv_cnt;
    }
    if (tid == 0) counters[b] = 0u;
}

// Pass 2: dx = rstd (dz gamma - S1 - xh S2) + dres + dres2, written per source tensor of the concat.  grid (chunks, B), block (CV, PY).
__global__ void gn_bwd_apply_kernel(const __nv_bfloat16* __restrict__ x1, int C1, const __nv_bfloat16* __restrict__ x2, int C2,
                                    const __nv_bfloat16* __restrict__ dy, int HW, int groups, int pix_per_cta,
                                    const float* __restrict__ stats, const float* __restrict__ gsum, float eps,
                                    const float* __restrict__ gamma, const float* __restrict__ beta, int silu,
                                    const __nv_bfloat16* __restrict__ dres, const __nv_bfloat16* __restrict__ dres2,
                                    __nv_bfloat16* __restrict__ dx1, __nv_bfloat16* __restrict__ dx2) {
    const int C = C1 + C2, cpg = C / groups;
    const int b = blockIdx.y, c0 = threadIdx.x * 8;
    const int PY = blockDim.y;
    const float inv_cnt = 1.0f / (static_cast<float>(HW) * cpg);
    pdl_trigger();
    pdl_wait();
    float mean[8], rstd[8], gm[8], bt[8], s1[8], s2[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const int c = c0 + e, g = c / cpg;
        const size_t gi = (static_cast<size_t>(b) * groups + g) * 2;
        const float sum = __ldg(&stats[gi]), sq = __ldg(&stats[gi + 1]);
        mean[e] = sum * inv_cnt;
        rstd[e] = rsqrtf(fmaxf(sq * inv_cnt - mean[e] * mean[e], 0.f) + eps);
        gm[e] = __ldg(&gamma[c]);
        bt[e] = __ldg(&beta[c]);
        s1[e] = __ldg(&gsum[gi]);
        s2[e] = __ldg(&gsum[gi + 1]);
    }
    const __nv_bfloat16* src;
    __nv_bfloat16* dst;
    int ld, cc;
    if (c0 < C1) { src = x1; dst = dx1; ld = C1; cc = c0; } else { src = x2; dst = dx2; ld = C2; cc = c0 - C1; }
    src += static_cast<size_t>(b) * HW * ld + cc;
    dst += static_cast<size_t>(b) * HW * ld + cc;
    const size_t full = static_cast<size_t>(b) * HW * C + c0;
    const int p0 = blockIdx.x * pix_per_cta, p1 = min(HW, p0 + pix_per_cta);
    auto one = [&](int p, const uint4& ux, const uint4& ud) {
        float fx[8], fd[8], r[8];
        unpack8b(ux, fx);
        unpack8b(ud, fd);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const float xh = (fx[e] - mean[e]) * rstd[e];
            const float dz = silu ? silu_grad(fmaf(gm[e], xh, bt[e]), fd[e]) : fd[e];
            r[e] = rstd[e] * (fmaf(dz, gm[e], -s1[e]) - xh * s2[e]);
        }
        if (dres) {
            float fr[8];
            unpack8b(__ldg(reinterpret_cast<const uint4*>(dres + full + static_cast<size_t>(p) * C)), fr);
#pragma unroll
            for (int e = 0; e < 8; ++e) r[e] += fr[e];
        }
        if (dres2) {
            float fr[8];
            unpack8b(__ldg(reinterpret_cast<const uint4*>(dres2 + full + static_cast<size_t>(p) * C)), fr);
#pragma unroll
            for (int e = 0; e < 8; ++e) r[e] += fr[e];
        }
        *reinterpret_cast<uint4*>(dst + static_cast<size_t>(p) * ld) = pack8b(r);
    };
    int p = p0 + threadIdx.y;
    for (; p + 3 * PY < p1; p += 4 * PY) {
        uint4 ux[4], ud[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            ux[k] = __ldg(reinterpret_cast<const uint4*>(src + static_cast<size_t>(p + k * PY) * ld));
            ud[k] = __ldg(reinterpret_cast<const uint4*>(dy + full + static_cast<size_t>(p + k * PY) * C));
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) one(p + k * PY, ux[k], ud[k]);
    }
    for (; p < p1; p += PY)
        one(p, __ldg(reinterpret_cast<const uint4*>(src + static_cast<size_t>(p) * ld)),
            __ldg(reinterpret_cast<const uint4*>(dy + full + static_cast<size_t>(p) * C)));
}

// dgamma[c] (+)= sum_b ws_dg[b][c], dbeta likewise, in image order
__global__ void gn_bwd2_final_kernel(const float* __restrict__ ws_dg, const float* __restrict__ ws_db, int B, int C,
                                     float* __restrict__ dgamma, float* __restrict__ dbeta, int accumulate) {
    pdl_trigger();
    pdl_wait();
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float a = 0.f, bsum = 0.f;
    for (int b = 0; b < B; ++b) {
        a += ws_dg[static_cast<size_t>(b) * C + c];
        bsum += ws_db[static_cast<size_t>(b) * C + c];
    }
    if (dgamma) dgamma[c] = a + (accumulate ? dgamma[c] : 0.f);
    if (dbeta) dbeta[c] = bsum + (accumulate ? dbeta[c] : 0.f);
}

// ------------------------------------------------------------------------------------------------ LayerNorm backward (data gradient)
// One warp per row, the row in registers: dx = rstd (g - mean(g) - xh mean(g xh)) + dres, g = dy * gamma (statistics recomputed
// from x: the row is read anyway).  NV = 16-byte vectors per lane.
template <int NV>
__global__ void __launch_bounds__(256) ln_bwd_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dy, int rows,
                                                     int C, float eps, const float* __restrict__ gamma,
                                                     const __nv_bfloat16* __restrict__ dres, __nv_bfloat16* __restrict__ dx) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    pdl_trigger();
    pdl_wait();
    if (row >= rows) return;
    const int nvec = C >> 3;
    const size_t off = static_cast<size_t>(row) * C;
    float fx[NV][8], g[NV][8];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int v = lane + i * 32;
        if (v < nvec) {
            unpack8b(__ldg(reinterpret_cast<const uint4*>(x + off + v * 8)), fx[i]);
            unpack8b(__ldg(reinterpret_cast<const uint4*>(dy + off + v * 8)), g[i]);
#pragma unroll
            for (int e = 0; e < 8; ++e) sum += fx[i][e];
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum / C;
    float var = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i)
        if (lane + i * 32 < nvec)
#pragma unroll
            for (int e = 0; e < 8; ++e) { const float d = fx[i][e] - mean; var = fmaf(d, d, var); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) var += __shfl_xor_sync(0xffffffffu, var, o);
    const float rstd = rsqrtf(var / C + eps);
    float sg = 0.f, sgx = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int v = lane + i * 32;
        if (v < nvec) {
            const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + v * 8));
            const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + v * 8 + 4));
            const float gv[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                fx[i][e] = (fx[i][e] - mean) * rstd;      // xh
                g[i][e] *= gv[e];
                sg += g[i][e];
                sgx = fmaf(g[i][e], fx[i][e], sgx);
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sg += __shfl_xor_sync(0xffffffffu, sg, o);
        sgx += __shfl_xor_sync(0xffffffffu, sgx, o);
    }
    const float mg = sg / C, mgx = sgx / C;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int v = lane + i * 32;
        if (v < nvec) {
            float r[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) r[e] = rstd * (g[i][e] - mg - fx[i][e] * mgx);
            if (dres) {
                float fr[8];
                unpack8b(__ldg(reinterpret_cast<const uint4*>(dres + off + v * 8)), fr);
#pragma unroll
                for (int e = 0; e < 8; ++e) r[e] += fr[e];
            }
            *reinterpret_cast<uint4*>(dx + off + v * 8) = pack8b(r);
        }
    }
}

// ------------------------------------------------------------------------------------------------ GEGLU on an un-fused projection
// proj [rows, 2C] = [h | gate] (activations.py:100-103): out = h * gelu_erf(gate) (forward, optional);
// d proj = [d out * gelu(gate) | d out * h * (Phi(gate) + gate phi(gate))] (backward, optional).  Thread = 8 channels of a row.
__global__ void __launch_bounds__(256) geglu_bf16_kernel(const __nv_bfloat16* __restrict__ proj, long long rows, int C,
                                                         __nv_bfloat16* __restrict__ out, const __nv_bfloat16* __restrict__ d_out,
                                                         __nv_bfloat16* __restrict__ d_proj) {
    pdl_trigger();
    pdl_wait();
    const int cv = C >> 3;
    const long long n = rows * cv;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long r = i / cv;
        const int c = static_cast<int>(i - r * cv) * 8;
        float hv[8], gt[8];
        unpack8b(__ldg(reinterpret_cast<const uint4*>(proj + r * 2 * C + c)), hv);
        unpack8b(__ldg(reinterpret_cast<const uint4*>(proj + r * 2 * C + C + c)), gt);
        float Phi[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) Phi[e] = 0.5f * (1.0f + erff(gt[e] * 0.70710678118654752f));
        if (out) {
            float o[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] = hv[e] * gt[e] * Phi[e];
            *reinterpret_cast<uint4*>(out + r * C + c) = pack8b(o);
        }
        if (d_proj) {
            float dov[8], dh[8], dg[8];
            unpack8b(__ldg(reinterpret_cast<const uint4*>(d_out + r * C + c)), dov);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const float phi = 0.39894228040143268f * __expf(-0.5f * gt[e] * gt[e]);
                dh[e] = dov[e] * gt[e] * Phi[e];
                dg[e] = dov[e] * hv[e] * fmaf(gt[e], phi, Phi[e]);
            }
            *reinterpret_cast<uint4*>(d_proj + r * 2 * C + c) = pack8b(dh);
            *reinterpret_cast<uint4*>(d_proj + r * 2 * C + C + c) = pack8b(dg);
        }
    }
}

// ------------------------------------------------------------------------------------------------ conv_out backward (data gradient)
// conv_out is 3x3, Cin (320) -> Cout (4), fp32 NCHW output (unet_2d_condition.py:1339).  dx[b, p, ci] = sum_{co, kh, kw}
// dy[b, co, h + 1 - kh, w + 1 - kw] w[co, kh, kw, ci]: 36 MACs per element.  Thread = (pixel, 8 input channels); the 9 x Cout
// gradient taps of a pixel are fetched once per thread (L1-resident across the channel vectors of the same pixel).
// 8 consecutive channels to / from bf16 (product path) or fp32 (parity mode) storage
__device__ __forceinline__ void store8(__nv_bfloat16* p, const float (&a)[8]) { *reinterpret_cast<uint4*>(p) = pack8b(a); }
__device__ __forceinline__ void store8(float* p, const float (&a)[8]) {
    *reinterpret_cast<float4*>(p) = make_float4(a[0], a[1], a[2], a[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(a[4], a[5], a[6], a[7]);
}
__device__ __forceinline__ void load8(const __nv_bfloat16* p, float (&a)[8]) { unpack8b(__ldg(reinterpret_cast<const uint4*>(p)), a); }
__device__ __forceinline__ void load8(const float* p, float (&a)[8]) {
    const float4 u = __ldg(reinterpret_cast<const float4*>(p)), v = __ldg(reinterpret_cast<const float4*>(p + 4));
    a[0] = u.x; a[1] = u.y; a[2] = u.z; a[3] = u.w; a[4] = v.x; a[5] = v.y; a[6] = v.z; a[7] = v.w;
}

template <typename T>
__global__ void __launch_bounds__(256) conv_out_bwd_kernel(const float* __restrict__ dy, int B, int H, int W, int Cin, int Cout,
                                                           const float* __restrict__ w, T* __restrict__ dx) {
    pdl_trigger();
    pdl_wait();
    const int cv = Cin >> 3;
    const long long n = static_cast<long long>(B) * H * W * cv;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int v = static_cast<int>(i % cv);
        const long long pix = i / cv;
        const int x_ = static_cast<int>(pix % W), y_ = static_cast<int>((pix / W) % H), b = static_cast<int>(pix / (static_cast<long long>(W) * H));
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int co = 0; co < Cout; ++co)
            for (int kh = 0; kh < 3; ++kh) {
                const int yy = y_ + 1 - kh;
                if (yy < 0 || yy >= H) continue;
                for (int kw = 0; kw < 3; ++kw) {
                    const int xx = x_ + 1 - kw;
                    if (xx < 0 || xx >= W) continue;
                    const float g = __ldg(dy + ((static_cast<size_t>(b) * Cout + co) * H + yy) * W + xx);
                    const float* wp = w + ((static_cast<size_t>(co) * 3 + kh) * 3 + kw) * Cin + v * 8;
                    const float4 w0 = __ldg(reinterpret_cast<const float4*>(wp)), w1 = __ldg(reinterpret_cast<const float4*>(wp + 4));
                    acc[0] = fmaf(g, w0.x, acc[0]); acc[1] = fmaf(g, w0.y, acc[1]); acc[2] = fmaf(g, w0.z, acc[2]); acc[3] = fmaf(g, w0.w, acc[3]);
                    acc[4] = fmaf(g, w1.x, acc[4]); acc[5] = fmaf(g, w1.y, acc[5]); acc[6] = fmaf(g, w1.z, acc[6]); acc[7] = fmaf(g, w1.w, acc[7]);
                }
            }
        store8(dx + pix * Cin + v * 8, acc);
    }
}

// ------------------------------------------------------------------------------------------------ data-gradient weight repack
// dx = conv(dy, W') with W'[ci, (k*k-1-t), co] = W[co, t, ci] (flip the taps, swap the channel roles; ops.pack_conv_dgrad_weight):
// after every optimizer step the trainable layers' data-gradient plans need W' re-derived from the updated packed weight.
// One 32x32 shared-memory transpose tile per (ci tile, co tile, tap): coalesced 2-byte reads along ci and writes along co.
__global__ void __launch_bounds__(256) dgrad_repack_kernel(const __nv_bfloat16* __restrict__ w, int Cout, int Cin, int taps,
                                                           __nv_bfloat16* __restrict__ wd) {
    __shared__ __nv_bfloat16 tile[32][33];
    pdl_trigger();
    pdl_wait();
    const int t = blockIdx.z, ci0 = blockIdx.x * 32, co0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;      // 32 x 8
    const size_t ktot = static_cast<size_t>(taps) * Cin, kd = static_cast<size_t>(taps) * Cout;
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
        const int co = co0 + r, ci = ci0 + tx;
        tile[r][tx] = (co < Cout && ci < Cin) ? w[static_cast<size_t>(co) * ktot + static_cast<size_t>(t) * Cin + ci] : __float2bfloat16(0.f);
    }
    __syncthreads();
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
        const int ci = ci0 + r, co = co0 + tx;
        if (ci < Cin && co < Cout) wd[static_cast<size_t>(ci) * kd + static_cast<size_t>(taps - 1 - t) * Cout + co] = tile[tx][r];
    }
}

// ------------------------------------------------------------------------------------------------ 2x2 sum-pool
// Adjoint of the nearest-x2 replication of Upsample2D (S/models/upsampling.py:167-173): dx[b, i, j, :] = sum of the four
// high-resolution gradients du[b, 2i + {0,1}, 2j + {0,1}, :].  Thread = (low-resolution pixel, 8 channels).
template <typename T>
__global__ void __launch_bounds__(256) sumpool2x2_kernel(const T* __restrict__ du, int B, int H, int W, int C, T* __restrict__ dx) {
    pdl_trigger();
    pdl_wait();
    const int cv = C >> 3;
    const long long n = static_cast<long long>(B) * H * W * cv;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int v = static_cast<int>(i % cv);
        const long long pix = i / cv;
        const int x_ = static_cast<int>(pix % W), y_ = static_cast<int>((pix / W) % H);
        const long long b = pix / (static_cast<long long>(W) * H);
        const T* src = du + ((b * 2 * H + 2 * y_) * 2 * W + 2 * x_) * C + v * 8;
        float acc[8], f[8];
        load8(src, acc);
        load8(src + C, f);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] += f[e];
        load8(src + static_cast<size_t>(2) * W * C, f);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] += f[e];
        load8(src + static_cast<size_t>(2) * W * C + C, f);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] += f[e];
        store8(dx + pix * C + v * 8, acc);
    }
}

// y += x over n fp32 elements (parity mode: residual / skip-path gradients that the bf16 kernels fold into their last pass)
__global__ void add32_kernel(float* __restrict__ y, const float* __restrict__ x, long long n) {
    pdl_trigger();
    pdl_wait();
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * blockDim.x)
        y[i] += x[i];
}

}  // namespace mfb

using namespace mfb;

extern "C" int mfb_dgrad_repack(const void* w, int Cout, int Cin, int ksize, void* wd, void* stream) {
    MFB_REQUIRE(w && wd && Cout > 0 && Cin > 0 && (ksize == 1 || ksize == 3), "bad arguments");
    const int taps = ksize * ksize;
    dim3 grid((Cin + 31) / 32, (Cout + 31) / 32, taps);
    MFB_REQUIRE(grid.y <= 65535, "Cout too large");
    MFB_CUDA_OK(launch_k(dgrad_repack_kernel, grid, dim3(256), 0, static_cast<cudaStream_t>(stream), 1,
                         static_cast<const __nv_bfloat16*>(w), Cout, Cin, taps, static_cast<__nv_bfloat16*>(wd)));
    return MFB_OK;
}

template <typename T>
static int sumpool2x2_launch(const void* du, int B, int H, int W, int C, void* dx, void* stream) {
    MFB_REQUIRE(du && dx && B > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0, "bad arguments (C must be a multiple of 8)");
    const long long n = static_cast<long long>(B) * H * W * (C / 8);
    long long blocks = (n + 255) / 256;
    const long long cap = 32LL * (device_sm_count() > 0 ? device_sm_count() : 148);
    if (blocks > cap) blocks = cap;
    MFB_CUDA_OK(launch_k(sumpool2x2_kernel<T>, dim3(static_cast<unsigned>(blocks)), dim3(256), 0, static_cast<cudaStream_t>(stream), 1,
                         static_cast<const T*>(du), B, H, W, C, static_cast<T*>(dx)));
    return MFB_OK;
}
extern "C" int mfb_sumpool2x2(const void* du, int B, int H, int W, int C, void* dx, void* stream) {
    return sumpool2x2_launch<__nv_bfloat16>(du, B, H, W, C, dx, stream);
}
extern "C" int mfb_sumpool2x2_f32(const float* du, int B, int H, int W, int C, float* dx, void* stream) {
    return sumpool2x2_launch<float>(du, B, H, W, C, dx, stream);
}
extern "C" int mfb_add_f32(float* y, const float* x, long long n, void* stream) {
    MFB_REQUIRE(y && x && n > 0, "bad arguments");
    long long blocks = (n + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    MFB_CUDA_OK(launch_k(add32_kernel, dim3(static_cast<unsigned>(blocks)), dim3(256), 0, static_cast<cudaStream_t>(stream), 1, y, x, n));
    return MFB_OK;
}

extern "C" long long mfb_groupnorm_bwd2_ws_floats(int B, int C, int groups) {
    return (2LL * B * C + 2LL * B * GNB_MAX_CHUNKS * C + 2LL * B * groups + B + 3) / 4 * 4;      // a multiple of 16 bytes
}

extern "C" int mfb_groupnorm_bwd2(const void* x1, int C1, const void* x2, int C2, const void* dy, int B, int HW, int groups, float eps,
                                  const float* gamma, const float* beta, int silu, const float* stats, const void* dres,
                                  const void* dres2, void* dx1, void* dx2, float* dgamma, float* dbeta, float* ws, int accumulate,
                                  void* stream) {
    MFB_REQUIRE(x1 && dy && gamma && beta && dx1 && ws && stats, "null pointer");
    if (!x2) C2 = 0;
    MFB_REQUIRE(C2 == 0 || dx2 != nullptr, "x2 / dx2 disagree");
    const int C = C1 + C2;
    MFB_REQUIRE(C1 % 8 == 0 && C2 % 8 == 0 && groups > 0 && C % groups == 0 && C / 8 <= 1024, "bad geometry C1=%d C2=%d groups=%d", C1, C2, groups);
    MFB_REQUIRE(B > 0 && B <= 65535 && HW > 0, "bad geometry B=%d HW=%d", B, HW);
    const int CV = C / 8;
    const int PY = CV >= 256 ? 1 : 256 / CV;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // one full wave of CTAs, equal pixel chunks (like the forward kernels), at least 2*PY pixels per CTA
    int chunks = (2 * device_sm_count()) / B;
    const int max_chunks = (HW + 2 * PY - 1) / (2 * PY);
    if (chunks > max_chunks) chunks = max_chunks;
    if (chunks > GNB_MAX_CHUNKS) chunks = GNB_MAX_CHUNKS;
    if (chunks < 1) chunks = 1;
    const int ppc = (HW + chunks - 1) / chunks;
    chunks = (HW + ppc - 1) / ppc;
    float* ws_dg = ws;
    float* ws_db = ws + static_cast<size_t>(B) * C;
    float* part = ws + 2 * static_cast<size_t>(B) * C;
    float* gsum = part + 2 * static_cast<size_t>(B) * GNB_MAX_CHUNKS * C;
    unsigned int* counters = reinterpret_cast<unsigned int*>(gsum + 2 * static_cast<size_t>(B) * groups);
    const size_t smem = static_cast<size_t>(2) * PY * C * sizeof(float);
    static bool configured = false;
    if (!configured) {
        MFB_CUDA_OK(cudaFuncSetAttribute(gn_bwd_sums_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
        configured = true;
    }
    MFB_REQUIRE(smem <= 96 * 1024, "C too large for the reduction buffer");
    auto X1 = static_cast<const __nv_bfloat16*>(x1);
    auto X2 = static_cast<const __nv_bfloat16*>(x2);
    auto DY = static_cast<const __nv_bfloat16*>(dy);
    dim3 grid(chunks, B), block(CV, PY);
    MFB_CUDA_OK(launch_k(gn_bwd_sums_kernel, grid, block, smem, st, 1, X1, C1, X2, C2, DY, HW, groups, ppc, stats, eps, gamma, beta, silu,
                         part, ws_dg, ws_db, gsum, counters));
    MFB_CUDA_OK(launch_k(gn_bwd_apply_kernel, grid, block, 0, st, 1, X1, C1, X2, C2, DY, HW, groups, ppc, stats,
                         static_cast<const float*>(gsum), eps, gamma, beta, silu, static_cast<const __nv_bfloat16*>(dres),
                         static_cast<const __nv_bfloat16*>(dres2), static_cast<__nv_bfloat16*>(dx1), static_cast<__nv_bfloat16*>(dx2)));
    if (dgamma || dbeta)
        MFB_CUDA_OK(launch_k(gn_bwd2_final_kernel, dim3((C + 127) / 128), dim3(128), 0, st, 1, static_cast<const float*>(ws_dg),
                             static_cast<const float*>(ws_db), B, C, dgamma, dbeta, accumulate));
    return MFB_OK;
}

extern "C" int mfb_layernorm_bwd(const void* x, const void* dy, int rows, int C, float eps, const float* gamma, const void* dres, void* dx,
                                 void* stream) {
    MFB_REQUIRE(x && dy && gamma && dx, "null pointer");
    MFB_REQUIRE(C % 8 == 0 && C <= 2048 && rows > 0, "C must be a multiple of 8 and <= 2048 (got %d)", C);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int warps = 8;
    dim3 grid((rows + warps - 1) / warps), block(warps * 32);
    const int nv = (C / 8 + 31) / 32;
    auto X = static_cast<const __nv_bfloat16*>(x);
    auto DY = static_cast<const __nv_bfloat16*>(dy);
    auto DR = static_cast<const __nv_bfloat16*>(dres);
    auto DX = static_cast<__nv_bfloat16*>(dx);
    if (nv <= 2) MFB_CUDA_OK(launch_k(ln_bwd_kernel<2>, grid, block, 0, st, 1, X, DY, rows, C, eps, gamma, DR, DX));
    else if (nv <= 5) MFB_CUDA_OK(launch_k(ln_bwd_kernel<5>, grid, block, 0, st, 1, X, DY, rows, C, eps, gamma, DR, DX));
    else MFB_CUDA_OK(launch_k(ln_bwd_kernel<8>, grid, block, 0, st, 1, X, DY, rows, C, eps, gamma, DR, DX));
    return MFB_OK;
}

extern "C" int mfb_geglu(const void* proj, long long rows, int C, void* out, const void* d_out, void* d_proj, void* stream) {
    MFB_REQUIRE(proj && rows > 0 && C > 0 && C % 8 == 0, "bad arguments (C must be a multiple of 8)");
    MFB_REQUIRE((d_out == nullptr) == (d_proj == nullptr), "d_out and d_proj go together");
    MFB_REQUIRE(out || d_proj, "nothing to compute");
    const long long n = rows * (C / 8);
    long long blocks = (n + 255) / 256;
    const long long cap = 16LL * (device_sm_count() > 0 ? device_sm_count() : 148);
    if (blocks > cap) blocks = cap;
    MFB_CUDA_OK(launch_k(geglu_bf16_kernel, dim3(static_cast<unsigned>(blocks)), dim3(256), 0, static_cast<cudaStream_t>(stream), 1,
                         static_cast<const __nv_bfloat16*>(proj), rows, C, static_cast<__nv_bfloat16*>(out),
                         static_cast<const __nv_bfloat16*>(d_out), static_cast<__nv_bfloat16*>(d_proj)));
    return MFB_OK;
}

template <typename T>
static int conv_out_bwd_launch(const float* dy, int B, int H, int W, int Cin, int Cout, const float* w, void* dx, void* stream) {
    MFB_REQUIRE(dy && w && dx, "null pointer");
    MFB_REQUIRE(Cin % 8 == 0 && Cout > 0 && B > 0 && H > 0 && W > 0, "bad geometry");
    const long long n = static_cast<long long>(B) * H * W * (Cin / 8);
    long long blocks = (n + 255) / 256;
    const long long cap = 32LL * (device_sm_count() > 0 ? device_sm_count() : 148);
    if (blocks > cap) blocks = cap;
    MFB_CUDA_OK(launch_k(conv_out_bwd_kernel<T>, dim3(static_cast<unsigned>(blocks)), dim3(256), 0, static_cast<cudaStream_t>(stream), 1, dy, B,
                         H, W, Cin, Cout, w, static_cast<T*>(dx)));
    return MFB_OK;
}
extern "C" int mfb_conv_out_bwd(const float* dy, int B, int H, int W, int Cin, int Cout, const float* w, void* dx, void* stream) {
    return conv_out_bwd_launch<__nv_bfloat16>(dy, B, H, W, Cin, Cout, w, dx, stream);
}
extern "C" int mfb_conv_out_bwd_f32(const float* dy, int B, int H, int W, int Cin, int Cout, const float* w, float* dx, void* stream) {
    return conv_out_bwd_launch<float>(dy, B, H, W, Cin, Cout, w, dx, stream);
}
