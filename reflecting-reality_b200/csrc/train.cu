// Training-step glue of the BrushNet fine-tune step (BASELINE config 4): everything of
// E/train_brushnet_mirror.py:1404-1466 that is not the forward/backward of the two nets.
//   add_noise / get_velocity  (S/schedulers/scheduling_ddpm.py:501-546)      one elementwise pass
//   F.mse_loss (+ min-SNR weights, train_brushnet_mirror.py:1433-1450)        loss AND d loss / d pred in one pass, deterministic
//   clip_grad_norm_ + AdamW + the bf16 working copy of the weights            ONE pass over flat fp32 buffers
//   weight gradient of a stride-1 conv / linear (CUDA-core, fp32 accumulate)  the partner of the tcgen05 data-gradient
// Layout decision (B200-first): all trainable tensors of the BrushNet branch live in ONE flat fp32 master buffer with flat
// gradient / moment buffers of the same shape (2.5 GB each at 618.8 M parameters — trivial against 180 GB), so the gradient
// all-reduce is a handful of large NCCL calls over contiguous memory and clip + AdamW + re-quantisation is a single launch
// instead of ~700 per-tensor launches x 3.  All reductions are fixed-order (no float atomics): a step is bit-reproducible.
#include <cuda_bf16.h>

#include "common.h"
#include "ptx.cuh"

namespace mfb {

// ---------------------------------------------------------------------------------------------- add_noise / velocity
// grid (chunks, B): the two per-sample scalars are computed once per thread from alphas_cumprod[t_b].
__global__ void add_noise_kernel(const float* __restrict__ x0, const float* __restrict__ noise, const long long* __restrict__ t,
                                 const float* __restrict__ acp, int T, long long n, float* __restrict__ noisy,
                                 float* __restrict__ velocity) {
    pdl_trigger();
    pdl_wait();
    const int b = blockIdx.y;
    long long ti = t[b];
    ti = ti < 0 ? 0 : (ti >= T ? T - 1 : ti);
    const float a = acp[ti];
    const float sa = sqrtf(a), so = sqrtf(1.0f - a);   // alphas_cumprod[t] ** 0.5, (1 - alphas_cumprod[t]) ** 0.5
    const size_t base = static_cast<size_t>(b) * n;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const float xv = x0[base + i], nv = noise[base + i];
        if (noisy) noisy[base + i] = sa * xv + so * nv;            // scheduling_ddpm.py:524
        if (velocity) velocity[base + i] = sa * nv - so * xv;      // scheduling_ddpm.py:545
    }
}

// ---------------------------------------------------------------------------------------------- block reduction (fixed order)
template <typename T>
__device__ __forceinline__ T block_sum(T v, T* sm /* >= 32 */) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) sm[warp] = v;
    __syncthreads();
    T r = 0;
    if (warp == 0) {
        r = lane < (blockDim.x + 31) / 32 ? sm[lane] : T(0);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
    }
    return r;  // valid in warp 0
}

// ---------------------------------------------------------------------------------------------- MSE loss + its gradient
// pass 1, grid (chunks, B): partial[b][c] = sum over the chunk of (pred - target)^2; also writes
// grad = 2 (pred - target) * w_b / (B n)  (the derivative of  mean_b( w_b * mean_n (p - t)^2 ) ).
__global__ void mse_partial_kernel(const float* __restrict__ pred, const float* __restrict__ target, const float* __restrict__ wts,
                                   int B, long long n, float* __restrict__ grad, float* __restrict__ partial) {
    pdl_trigger();
    pdl_wait();
    __shared__ float sm[32];
    const int b = blockIdx.y;
    const size_t base = static_cast<size_t>(b) * n;
    const float gs = 2.0f * (wts ? wts[b] : 1.0f) / (static_cast<float>(B) * static_cast<float>(n));
    const long long per = (n + gridDim.x - 1) / gridDim.x;
    const long long lo = blockIdx.x * per, hi = (lo + per < n) ? lo + per : n;
    float acc = 0.f;
    for (long long i = lo + threadIdx.x; i < hi; i += blockDim.x) {
        const float d = pred[base + i] - target[base + i];
        acc += d * d;
        if (grad) grad[base + i] = gs * d;
    }
    const float s = block_sum(acc, sm);
    if (threadIdx.x == 0) partial[static_cast<size_t>(b) * gridDim.x + blockIdx.x] = s;
}

// pass 2, one CTA: per_sample[b] = (sum_c partial[b][c]) / n ; loss = (sum_b w_b * per_sample[b]) / B.  Serial, fixed order.
__global__ void mse_final_kernel(const float* __restrict__ partial, int chunks, const float* __restrict__ wts, int B, long long n,
                                 float* __restrict__ per_sample, float* __restrict__ loss) {
    pdl_trigger();
    pdl_wait();
    __shared__ double sm[32];
    double acc = 0.0;
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
        double s = 0.0;
        for (int c = 0; c < chunks; ++c) s += static_cast<double>(partial[static_cast<size_t>(b) * chunks + c]);
        const double ps = s / static_cast<double>(n);
        if (per_sample) per_sample[b] = static_cast<float>(ps);
        acc += ps * (wts ? static_cast<double>(wts[b]) : 1.0);
    }
    const double tot = block_sum(acc, sm);
    if (threadIdx.x == 0) *loss = static_cast<float>(tot / static_cast<double>(B));
}

// ---------------------------------------------------------------------------------------------- squared L2 norm of a flat buffer
__global__ void sqnorm_partial_kernel(const float* __restrict__ g, long long n, float* __restrict__ partial) {
    pdl_trigger();
    pdl_wait();
    __shared__ float sm[32];
    float acc = 0.f;
    const long long n4 = n >> 2;
    const float4* g4 = reinterpret_cast<const float4*>(g);
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n4;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const float4 v = g4[i];
        acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
        const float v = g[(n4 << 2) + threadIdx.x];
        acc += v * v;
    }
    const float s = block_sum(acc, sm);
    if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

__global__ void sqnorm_final_kernel(const float* __restrict__ partial, int m, float* __restrict__ out, int accumulate) {
    pdl_trigger();
    pdl_wait();
    __shared__ double sm[32];
    double acc = 0.0;
    for (int i = threadIdx.x; i < m; i += blockDim.x) acc += static_cast<double>(partial[i]);
    const double tot = block_sum(acc, sm);
    if (threadIdx.x == 0) *out = static_cast<float>(tot + (accumulate ? static_cast<double>(*out) : 0.0));
}

// ---------------------------------------------------------------------------------------------- clip + AdamW + bf16 working copy
// hyper (device, 12 floats): lr, beta1, beta2, eps, weight_decay, bias_correction1 = 1 - beta1^step,
// sqrt(bias_correction2) = sqrt(1 - beta2^step), grad_scale (1 / world size after a SUM all-reduce, 1 / accumulation ...),
// then the derived scalars torch evaluates in Python float64 before they reach a kernel: 1 - beta1, 1 - beta2,
// 1 - lr*weight_decay, lr / bias_correction1  (1 - 0.999f evaluated in fp32 is off by 1.3e-5: it has to come from the host).
// torch.optim.AdamW single-tensor update order (torch/optim/adamw.py _single_tensor_adamw):
//   p *= 1 - lr*wd ; m = lerp(m, g, 1-b1) ; v = b2*v + (1-b2) g^2 ; p -= (lr/bc1) * m / (sqrt(v)/sqrt(bc2) + eps)
// clip_grad_norm_ (torch/nn/utils/clip_grad.py): g *= min(1, max_norm / (total_norm + 1e-6)).
struct AdamScalars {
    float lr, b1, b2, eps, wd, bc1, sbc2, gscale, omb1, omb2, decay, step_size;
};

__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, const AdamScalars& h) {
    p *= h.decay;
    m += h.omb1 * (g - m);
    v = v * h.b2 + h.omb2 * (g * g);
    const float denom = sqrtf(v) / h.sbc2 + h.eps;
    p -= h.step_size * (m / denom);
}

__global__ void adamw_kernel(float* __restrict__ param, const float* __restrict__ grad, float* __restrict__ m, float* __restrict__ v,
                             __nv_bfloat16* __restrict__ pbf, long long n, const float* __restrict__ hyper,
                             const float* __restrict__ sqnorm, float max_norm) {
    pdl_trigger();
    pdl_wait();
    AdamScalars h{hyper[0], hyper[1], hyper[2], hyper[3], hyper[4], hyper[5], hyper[6], hyper[7], hyper[8], hyper[9], hyper[10], hyper[11]};
    float gmul = h.gscale;
    if (sqnorm && max_norm > 0.f) {
        const float total = sqrtf(*sqnorm) * fabsf(h.gscale);       // norm of the SCALED gradient
        gmul *= fminf(1.0f, max_norm / (total + 1e-6f));
    }
    const long long n4 = n >> 2;
    float4* p4 = reinterpret_cast<float4*>(param);
    const float4* g4 = reinterpret_cast<const float4*>(grad);
    float4* m4 = reinterpret_cast<float4*>(m);
    float4* v4 = reinterpret_cast<float4*>(v);
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n4;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        float4 p = p4[i], mm = m4[i], vv = v4[i];
        const float4 g = g4[i];
        adam_one(p.x, g.x * gmul, mm.x, vv.x, h);
        adam_one(p.y, g.y * gmul, mm.y, vv.y, h);
        adam_one(p.z, g.z * gmul, mm.z, vv.z, h);
        adam_one(p.w, g.w * gmul, mm.w, vv.w, h);
        p4[i] = p;
        m4[i] = mm;
        v4[i] = vv;
        if (pbf) {
            __nv_bfloat162 lo = __floats2bfloat162_rn(p.x, p.y), hi = __floats2bfloat162_rn(p.z, p.w);
            uint2 u;
            u.x = *reinterpret_cast<uint32_t*>(&lo);
            u.y = *reinterpret_cast<uint32_t*>(&hi);
            reinterpret_cast<uint2*>(pbf)[i] = u;
        }
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
        const long long i = (n4 << 2) + threadIdx.x;
        float p = param[i], mm = m[i], vv = v[i];
        adam_one(p, grad[i] * gmul, mm, vv, h);
        param[i] = p;
        m[i] = mm;
        v[i] = vv;
        if (pbf) pbf[i] = __float2bfloat16_rn(p);
    }
}

// ---------------------------------------------------------------------------------------------- conv / linear weight gradient
// dW[co][(t, ci)] (+)= sum over pixels p of dy[p][co] * x[p + offset_t][ci]   (stride 1, padding k/2; NHWC operands).
// As a GEMM: M = Cout, N = taps*Cin, K = B*H*W (the reduction runs over PIXELS, so both operands are "transposed" with
// respect to the forward implicit GEMM).  CUDA-core version: CTA = 64 co x 64 ci of one tap, 16-pixel K slabs staged in
// smem as fp32, 4x4 register micro-tile per thread, ONE CTA per output tile walking all pixels in order => deterministic.
// The bias gradient (column sums of dy) rides along in the CTAs of tap 0 / ci block 0.
template <typename T>
__device__ __forceinline__ float ld_as_float(const T* p);
template <>
__device__ __forceinline__ float ld_as_float<float>(const float* p) { return *p; }
template <>
__device__ __forceinline__ float ld_as_float<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }

constexpr int WG_T = 64, WG_K = 16;

template <typename T>
__global__ void __launch_bounds__(256) wgrad_kernel(const T* __restrict__ x, const T* __restrict__ dy, int B, int H, int W, int Cin,
                                                    int Cout, int ksize, float* __restrict__ dw, float* __restrict__ dbias,
                                                    int accumulate) {
    pdl_trigger();
    pdl_wait();
    __shared__ __align__(16) float sdy[WG_K][WG_T];
    __shared__ __align__(16) float sx[WG_K][WG_T];
    const int ci0 = blockIdx.x * WG_T, co0 = blockIdx.y * WG_T, tap = blockIdx.z;
    const int pad = ksize / 2;
    const int dh = tap / ksize - pad, dwo = tap % ksize - pad;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const long long P = static_cast<long long>(B) * H * W;
    const int ktot = ksize * ksize * Cin;
    float acc[4][4] = {};
    float bsum[4] = {};
    const bool do_bias = dbias != nullptr && blockIdx.x == 0 && tap == 0 && tx == 0;
    for (long long p0 = 0; p0 < P; p0 += WG_K) {
#pragma unroll
        for (int r = 0; r < (WG_K * WG_T) / 256; ++r) {
            const int e = r * 256 + threadIdx.x;
            const int k = e >> 6, c = e & 63;
            const long long p = p0 + k;
            float a = 0.f, bv = 0.f;
            if (p < P) {
                if (co0 + c < Cout) a = ld_as_float(dy + p * Cout + co0 + c);
                const int w_ = static_cast<int>(p % W), h_ = static_cast<int>((p / W) % H);
                const int hs = h_ + dh, ws = w_ + dwo;
                if (hs >= 0 && hs < H && ws >= 0 && ws < W && ci0 + c < Cin)
                    bv = ld_as_float(x + (p + static_cast<long long>(dh) * W + dwo) * Cin + ci0 + c);
            }
            sdy[k][c] = a;
            sx[k][c] = bv;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < WG_K; ++k) {
            const float4 a = *reinterpret_cast<const float4*>(&sdy[k][ty * 4]);
            const float4 b = *reinterpret_cast<const float4*>(&sx[k][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
                bsum[i] += av[i];
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int co = co0 + ty * 4 + i;
        if (co >= Cout) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int ci = ci0 + tx * 4 + j;
            if (ci >= Cin) continue;
            float* o = dw + static_cast<size_t>(co) * ktot + static_cast<size_t>(tap) * Cin + ci;
            *o = acc[i][j] + (accumulate ? *o : 0.f);
        }
        if (do_bias) dbias[co] = bsum[i] + (accumulate ? dbias[co] : 0.f);
    }
}

// ---------------------------------------------------------------------------------------------- GroupNorm (+SiLU) backward
// y = silu(gamma * xhat + beta), xhat = (x - mean) * rstd over the (HW x C/groups) slab of (image, group); NHWC, x optionally the
// channel concat of two tensors (the up blocks' skip concat, as in the forward kernel).  With dz = dy * silu'(z):
//   dgamma[c] = sum dz * xhat ; dbeta[c] = sum dz ; dx = rstd * (dz*gamma - mean_slab(dz*gamma) - xhat * mean_slab(dz*gamma*xhat))
// One CTA per (group, image).  Thread t owns channel t % cpg and every (blockDim / cpg)-th pixel, so the per-channel sums
// are private until a fixed-order smem reduction; slab sums in fp64; statistics are recomputed from x (nothing is saved by the
// forward pass).  Three reads of x, two of dy, one write: a first CUDA-core version (correct and deterministic, not yet tuned).
template <typename T>
__device__ __forceinline__ void st_from_float(T* p, float v);
template <>
__device__ __forceinline__ void st_from_float<float>(float* p, float v) { *p = v; }
template <>
__device__ __forceinline__ void st_from_float<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

__device__ __forceinline__ double block_sum_all(double v, double* sm /* >= 33 */) {
    const double r = block_sum(v, sm);
    if (threadIdx.x == 0) sm[32] = r;
    __syncthreads();
    return sm[32];
}

template <typename T>
__global__ void __launch_bounds__(256) gn_bwd_kernel(const T* __restrict__ x1, int C1, const T* __restrict__ x2, int C2,
                                                     const T* __restrict__ dy, int HW, int groups, float eps,
                                                     const float* __restrict__ gamma, const float* __restrict__ beta, int silu,
                                                     T* __restrict__ dx1, T* __restrict__ dx2, float* __restrict__ ws_dgamma,
                                                     float* __restrict__ ws_dbeta) {
    pdl_trigger();
    pdl_wait();
    __shared__ double smd[33];
    __shared__ float red[2][256];
    const int C = C1 + C2, cpg = C / groups;
    const int b = blockIdx.y, c0 = blockIdx.x * cpg;
    const int rows = blockDim.x / cpg;
    const bool active = static_cast<int>(threadIdx.x) < rows * cpg;
    const int cl = threadIdx.x % cpg, r = threadIdx.x / cpg, c = c0 + cl;
    const bool in1 = c < C1;
    const T* xs = in1 ? x1 + static_cast<size_t>(b) * HW * C1 + c : x2 + static_cast<size_t>(b) * HW * C2 + (c - C1);
    T* dxs = in1 ? dx1 + static_cast<size_t>(b) * HW * C1 + c : dx2 + static_cast<size_t>(b) * HW * C2 + (c - C1);
    const int ldx = in1 ? C1 : C2;
    const T* dys = dy + static_cast<size_t>(b) * HW * C + c;
    const double m = static_cast<double>(HW) * cpg;

    double s = 0.0, ss = 0.0;
    if (active)
        for (int p = r; p < HW; p += rows) {
            const double v = ld_as_float(xs + static_cast<size_t>(p) * ldx);
            s += v;
            ss += v * v;
        }
    const double mean_d = block_sum_all(s, smd) / m;
    __syncthreads();
    const double var_d = fmax(block_sum_all(ss, smd) / m - mean_d * mean_d, 0.0);
    __syncthreads();
    const float mean = static_cast<float>(mean_d);
    const float rstd = static_cast<float>(1.0 / sqrt(var_d + static_cast<double>(eps)));
    const float gm = active ? gamma[c] : 0.f, bt = active ? beta[c] : 0.f;

    float dg = 0.f, db = 0.f;
    double s1 = 0.0, s2 = 0.0;
    if (active)
        for (int p = r; p < HW; p += rows) {
            const float xh = (ld_as_float(xs + static_cast<size_t>(p) * ldx) - mean) * rstd;
            float dz = ld_as_float(dys + static_cast<size_t>(p) * C);
            if (silu) {
                const float z = gm * xh + bt;
                const float sg = 1.0f / (1.0f + expf(-z));
                dz *= sg * (1.0f + z * (1.0f - sg));
            }
            dg += dz * xh;
            db += dz;
            s1 += static_cast<double>(dz * gm);
            s2 += static_cast<double>(dz * gm * xh);
        }
    const float S1 = static_cast<float>(block_sum_all(s1, smd) / m);
    __syncthreads();
    const float S2 = static_cast<float>(block_sum_all(s2, smd) / m);
    red[0][threadIdx.x] = dg;
    red[1][threadIdx.x] = db;
    __syncthreads();
    if (static_cast<int>(threadIdx.x) < cpg) {
        float a = 0.f, bsum = 0.f;
        for (int q = 0; q < rows; ++q) {
            a += red[0][q * cpg + threadIdx.x];
            bsum += red[1][q * cpg + threadIdx.x];
        }
        ws_dgamma[static_cast<size_t>(b) * C + c0 + threadIdx.x] = a;
        ws_dbeta[static_cast<size_t>(b) * C + c0 + threadIdx.x] = bsum;
    }
    if (active)
        for (int p = r; p < HW; p += rows) {
            const float xh = (ld_as_float(xs + static_cast<size_t>(p) * ldx) - mean) * rstd;
            float dz = ld_as_float(dys + static_cast<size_t>(p) * C);
            if (silu) {
                const float z = gm * xh + bt;
                const float sg = 1.0f / (1.0f + expf(-z));
                dz *= sg * (1.0f + z * (1.0f - sg));
            }
            st_from_float(dxs + static_cast<size_t>(p) * ldx, rstd * (dz * gm - S1 - xh * S2));
        }
}

// dgamma[c] (+)= sum_b ws[b][c] in image order (same for dbeta): thread per channel
__global__ void gn_bwd_final_kernel(const float* __restrict__ ws_dgamma, const float* __restrict__ ws_dbeta, int B, int C,
                                    float* __restrict__ dgamma, float* __restrict__ dbeta, int accumulate) {
    pdl_trigger();
    pdl_wait();
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float a = 0.f, bsum = 0.f;
    for (int b = 0; b < B; ++b) {
        a += ws_dgamma[static_cast<size_t>(b) * C + c];
        bsum += ws_dbeta[static_cast<size_t>(b) * C + c];
    }
    if (dgamma) dgamma[c] = a + (accumulate ? dgamma[c] : 0.f);
    if (dbeta) dbeta[c] = bsum + (accumulate ? dbeta[c] : 0.f);
}

}  // namespace mfb

using namespace mfb;

static inline int chunks_for(long long n, int block, int cap) {
    long long g = (n + block - 1) / block;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return static_cast<int>(g);
}

extern "C" int mfb_add_noise(const float* x0, const float* noise, const long long* timesteps, const float* alphas_cumprod,
                             int num_train_timesteps, int B, long long n, float* noisy, float* velocity, void* stream) {
    MFB_REQUIRE(x0 && noise && timesteps && alphas_cumprod && (noisy || velocity), "null pointer");
    MFB_REQUIRE(B > 0 && B <= 65535 && n > 0 && num_train_timesteps > 0, "bad geometry B=%d n=%lld T=%d", B, n, num_train_timesteps);
    const int chunks = chunks_for(n, 256 * 4, 148 * 8 / (B < 148 * 8 ? B : 148 * 8) + 1);
    MFB_CUDA_OK(launch_k(add_noise_kernel, dim3(chunks, B), dim3(256), 0, static_cast<cudaStream_t>(stream), 1, x0, noise,
                         timesteps, alphas_cumprod, num_train_timesteps, n, noisy, velocity));
    return MFB_OK;
}

extern "C" int mfb_mse_loss(const float* pred, const float* target, const float* weights, int B, long long n, float* per_sample,
                            float* loss, float* grad, float* ws, void* stream) {
    MFB_REQUIRE(pred && target && loss && ws, "null pointer");
    MFB_REQUIRE(B > 0 && B <= 65535 && n > 0, "bad geometry B=%d n=%lld", B, n);
    const int chunks = chunks_for(n, 256 * 8, MFB_MSE_MAX_CHUNKS);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    MFB_CUDA_OK(launch_k(mse_partial_kernel, dim3(chunks, B), dim3(256), 0, st, 1, pred, target, weights, B, n, grad, ws));
    MFB_CUDA_OK(launch_k(mse_final_kernel, dim3(1), dim3(256), 0, st, 1, static_cast<const float*>(ws), chunks, weights, B, n,
                         per_sample, loss));
    return MFB_OK;
}

extern "C" int mfb_grad_sqnorm(const float* g, long long n, float* ws, float* out_sq, int accumulate, void* stream) {
    MFB_REQUIRE(g && ws && out_sq && n > 0, "null pointer / empty buffer");
    MFB_REQUIRE((reinterpret_cast<uintptr_t>(g) & 15) == 0, "gradient buffer must be 16-byte aligned");
    const int grid = chunks_for(n >> 2, 256, MFB_SQNORM_WS_FLOATS);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    MFB_CUDA_OK(launch_k(sqnorm_partial_kernel, dim3(grid), dim3(256), 0, st, 1, g, n, ws));
    MFB_CUDA_OK(launch_k(sqnorm_final_kernel, dim3(1), dim3(256), 0, st, 1, static_cast<const float*>(ws), grid, out_sq, accumulate));
    return MFB_OK;
}

extern "C" int mfb_adamw_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, void* param_bf16, long long n,
                              const float* hyper, const float* grad_sqnorm, float max_grad_norm, void* stream) {
    MFB_REQUIRE(param && grad && exp_avg && exp_avg_sq && hyper && n > 0, "null pointer / empty buffer");
    MFB_REQUIRE(((reinterpret_cast<uintptr_t>(param) | reinterpret_cast<uintptr_t>(grad) | reinterpret_cast<uintptr_t>(exp_avg) |
                  reinterpret_cast<uintptr_t>(exp_avg_sq)) & 15) == 0 && (reinterpret_cast<uintptr_t>(param_bf16) & 7) == 0,
                "flat buffers must be 16-byte aligned (bf16 copy: 8)");
    const int grid = chunks_for(n >> 2, 256, 148 * 8);
    MFB_CUDA_OK(launch_k(adamw_kernel, dim3(grid), dim3(256), 0, static_cast<cudaStream_t>(stream), 1, param, grad, exp_avg,
                         exp_avg_sq, static_cast<__nv_bfloat16*>(param_bf16), n, hyper, grad_sqnorm, max_grad_norm));
    return MFB_OK;
}

extern "C" int mfb_conv_wgrad(const void* x, const void* dy, int dtype, int B, int H, int W, int Cin, int Cout, int ksize,
                              float* dw, float* dbias, int accumulate, void* stream) {
    MFB_REQUIRE(x && dy && dw, "null pointer");
    MFB_REQUIRE((ksize == 1 || ksize == 3) && B > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0, "bad geometry");
    MFB_REQUIRE(dtype == 0 || dtype == 1, "dtype must be 0 (bf16) or 1 (fp32)");
    const dim3 grid((Cin + WG_T - 1) / WG_T, (Cout + WG_T - 1) / WG_T, ksize * ksize);
    MFB_REQUIRE(grid.y <= 65535, "Cout too large");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (dtype == 1) {
        MFB_CUDA_OK(launch_k(wgrad_kernel<float>, grid, dim3(256), 0, st, 1, static_cast<const float*>(x),
                             static_cast<const float*>(dy), B, H, W, Cin, Cout, ksize, dw, dbias, accumulate));
    } else {
        MFB_CUDA_OK(launch_k(wgrad_kernel<__nv_bfloat16>, grid, dim3(256), 0, st, 1, static_cast<const __nv_bfloat16*>(x),
                             static_cast<const __nv_bfloat16*>(dy), B, H, W, Cin, Cout, ksize, dw, dbias, accumulate));
    }
    return MFB_OK;
}

extern "C" int mfb_groupnorm_bwd(const void* x1, int C1, const void* x2, int C2, const void* dy, int dtype, int B, int HW, int groups,
                                 float eps, const float* gamma, const float* beta, int silu, void* dx1, void* dx2, float* dgamma,
                                 float* dbeta, float* ws, int accumulate, void* stream) {
    MFB_REQUIRE(x1 && dy && gamma && beta && dx1 && ws, "null pointer");
    MFB_REQUIRE(dtype == 0 || dtype == 1, "dtype must be 0 (bf16) or 1 (fp32)");
    MFB_REQUIRE((C2 == 0) == (x2 == nullptr) && (C2 == 0 || dx2 != nullptr), "x2 / dx2 / C2 disagree");
    const int C = C1 + C2;
    MFB_REQUIRE(B > 0 && B <= 65535 && HW > 0 && groups > 0 && C % groups == 0 && C / groups <= 256, "bad geometry B=%d HW=%d C=%d groups=%d",
                B, HW, C, groups);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    float* wg = ws;
    float* wb = ws + static_cast<size_t>(B) * C;
    if (dtype == 1) {
        MFB_CUDA_OK(launch_k(gn_bwd_kernel<float>, dim3(groups, B), dim3(256), 0, st, 1, static_cast<const float*>(x1), C1,
                             static_cast<const float*>(x2), C2, static_cast<const float*>(dy), HW, groups, eps, gamma, beta, silu,
                             static_cast<float*>(dx1), static_cast<float*>(dx2), wg, wb));
    } else {
        MFB_CUDA_OK(launch_k(gn_bwd_kernel<__nv_bfloat16>, dim3(groups, B), dim3(256), 0, st, 1,
                             static_cast<const __nv_bfloat16*>(x1), C1, static_cast<const __nv_bfloat16*>(x2), C2,
                             static_cast<const __nv_bfloat16*>(dy), HW, groups, eps, gamma, beta, silu,
                             static_cast<__nv_bfloat16*>(dx1), static_cast<__nv_bfloat16*>(dx2), wg, wb));
    }
    if (dgamma || dbeta) {
        MFB_CUDA_OK(launch_k(gn_bwd_final_kernel, dim3((C + 127) / 128), dim3(128), 0, st, 1, static_cast<const float*>(wg),
                             static_cast<const float*>(wb), B, C, dgamma, dbeta, accumulate));
    }
    return MFB_OK;
}
