// fp32 parity mode of the denoise step (BASELINE.json configs[0]: "fp32 on CPU" is the reference's own correctness
// anchor; north_star asks for rel-L2 1e-4 per step against it).  The SAME host program (engine.py: same fusions, same
// K-segments, same tap folding, same buffers) runs with fp32 storage and these CUDA-core kernels instead of the bf16
// tcgen05 ones, selected by tensor dtype.  What the mode proves is the wiring of the path — every fusion and layout is
// shared — at a precision where a wrong epsilon, a dropped bias or a mis-ordered segment cannot hide behind bf16
// rounding.  Not a performance path: plain shared-memory-tiled FFMA, ~1 % of the tensor-core rate.
//
// Reference semantics per kernel: see the bf16 counterparts (igemm.cu, norm.cu, attn.cu, misc.cu) — the file:line
// citations there apply unchanged.
#include <math.h>
#include <string.h>

#include "fp32mode.h"

namespace mfb {

// ------------------------------------------------------------------------------------------------ conv / linear
constexpr int C32_BM = 64, C32_BN = 128, C32_BK = 16;

__global__ void __launch_bounds__(256) conv32_kernel(const Conv32Params q) {
    __shared__ float As[C32_BK][C32_BM + 4];
    __shared__ float Bs[C32_BK][C32_BN + 4];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.x * C32_BM, n0 = blockIdx.y * C32_BN;
    const int M = q.B * q.Ho * q.Wo;

    // A-load role: row = tid / 4, 4 consecutive k (channels) = (tid % 4) * 4
    const int ar = tid >> 2, ak = (tid & 3) * 4;
    const int am = m0 + ar;
    const bool arow_ok = am < M;
    const int a_b = arow_ok ? am / (q.Ho * q.Wo) : 0;
    const int a_oh = arow_ok ? (am / q.Wo) % q.Ho : 0;
    const int a_ow = arow_ok ? am % q.Wo : 0;
    // B-load role: col = tid / 2, 8 consecutive k = (tid % 2) * 8
    const int bc = tid >> 1, bk = (tid & 1) * 8;
    const bool bcol_ok = n0 + bc < q.Cout;
    const float* wrow = q.w + static_cast<size_t>(bcol_ok ? n0 + bc : 0) * q.ktot;

    float acc[4][8];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    int kglob = 0;
    const int nseg = q.ntaps + q.n_extra;
    for (int s = 0; s < nseg; ++s) {
        const float* src;
        int C;
        bool pix_ok = arow_ok;
        size_t pix = 0;
        if (s < q.ntaps) {
            src = q.x; C = q.Cin;
            const int ih = a_oh * q.stride + q.dh[s], iw = a_ow * q.stride + q.dw[s];
            pix_ok = pix_ok && ih >= 0 && ih < q.Hin && iw >= 0 && iw < q.Win;
            pix = (static_cast<size_t>(a_b) * q.Hin + (pix_ok ? ih : 0)) * q.Win + (pix_ok ? iw : 0);
        } else {
            src = q.ex[s - q.ntaps]; C = q.exC[s - q.ntaps];
            pix = (static_cast<size_t>(a_b) * q.Hf + a_oh * q.o_step + q.o_py) * q.Wf + a_ow * q.o_step + q.o_px;
        }
        const float* arow = src + pix * C;
        for (int c0 = 0; c0 < C; c0 += C32_BK, kglob += C32_BK) {
            float4 av = make_float4(0.f, 0.f, 0.f, 0.f);
            if (pix_ok) av = *reinterpret_cast<const float4*>(arow + c0 + ak);
            float4 b0 = make_float4(0.f, 0.f, 0.f, 0.f), b1 = b0;
            if (bcol_ok) {
                b0 = *reinterpret_cast<const float4*>(wrow + kglob + bk);
                b1 = *reinterpret_cast<const float4*>(wrow + kglob + bk + 4);
            }
            __syncthreads();
            As[ak + 0][ar] = av.x; As[ak + 1][ar] = av.y; As[ak + 2][ar] = av.z; As[ak + 3][ar] = av.w;
            Bs[bk + 0][bc] = b0.x; Bs[bk + 1][bc] = b0.y; Bs[bk + 2][bc] = b0.z; Bs[bk + 3][bc] = b0.w;
            Bs[bk + 4][bc] = b1.x; Bs[bk + 5][bc] = b1.y; Bs[bk + 6][bc] = b1.z; Bs[bk + 7][bc] = b1.w;
            __syncthreads();
#pragma unroll
            for (int k = 0; k < C32_BK; ++k) {
                const float4 a4 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
                const float4 v0 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
                const float4 v1 = *reinterpret_cast<const float4*>(&Bs[k][64 + tx * 4]);
                const float a[4] = {a4.x, a4.y, a4.z, a4.w};
                const float b[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
            }
        }
    }

    // epilogue: (acc + bias + rowbias) * alpha + res1 + res2, or GEGLU over the [64 value | 64 gate] column tile
    const float alpha = q.alpha ? *q.alpha : 1.0f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= M) continue;
        const int b = m / (q.Ho * q.Wo), oh = (m / q.Wo) % q.Ho, ow = m % q.Wo;
        const size_t opix = (static_cast<size_t>(b) * q.Hf + oh * q.o_step + q.o_py) * q.Wf + ow * q.o_step + q.o_px;
        if (q.geglu) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int nv = n0 + tx * 4 + j, ng = nv + 64;
                if (ng >= q.Cout) continue;
                const float val = acc[i][j] + (q.bias ? q.bias[nv] : 0.f);
                const float gate = acc[i][4 + j] + (q.bias ? q.bias[ng] : 0.f);
                const float gelu = 0.5f * gate * (1.0f + erff(gate * 0.70710678118654752f));   // exact (erf) GELU
                q.out[opix * q.out_ld + (n0 >> 1) + tx * 4 + j] = val * gelu;
            }
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int n = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
                if (n >= q.Cout) continue;
                float v = acc[i][j];
                if (q.bias) v += q.bias[n];
                if (q.rowbias) v += q.rowbias[static_cast<size_t>(b) * q.rowbias_ld + n];
                v *= alpha;
                if (q.res1) v += q.res1[opix * q.out_ld + n];
                if (q.res2) v += q.res2[opix * q.out_ld + n];
                q.out[opix * q.out_ld + n] = v;
            }
        }
    }
}

int conv32_build(const mfb_conv_desc* d, int up_py, int up_px, const void* w, Conv32Params& q) {
    memset(&q, 0, sizeof(q));
    const bool up = up_py >= 0;
    q.x = static_cast<const float*>(d->x);
    q.B = d->B; q.Hin = d->H; q.Win = d->W; q.Cin = d->Cin;
    q.stride = up ? 1 : d->stride;
    if (up) {
        // sub-pixel phase of conv3x3(nearest2x(x)): 2x2 taps over the low-resolution input (ops.pack_upconv_weight)
        q.ntaps = 4;
        for (int ty = 0; ty < 2; ++ty)
            for (int tx = 0; tx < 2; ++tx) { q.dh[ty * 2 + tx] = up_py - 1 + ty; q.dw[ty * 2 + tx] = up_px - 1 + tx; }
        q.Ho = d->H; q.Wo = d->W; q.o_step = 2; q.o_py = up_py; q.o_px = up_px; q.Hf = 2 * d->H; q.Wf = 2 * d->W;
    } else {
        q.ntaps = d->ksize * d->ksize;
        for (int kh = 0; kh < d->ksize; ++kh)
            for (int kw = 0; kw < d->ksize; ++kw) {
                q.dh[kh * d->ksize + kw] = kh - (d->pad0 ? 0 : d->ksize / 2);      // pad0: zero row/column at the bottom/right
                q.dw[kh * d->ksize + kw] = kw - (d->pad0 ? 0 : d->ksize / 2);
            }
        q.Ho = (d->H + d->stride - 1) / d->stride; q.Wo = (d->W + d->stride - 1) / d->stride;
        q.o_step = 1; q.o_py = 0; q.o_px = 0; q.Hf = q.Ho; q.Wf = q.Wo;
    }
    q.n_extra = d->n_extra;
    q.ktot = q.ntaps * d->Cin;
    for (int e = 0; e < d->n_extra; ++e) {
        MFB_REQUIRE(d->extra_C[e] % C32_BK == 0, "extra segment channels must be a multiple of 16");
        q.ex[e] = static_cast<const float*>(d->extra_x[e]);
        q.exC[e] = d->extra_C[e];
        q.ktot += d->extra_C[e];
    }
    q.w = static_cast<const float*>(w);
    q.Cout = d->Cout;
    q.bias = d->bias; q.rowbias = d->rowbias; q.rowbias_ld = d->rowbias_ld; q.alpha = d->alpha;
    q.res1 = static_cast<const float*>(d->res1); q.res2 = static_cast<const float*>(d->res2);
    q.out = static_cast<float*>(d->out);
    q.out_ld = d->geglu ? d->Cout / 2 : d->Cout;
    q.geglu = d->geglu;
    return MFB_OK;
}

int conv32_launch(const Conv32Params& q, cudaStream_t st) {
    const int M = q.B * q.Ho * q.Wo;
    dim3 grid((M + C32_BM - 1) / C32_BM, (q.Cout + C32_BN - 1) / C32_BN);
    conv32_kernel<<<grid, 256, 0, st>>>(q);
    MFB_CUDA_OK(cudaGetLastError());
    return MFB_OK;
}

// ------------------------------------------------------------------------------------------------ GroupNorm (+SiLU)
// one CTA per (image, group); two passes (mean, then centred second moment) with fp64 block sums
__device__ double block_sum(double v, double* sh) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    for (int i = 0; i < int(blockDim.x >> 5); ++i) t += sh[i];
    return t;
}

__global__ void __launch_bounds__(256) gn32_kernel(const float* __restrict__ x1, int C1, const float* __restrict__ x2, int C2, int HW,
                                                   int groups, float eps, const float* __restrict__ gamma,
                                                   const float* __restrict__ beta, int silu, float* __restrict__ out) {
    __shared__ double sh[8];
    const int b = blockIdx.y, g = blockIdx.x;
    const int C = C1 + C2, Cg = C / groups;
    const int n = HW * Cg;
    auto load = [&](int i) {
        const int p = i / Cg, c = g * Cg + i % Cg;
        return c < C1 ? x1[(static_cast<size_t>(b) * HW + p) * C1 + c] : x2[(static_cast<size_t>(b) * HW + p) * C2 + (c - C1)];
    };
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) s += load(i);
    const double mean = block_sum(s, sh) / n;
    double v = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const double dlt = load(i) - mean;
        v += dlt * dlt;
    }
    const double var = block_sum(v, sh) / n;
    const float rstd = float(1.0 / sqrt(var + double(eps)));
    const float fmean = float(mean);
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int p = i / Cg, c = g * Cg + i % Cg;
        float y = (load(i) - fmean) * rstd * gamma[c] + beta[c];
        if (silu) y = y / (1.0f + expf(-y));
        out[(static_cast<size_t>(b) * HW + p) * C + c] = y;
    }
}

// ------------------------------------------------------------------------------------------------ LayerNorm
__global__ void __launch_bounds__(256) ln32_kernel(const float* __restrict__ x, int rows, int C, float eps, const float* __restrict__ gamma,
                                                   const float* __restrict__ beta, float* __restrict__ out) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float* xr = x + static_cast<size_t>(row) * C;
    double s = 0.0;
    for (int c = lane; c < C; c += 32) s += xr[c];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const double mean = s / C;
    double v = 0.0;
    for (int c = lane; c < C; c += 32) { const double dlt = xr[c] - mean; v += dlt * dlt; }
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const float rstd = float(1.0 / sqrt(v / C + double(eps))), fmean = float(mean);
    for (int c = lane; c < C; c += 32) out[static_cast<size_t>(row) * C + c] = (xr[c] - fmean) * rstd * gamma[c] + beta[c];
}

// ------------------------------------------------------------------------------------------------ attention
// one warp per (batch, head, query row); keys in chunks of 32: lane = key for the scores (online softmax across
// chunks), lane = output dims {lane, lane+32, ...} for the P V accumulation
constexpr int A32_MAXD = 512;       // 512 = the VAE mid block's single head
__global__ void __launch_bounds__(128) attn32_kernel(const float* __restrict__ q, int ldq, const float* __restrict__ k, int ldk,
                                                     const float* __restrict__ v, int ldv, float* __restrict__ out, int ldo, int heads,
                                                     int d, int Tq, int Tk, float scale) {
    __shared__ float qs[4][A32_MAXD];
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row = blockIdx.x * 4 + wid;
    const int h = blockIdx.y, b = blockIdx.z;
    if (row >= Tq) return;
    const float* qr = q + (static_cast<size_t>(b) * Tq + row) * ldq + h * d;
    for (int c = lane; c < d; c += 32) qs[wid][c] = qr[c] * scale;
    __syncwarp();
    float o[A32_MAXD / 32];
#pragma unroll
    for (int i = 0; i < A32_MAXD / 32; ++i) o[i] = 0.f;
    float mrun = -INFINITY, l = 0.f;
    for (int j0 = 0; j0 < Tk; j0 += 32) {
        const int j = j0 + lane;
        float s = -INFINITY;
        if (j < Tk) {
            const float* kr = k + (static_cast<size_t>(b) * Tk + j) * ldk + h * d;
            float acc = 0.f;
            for (int c = 0; c < d; ++c) acc = fmaf(qs[wid][c], kr[c], acc);
            s = acc;
        }
        float mx = s;
        for (int o2 = 16; o2 > 0; o2 >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o2));
        const float mnew = fmaxf(mrun, mx);
        const float corr = expf(mrun - mnew);          // 0 on the first chunk (mrun = -inf)
        const float pj = j < Tk ? expf(s - mnew) : 0.f;
        float ps = pj;
        for (int o2 = 16; o2 > 0; o2 >>= 1) ps += __shfl_xor_sync(0xffffffffu, ps, o2);
        l = l * corr + ps;
        mrun = mnew;
#pragma unroll
        for (int i = 0; i < A32_MAXD / 32; ++i) o[i] *= corr;
        const int nk = min(32, Tk - j0);
        for (int jj = 0; jj < nk; ++jj) {
            const float pb = __shfl_sync(0xffffffffu, pj, jj);
            const float* vr = v + (static_cast<size_t>(b) * Tk + j0 + jj) * ldv + h * d;
#pragma unroll
            for (int i = 0; i < A32_MAXD / 32; ++i) {
                const int c = lane + 32 * i;
                if (c < d) o[i] = fmaf(pb, vr[c], o[i]);
            }
        }
    }
    float* orow = out + (static_cast<size_t>(b) * Tq + row) * ldo + h * d;
    const float inv = 1.0f / l;
#pragma unroll
    for (int i = 0; i < A32_MAXD / 32; ++i) {
        const int c = lane + 32 * i;
        if (c < d) orow[c] = o[i] * inv;
    }
}

// ------------------------------------------------------------------------------------------------ boundary layers
// conv_in: NCHW fp32 (sample | cond) -> NHWC fp32; thread per output element
__global__ void conv_in32_kernel(const float* __restrict__ xa, int Ca, const float* __restrict__ xb, int Cb, int B, int H, int W,
                                 const float* __restrict__ w, const float* __restrict__ bias, int Cout, float* __restrict__ out,
                                 const float* __restrict__ tap, float* __restrict__ out_post) {
    const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    const size_t total = static_cast<size_t>(B) * H * W * Cout;
    if (idx >= total) return;
    const int co = idx % Cout;
    const int ow = (idx / Cout) % W, oh = (idx / Cout / W) % H, b = idx / Cout / W / H;
    const int Cin = Ca + Cb;
    float acc = bias[co];
    for (int kh = 0; kh < 3; ++kh)
        for (int kw = 0; kw < 3; ++kw) {
            const int ih = oh + kh - 1, iw = ow + kw - 1;
            if (ih < 0 || ih >= H || iw < 0 || iw >= W) continue;
            for (int c = 0; c < Cin; ++c) {
                const float xv = c < Ca ? xa[((static_cast<size_t>(b) * Ca + c) * H + ih) * W + iw]
                                        : xb[((static_cast<size_t>(b) * Cb + (c - Ca)) * H + ih) * W + iw];
                acc = fmaf(xv, w[(static_cast<size_t>(kh * 3 + kw) * Cin + c) * Cout + co], acc);
            }
        }
    out[idx] = acc;
    if (tap) out_post[idx] = acc + tap[idx];
}

// conv_out: NHWC fp32 -> NCHW fp32, w [Cout][3][3][Cin]; one warp per output element (lanes over input channels)
__global__ void __launch_bounds__(256) conv_out32_kernel(const float* __restrict__ x, int Cin, int B, int H, int W,
                                                         const float* __restrict__ w, const float* __restrict__ bias, int Cout,
                                                         float* __restrict__ out) {
    const size_t wi = (static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const size_t total = static_cast<size_t>(B) * Cout * H * W;
    if (wi >= total) return;
    const int ow = wi % W, oh = (wi / W) % H, co = (wi / W / H) % Cout, b = wi / W / H / Cout;
    float acc = 0.f;
    for (int kh = 0; kh < 3; ++kh)
        for (int kw = 0; kw < 3; ++kw) {
            const int ih = oh + kh - 1, iw = ow + kw - 1;
            if (ih < 0 || ih >= H || iw < 0 || iw >= W) continue;
            const float* xr = x + ((static_cast<size_t>(b) * H + ih) * W + iw) * Cin;
            const float* wr = w + (static_cast<size_t>(co) * 9 + kh * 3 + kw) * Cin;
            for (int c = lane; c < Cin; c += 32) acc = fmaf(xr[c], wr[c], acc);
        }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) out[wi] = acc + bias[co];
}

// y[M,N] = act_out(W[N,K] . act_in(x[M,K]) + b[N]), fp32 weights; one warp per output column
__global__ void __launch_bounds__(256) linear_small32_kernel(const float* __restrict__ x, int M, int K, const float* __restrict__ w,
                                                             const float* __restrict__ bias, int N, int act_in, int act_out,
                                                             float* __restrict__ y) {
    const int n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (n >= N) return;
    const float* wr = w + static_cast<size_t>(n) * K;
    for (int m = 0; m < M; ++m) {
        float acc = 0.f;
        for (int kk = lane; kk < K; kk += 32) {
            float xv = x[static_cast<size_t>(m) * K + kk];
            if (act_in) xv = xv / (1.0f + expf(-xv));
            acc = fmaf(xv, wr[kk], acc);
        }
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) {
            float r = acc + (bias ? bias[n] : 0.f);
            if (act_out) r = r / (1.0f + expf(-r));
            y[static_cast<size_t>(m) * N + n] = r;
        }
    }
}

}  // namespace mfb

using namespace mfb;

extern "C" int mfb_groupnorm_f32(const float* x1, int C1, const float* x2, int C2, int B, int HW, int groups, float eps,
                                 const float* gamma, const float* beta, int silu, float* out, void* stream) {
    MFB_RECORD(mfb_groupnorm_f32(x1, C1, x2, C2, B, HW, groups, eps, gamma, beta, silu, out, stream));
    MFB_REQUIRE(x1 && gamma && beta && out, "null pointer");
    MFB_REQUIRE((C1 + C2) % groups == 0 && (x2 != nullptr) == (C2 > 0), "bad channel split");
    gn32_kernel<<<dim3(groups, B), 256, 0, static_cast<cudaStream_t>(stream)>>>(x1, C1, x2, C2, HW, groups, eps, gamma, beta, silu, out);
    MFB_CUDA_OK(cudaGetLastError());
    return MFB_OK;
}

extern "C" int mfb_layernorm_f32(const float* x, int rows, int C, float eps, const float* gamma, const float* beta, float* out,
                                 void* stream) {
    MFB_RECORD(mfb_layernorm_f32(x, rows, C, eps, gamma, beta, out, stream));
    MFB_REQUIRE(x && gamma && beta && out, "null pointer");
    ln32_kernel<<<(rows + 7) / 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, rows, C, eps, gamma, beta, out);
    MFB_CUDA_OK(cudaGetLastError());
    return MFB_OK;
}

extern "C" int mfb_attention_f32(const float* q, int ldq, const float* k, int ldk, const float* v, int ldv, float* out, int ldo,
                                 int B, int heads, int head_dim, int Tq, int Tk, void* stream) {
    MFB_RECORD(mfb_attention_f32(q, ldq, k, ldk, v, ldv, out, ldo, B, heads, head_dim, Tq, Tk, stream));
    MFB_REQUIRE(q && k && v && out, "null pointer");
    MFB_REQUIRE(head_dim > 0 && head_dim <= A32_MAXD && Tq > 0 && Tk > 0, "unsupported attention shape");
    attn32_kernel<<<dim3((Tq + 3) / 4, heads, B), 128, 0, static_cast<cudaStream_t>(stream)>>>(
        q, ldq, k, ldk, v, ldv, out, ldo, heads, head_dim, Tq, Tk, 1.0f / sqrtf(static_cast<float>(head_dim)));
    MFB_CUDA_OK(cudaGetLastError());
    return MFB_OK;
}

extern "C" int mfb_conv_in_f32(const float* sample, int Ca, const float* cond, int Cb, int B, int H, int W, const float* w,
                               const float* bias, int Cout, float* out, const float* tap, float* out_post, void* stream) {
    MFB_RECORD(mfb_conv_in_f32(sample, Ca, cond, Cb, B, H, W, w, bias, Cout, out, tap, out_post, stream));
    MFB_REQUIRE(sample && w && bias && out && (cond != nullptr) == (Cb > 0) && (tap == nullptr || out_post != nullptr), "bad arguments");
    const size_t total = static_cast<size_t>(B) * H * W * Cout;
    conv_in32_kernel<<<unsigned((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(sample, Ca, cond, Cb, B, H, W, w, bias,
                                                                                                  Cout, out, tap, out_post);
    MFB_CUDA_OK(cudaGetLastError());
    return MFB_OK;
}

extern "C" int mfb_conv_out_f32(const float* x, int Cin, int B, int H, int W, const float* w, const float* bias, int Cout, float* out,
                                void* stream) {
    MFB_RECORD(mfb_conv_out_f32(x, Cin, B, H, W, w, bias, Cout, out, stream));
    MFB_REQUIRE(x && w && bias && out, "null pointer");
    const size_t warps = static_cast<size_t>(B) * Cout * H * W;
    conv_out32_kernel<<<unsigned((warps * 32 + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, Cin, B, H, W, w, bias, Cout, out);
    MFB_CUDA_OK(cudaGetLastError());
    return MFB_OK;
}

extern "C" int mfb_linear_small_f32(const float* x, int M, int K, const float* w, const float* b, int N, int act_in, int act_out,
                                    float* y, void* stream) {
    MFB_RECORD(mfb_linear_small_f32(x, M, K, w, b, N, act_in, act_out, y, stream));
    MFB_REQUIRE(x && w && y, "null pointer");
    linear_small32_kernel<<<(N * 32 + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, M, K, w, b, N, act_in, act_out, y);
    MFB_CUDA_OK(cudaGetLastError());
    return MFB_OK;
}
