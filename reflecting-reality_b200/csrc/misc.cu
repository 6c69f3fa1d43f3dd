// Boundary and small layers of the denoise step: conv_in / conv_out (4..10 and 4 channels: CUDA-core direct
// convolutions, <0.1 % of the FLOPs), nearest upsample, layout conversion, the timestep path and the fused
// CFG + scheduler update.  All are bandwidth/latency bound; accesses are coalesced and vectorised where the
// shapes allow.
#include "common.h"
#include "ptx.cuh"

namespace mfb {

// ---------------------------------------------------------------------------------------------- conv_in
// grid (ceil(W/8), ceil(H/8), B), block = 8 * Cout/8 threads (320 for Cout = 320).  Input patch (10x10xCin fp32)
// staged in smem; a thread owns an 8-channel group and one 8-pixel row of the tile, so each weight vector
// ([3][3][Cin][Cout] fp32, coalesced over the channel groups of consecutive threads) is loaded once per 8 pixels.
__global__ void conv_in_kernel(const float* __restrict__ xa, int Ca, const float* __restrict__ xb, int Cb, int H, int W,
                               const float* __restrict__ w, const float* __restrict__ bias, int Cout,
                               __nv_bfloat16* __restrict__ out, const __nv_bfloat16* __restrict__ tap,
                               __nv_bfloat16* __restrict__ out_post) {
    extern __shared__ float patch[];  // [Cin][10][10]
    const int Cin = Ca + Cb;
    const int b = blockIdx.z;
    const int h0 = blockIdx.y * 8, w0 = blockIdx.x * 8;
    pdl_trigger();
    pdl_wait();
    for (int i = threadIdx.x; i < Cin * 100; i += blockDim.x) {
        const int c = i / 100, r = i % 100;
        const int hh = h0 + r / 10 - 1, ww = w0 + r % 10 - 1;
        float v = 0.f;
        if (hh >= 0 && hh < H && ww >= 0 && ww < W) {
            v = c < Ca ? xa[((static_cast<size_t>(b) * Ca + c) * H + hh) * W + ww]
                       : xb[((static_cast<size_t>(b) * Cb + (c - Ca)) * H + hh) * W + ww];
        }
        patch[i] = v;
    }
    __syncthreads();
    const int ngroups = Cout / 8;
    for (int item = threadIdx.x; item < 8 * ngroups; item += blockDim.x) {
        const int cg = item % ngroups, ph = item / ngroups;
        const int oh = h0 + ph;
        if (oh >= H) continue;
        float acc[8][8];  // [pixel in row][channel]
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const float bv = __ldg(&bias[cg * 8 + e]);
#pragma unroll
            for (int px = 0; px < 8; ++px) acc[px][e] = bv;
        }
        for (int kh = 0; kh < 3; ++kh)
            for (int kw = 0; kw < 3; ++kw)
                for (int c = 0; c < Cin; ++c) {
                    const float* wp = w + ((static_cast<size_t>(kh * 3 + kw) * Cin + c) * Cout + cg * 8);
                    const float4 w0v = __ldg(reinterpret_cast<const float4*>(wp));
                    const float4 w1v = __ldg(reinterpret_cast<const float4*>(wp + 4));
                    const float* prow = patch + c * 100 + (ph + kh) * 10 + kw;
#pragma unroll
                    for (int px = 0; px < 8; ++px) {
                        const float xv = prow[px];
                        acc[px][0] = fmaf(xv, w0v.x, acc[px][0]); acc[px][1] = fmaf(xv, w0v.y, acc[px][1]);
                        acc[px][2] = fmaf(xv, w0v.z, acc[px][2]); acc[px][3] = fmaf(xv, w0v.w, acc[px][3]);
                        acc[px][4] = fmaf(xv, w1v.x, acc[px][4]); acc[px][5] = fmaf(xv, w1v.y, acc[px][5]);
                        acc[px][6] = fmaf(xv, w1v.z, acc[px][6]); acc[px][7] = fmaf(xv, w1v.w, acc[px][7]);
                    }
                }
#pragma unroll
        for (int px = 0; px < 8; ++px) {
            const int ow = w0 + px;
            if (ow >= W) continue;
            const size_t o = ((static_cast<size_t>(b) * H + oh) * W + ow) * Cout + cg * 8;
            uint4 pk;
            pk.x = pack_bf16x2(acc[px][0], acc[px][1]); pk.y = pack_bf16x2(acc[px][2], acc[px][3]);
            pk.z = pack_bf16x2(acc[px][4], acc[px][5]); pk.w = pack_bf16x2(acc[px][6], acc[px][7]);
            *reinterpret_cast<uint4*>(out + o) = pk;
            if (tap) {
                const uint4 tv = __ldg(reinterpret_cast<const uint4*>(tap + o));
                float2 t;
                float f[8];
                t = unpack_bf16x2(tv.x); f[0] = acc[px][0] + t.x; f[1] = acc[px][1] + t.y;
                t = unpack_bf16x2(tv.y); f[2] = acc[px][2] + t.x; f[3] = acc[px][3] + t.y;
                t = unpack_bf16x2(tv.z); f[4] = acc[px][4] + t.x; f[5] = acc[px][5] + t.y;
                t = unpack_bf16x2(tv.w); f[6] = acc[px][6] + t.x; f[7] = acc[px][7] + t.y;
                pk.x = pack_bf16x2(f[0], f[1]); pk.y = pack_bf16x2(f[2], f[3]);
                pk.z = pack_bf16x2(f[4], f[5]); pk.w = pack_bf16x2(f[6], f[7]);
                *reinterpret_cast<uint4*>(out_post + o) = pk;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------- conv_out
// 320 -> 4 channel 3x3 conv at the fp32 NCHW boundary, CUDA cores (fp32 accumulate AND fp32 output: rounding the
// noise prediction to bf16 would eat the parity margin).  Persistent CTAs of 256 threads = 32 pixel quads x 8 channel
// slices.  A thread owns 4 consecutive pixels of a row and the 16-byte channel vectors v = 8k + s of its slice s (the 8
// lanes of a quad read 128 contiguous bytes per pixel); per (k, filter row) it loads the 6 input pixels the quad needs
// once and reuses every weight float4 (4 outputs per input channel) for all 4 pixels — a weight LDS.128 costs four LSU
// cycles per warp, so without that reuse the kernel is bound by shared-memory reads (131 us), not by the 755 MFMA.
// Weights are staged ONCE per CTA as [tap][k][e][s] float4, so the 8 slices of a warp read 128 contiguous bytes
// (conflict-free) and the 4 quads of the warp share them by broadcast.  Partial sums of the 8 slices are combined with
// a fixed-order shuffle butterfly.
__global__ void __launch_bounds__(256) conv_out_kernel(const __nv_bfloat16* __restrict__ x, int Cin, int B, int H, int W,
                                                        const float* __restrict__ w, const float* __restrict__ bias, int Cout,
                                                        float* __restrict__ out) {
    extern __shared__ float4 ws4[];  // [9][KV][8][8] : tap, k, element e, slice s  ->  channel c = (8k + s) * 8 + e
    const int K = 9 * Cin;
    const int nvec = Cin / 8;                 // 16-byte vectors per pixel
    const int KV = (nvec + 7) / 8;            // vectors per slice
    for (int i = threadIdx.x; i < 9 * KV * 64; i += blockDim.x) {
        const int s = i & 7, e = (i >> 3) & 7, k = (i >> 6) % KV, tap = i / (64 * KV);
        const int v = 8 * k + s;
        float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
        if (v < nvec) {
            const int src = tap * Cin + v * 8 + e;
            val.x = w[src];
            if (Cout > 1) val.y = w[K + src];
            if (Cout > 2) val.z = w[2 * K + src];
            if (Cout > 3) val.w = w[3 * K + src];
        }
        ws4[i] = val;
    }
    __syncthreads();
    pdl_trigger();
    pdl_wait();      // the weights above are static; the activations below come from the previous kernel
    const int s = threadIdx.x & 7;
    const int qpr = (W + 3) / 4;                                   // pixel quads per row
    const long long nquad = static_cast<long long>(B) * H * qpr;
    const size_t hw = static_cast<size_t>(H) * W;
    for (long long q = static_cast<long long>(blockIdx.x) * 32 + (threadIdx.x >> 3); q < ((nquad + 31) / 32) * 32;
         q += static_cast<long long>(gridDim.x) * 32) {
        const bool live = q < nquad;           // whole 8-lane groups are live or not; dead ones still join the shuffles
        const int ow0 = live ? int(q % qpr) * 4 : 0;
        const int oh = live ? int((q / qpr) % H) : 0;
        const int b = live ? int(q / (static_cast<long long>(qpr) * H)) : 0;
        float4 acc[4];
#pragma unroll
        for (int px = 0; px < 4; ++px) acc[px] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (live) {
#pragma unroll 1
            for (int k = 0; k < KV; ++k) {
                const int v = 8 * k + s;
                if (v >= nvec) continue;
#pragma unroll 1
                for (int kh = 0; kh < 3; ++kh) {
                    const int hh = oh + kh - 1;
                    if (hh < 0 || hh >= H) continue;
                    const uint4* row = reinterpret_cast<const uint4*>(x + ((static_cast<size_t>(b) * H + hh) * W) * Cin) + v;
                    float f[6][8];
#pragma unroll
                    for (int cc = 0; cc < 6; ++cc) {
                        const int col = ow0 - 1 + cc;
                        uint4 u = make_uint4(0, 0, 0, 0);
                        if (col >= 0 && col < W) u = __ldg(row + static_cast<size_t>(col) * nvec);
                        float2 t;
                        t = unpack_bf16x2(u.x); f[cc][0] = t.x; f[cc][1] = t.y;
                        t = unpack_bf16x2(u.y); f[cc][2] = t.x; f[cc][3] = t.y;
                        t = unpack_bf16x2(u.z); f[cc][4] = t.x; f[cc][5] = t.y;
                        t = unpack_bf16x2(u.w); f[cc][6] = t.x; f[cc][7] = t.y;
                    }
#pragma unroll
                    for (int kw = 0; kw < 3; ++kw) {
                        const float4* wt = ws4 + ((kh * 3 + kw) * KV + k) * 64 + s;
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            const float4 wv = wt[e * 8];
#pragma unroll
                            for (int px = 0; px < 4; ++px) {
                                const float xv = f[px + kw][e];
                                acc[px].x = fmaf(xv, wv.x, acc[px].x);
                                acc[px].y = fmaf(xv, wv.y, acc[px].y);
                                acc[px].z = fmaf(xv, wv.z, acc[px].z);
                                acc[px].w = fmaf(xv, wv.w, acc[px].w);
                            }
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int px = 0; px < 4; ++px) {
#pragma unroll
            for (int d = 1; d < 8; d <<= 1) {      // fixed-order butterfly over the 8 channel slices
                acc[px].x += __shfl_xor_sync(0xffffffffu, acc[px].x, d);
                acc[px].y += __shfl_xor_sync(0xffffffffu, acc[px].y, d);
                acc[px].z += __shfl_xor_sync(0xffffffffu, acc[px].z, d);
                acc[px].w += __shfl_xor_sync(0xffffffffu, acc[px].w, d);
            }
        }
        if (live && s < Cout) {
#pragma unroll
            for (int px = 0; px < 4; ++px) {
                if (ow0 + px < W) {
                    const float r = s == 0 ? acc[px].x : s == 1 ? acc[px].y : s == 2 ? acc[px].z : acc[px].w;
                    out[static_cast<size_t>(b) * Cout * hw + s * hw + static_cast<size_t>(oh) * W + ow0 + px] = r + bias[s];
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------- upsample / layout
__global__ void upsample2x_kernel(const uint4* __restrict__ x, int H, int W, int CV, uint4* __restrict__ out, long long total) {
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int cv = i % CV;
        long long r = i / CV;
        const int ow = r % (2 * W); r /= (2 * W);
        const int oh = r % (2 * H);
        const long long b = r / (2 * H);
        out[i] = __ldg(&x[((b * H + (oh >> 1)) * W + (ow >> 1)) * CV + cv]);
    }
}

// NCHW fp32 -> NHWC bf16 via a 32x32 smem transpose over (C, HW)
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ x, int C, int HW, __nv_bfloat16* __restrict__ out) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z;
    const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int c = c0 + i, p = p0 + threadIdx.x;
        tile[i][threadIdx.x] = (c < C && p < HW) ? x[(static_cast<size_t>(b) * C + c) * HW + p] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int p = p0 + i, c = c0 + threadIdx.x;
        if (c < C && p < HW) out[(static_cast<size_t>(b) * HW + p) * C + c] = __float2bfloat16(tile[threadIdx.x][i]);
    }
}
__global__ void nhwc_to_nchw_kernel(const __nv_bfloat16* __restrict__ x, int C, int HW, float* __restrict__ out) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z;
    const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int p = p0 + i, c = c0 + threadIdx.x;
        tile[i][threadIdx.x] = (c < C && p < HW) ? __bfloat162float(x[(static_cast<size_t>(b) * HW + p) * C + c]) : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int c = c0 + i, p = p0 + threadIdx.x;
        if (c < C && p < HW) out[(static_cast<size_t>(b) * C + c) * HW + p] = tile[threadIdx.x][i];
    }
}
__global__ void f32_to_bf16_kernel(const float* __restrict__ x, long long n, __nv_bfloat16* __restrict__ out) {
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x)
        out[i] = __float2bfloat16(x[i]);
}

// [B, T, ld] columns [col0, col0+C) -> [B, C, ldt] (token index contiguous), zero padding for t in [T, ldt).
// 64 x 64 tile per CTA, 16-byte global accesses on both sides (a row of 64 channels / 64 tokens = one 128 B line):
// thread (tl, cv) loads 8 channels of tokens tl, tl+32, scatters them into a [channel][token] smem tile (pitch 72
// elements = 144 B keeps the 16-byte rows of the read side aligned), then thread (cl, tv) stores 8 tokens of channels
// cl, cl+32.  Requires col0 % 8 == 0, C % 8 == 0, ld % 8 == 0, ldt % 8 == 0 (checked on the host).
__global__ void __launch_bounds__(256) transpose_tokens_kernel(const __nv_bfloat16* __restrict__ x, int ld, int col0, int C, int T,
                                                                __nv_bfloat16* __restrict__ out, int ldt) {
    __shared__ __align__(16) __nv_bfloat16 tile[64][72];
    const int b = blockIdx.z;
    const int t0 = blockIdx.x * 64, c0 = blockIdx.y * 64;
    const int v = threadIdx.x & 7, l = threadIdx.x >> 3;     // 8 vectors x 32 rows
    pdl_trigger();
    pdl_wait();
    uint4 u[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const int t = t0 + l + 32 * k, c = c0 + v * 8;
        u[k] = (t < T && c < C) ? __ldg(reinterpret_cast<const uint4*>(x + (static_cast<size_t>(b) * T + t) * ld + col0 + c))
                                : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const __nv_bfloat16* e = reinterpret_cast<const __nv_bfloat16*>(&u[k]);
#pragma unroll
        for (int i = 0; i < 8; ++i) tile[v * 8 + i][l + 32 * k] = e[i];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const int c = c0 + l + 32 * k, t = t0 + v * 8;
        if (c < C && t < ldt)
            *reinterpret_cast<uint4*>(out + (static_cast<size_t>(b) * C + c) * ldt + t) = *reinterpret_cast<const uint4*>(&tile[l + 32 * k][v * 8]);
    }
}

// ---------------------------------------------------------------------------------------------- timestep path
__global__ void sinusoid_kernel(const float* __restrict__ t, int M, int dim, float* __restrict__ out) {
    const int half = dim / 2;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M * half) return;
    const int m = i / half, j = i % half;
    // exponent = -ln(10000) * j / half   (embeddings.py:45-49, downscale_freq_shift = 0)
    const float freq = expf(-9.210340371976184f * static_cast<float>(j) / static_cast<float>(half));
    const float a = t[m] * freq;
    out[static_cast<size_t>(m) * dim + j] = cosf(a);          // flip_sin_to_cos: cos first
    out[static_cast<size_t>(m) * dim + half + j] = sinf(a);
}

// y[m, n] = act_out(b[n] + sum_k W[n,k] * act_in(x[m,k])).  grid (ceil(N/64), ceil(M/8)), block 256:
// the 8 activation rows of the block are staged once in shared memory (with act_in applied) and reused by
// the 64 output columns of the block (8 warps x 8 columns); weights stream through as 16-byte bf16 vectors.
__global__ void linear_small_kernel(const float* __restrict__ x, int M, int K, const __nv_bfloat16* __restrict__ w,
                                    const float* __restrict__ b, int N, int act_in, int act_out, float* __restrict__ y) {
    extern __shared__ float xs[];  // [8][K]
    const int m0 = blockIdx.y * 8;
    for (int i = threadIdx.x; i < 8 * K; i += blockDim.x) {
        const int r = i / K, k = i % K;
        float v = (m0 + r < M) ? x[static_cast<size_t>(m0 + r) * K + k] : 0.f;
        if (act_in) v = v / (1.f + expf(-v));
        xs[i] = v;
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int j = 0; j < 8; ++j) {
        const int n = blockIdx.x * 64 + warp * 8 + j;
        if (n >= N) break;
        float acc[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) acc[r] = 0.f;
        const __nv_bfloat16* wr = w + static_cast<size_t>(n) * K;
        for (int k = lane * 8; k < K; k += 256) {
            const uint4 u = __ldg(reinterpret_cast<const uint4*>(wr + k));
            float wv[8];
            float2 t;
            t = unpack_bf16x2(u.x); wv[0] = t.x; wv[1] = t.y;
            t = unpack_bf16x2(u.y); wv[2] = t.x; wv[3] = t.y;
            t = unpack_bf16x2(u.z); wv[4] = t.x; wv[5] = t.y;
            t = unpack_bf16x2(u.w); wv[6] = t.x; wv[7] = t.y;
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const float4 x0 = *reinterpret_cast<const float4*>(xs + r * K + k);
                const float4 x1 = *reinterpret_cast<const float4*>(xs + r * K + k + 4);
                acc[r] += wv[0] * x0.x + wv[1] * x0.y + wv[2] * x0.z + wv[3] * x0.w + wv[4] * x1.x + wv[5] * x1.y +
                          wv[6] * x1.z + wv[7] * x1.w;
            }
        }
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            float a = acc[r];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
            if (lane == 0 && m0 + r < M) {
                a += b ? b[n] : 0.f;
                if (act_out) a = a / (1.f + expf(-a));
                y[static_cast<size_t>(m0 + r) * N + n] = a;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------- CFG + scheduler
__global__ void cfg_sched_kernel(const float* __restrict__ eps_u, const float* __restrict__ eps_c, float* __restrict__ x, float* __restrict__ last,
                                 float* __restrict__ m0, float* __restrict__ m1, const float* __restrict__ coef,
                                 long long total /* Bimg*n */) {
    pdl_trigger();
    pdl_wait();
    const float g = coef[0], c_x = coef[1], c_eps = coef[2];
    const float a_last = coef[3], a_m0 = coef[4], a_m1 = coef[5], a_mt = coef[6], use_corr = coef[7];
    const float b_x = coef[8], b_mt = coef[9], b_m0 = coef[10], b_eps = coef[11];
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const float eu = eps_u[i], ec = eps_c[i];
        const float e = eu + g * (ec - eu);                    // pipeline_brushnet.py:1311-1312
        const float xv = x[i];
        const float mt = c_x * xv + c_eps * e;                 // convert_model_output (x0 prediction)
        const float m0v = m0[i], m1v = m1[i];
        float xc = xv;
        if (use_corr != 0.f) xc = a_last * last[i] + a_m0 * m0v + a_m1 * m1v + a_mt * mt;   // UniC
        const float xn = b_x * xc + b_mt * mt + b_m0 * m0v + b_eps * e;                      // UniP / DDIM
        x[i] = xn;
        last[i] = xc;
        m1[i] = m0v;
        m0[i] = mt;
    }
}

}  // namespace mfb

using namespace mfb;

static inline int grid_for(long long n, int block, int cap = 148 * 8) {
    long long g = (n + block - 1) / block;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return static_cast<int>(g);
}

extern "C" int mfb_conv_in(const float* sample, int Ca, const float* cond, int Cb, int B, int H, int W, const float* w,
                           const float* bias, int Cout, void* out, const void* tap, void* out_post, void* stream) {
    MFB_RECORD(mfb_conv_in(sample, Ca, cond, Cb, B, H, W, w, bias, Cout, out, tap, out_post, stream));
    MFB_REQUIRE(sample && w && bias && out, "null pointer");
    if (!cond) Cb = 0;
    MFB_REQUIRE(Cout % 8 == 0, "Cout must be a multiple of 8");
    MFB_REQUIRE(!tap || out_post, "tap given without out_post");
    const int Cin = Ca + Cb;
    MFB_REQUIRE(Cin <= 64, "conv_in supports at most 64 input channels");
    dim3 grid((W + 7) / 8, (H + 7) / 8, B);
    int threads = 8 * (Cout / 8);
    if (threads > 640) threads = 640;
    threads = (threads + 31) / 32 * 32;
    MFB_CUDA_OK(launch_k(conv_in_kernel, grid, dim3(threads), Cin * 100 * sizeof(float), static_cast<cudaStream_t>(stream), 1,
                         sample, Ca, cond, Cb, H, W, w, bias, Cout, static_cast<__nv_bfloat16*>(out),
                         static_cast<const __nv_bfloat16*>(tap), static_cast<__nv_bfloat16*>(out_post)));
    return MFB_OK;
}

extern "C" int mfb_conv_out(const void* x, int Cin, int B, int H, int W, const float* w, const float* bias, int Cout,
                            float* out, void* stream) {
    MFB_RECORD(mfb_conv_out(x, Cin, B, H, W, w, bias, Cout, out, stream));
    MFB_REQUIRE(x && w && bias && out, "null pointer");
    MFB_REQUIRE(Cout >= 1 && Cout <= 4 && Cin % 8 == 0, "conv_out supports Cout <= 4, Cin %% 8 == 0");
    const size_t smem = static_cast<size_t>(9) * ((Cin / 8 + 7) / 8) * 64 * sizeof(float4);
    MFB_REQUIRE(smem <= 200 * 1024, "conv_out weights do not fit shared memory");
    static bool configured = false;
    if (!configured) {
        MFB_CUDA_OK(cudaFuncSetAttribute(conv_out_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        configured = true;
    }
    const long long nquad = static_cast<long long>(B) * H * ((W + 3) / 4);
    long long grid = (nquad + 31) / 32;
    const long long resident = 2LL * device_sm_count();      // persistent: the weight staging is paid once per CTA
    if (grid > resident) grid = resident;
    MFB_CUDA_OK(launch_k(conv_out_kernel, dim3(static_cast<unsigned>(grid)), dim3(256), smem, static_cast<cudaStream_t>(stream), 1,
                         static_cast<const __nv_bfloat16*>(x), Cin, B, H, W, w, bias, Cout, out));
    return MFB_OK;
}

extern "C" int mfb_upsample2x(const void* x, int B, int H, int W, int C, void* out, void* stream) {
    MFB_RECORD(mfb_upsample2x(x, B, H, W, C, out, stream));
    MFB_REQUIRE(x && out && C % 8 == 0, "bad arguments");
    const long long total = static_cast<long long>(B) * 4 * H * W * (C / 8);
    upsample2x_kernel<<<grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const uint4*>(x), H, W, C / 8, static_cast<uint4*>(out), total);
    MFB_CUDA_OK(cudaGetLastError());
    return MFB_OK;
}

extern "C" int mfb_nchw_f32_to_nhwc_bf16(const float* x, int B, int C, int H, int W, void* out, void* stream) {
    MFB_RECORD(mfb_nchw_f32_to_nhwc_bf16(x, B, C, H, W, out, stream));
    MFB_REQUIRE(x && out, "null pointer");
    dim3 grid((H * W + 31) / 32, (C + 31) / 32, B), block(32, 8);
    nchw_to_nhwc_kernel<<<grid, block, 0, static_cast<cudaStream_t>(stream)>>>(x, C, H * W, static_cast<__nv_bfloat16*>(out));
    MFB_CUDA_OK(cudaGetLastError());
    return MFB_OK;
}

extern "C" int mfb_nhwc_bf16_to_nchw_f32(const void* x, int B, int C, int H, int W, float* out, void* stream) {
    MFB_RECORD(mfb_nhwc_bf16_to_nchw_f32(x, B, C, H, W, out, stream));
    MFB_REQUIRE(x && out, "null pointer");
    dim3 grid((H * W + 31) / 32, (C + 31) / 32, B), block(32, 8);
    nhwc_to_nchw_kernel<<<grid, block, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const __nv_bfloat16*>(x), C, H * W, out);
    MFB_CUDA_OK(cudaGetLastError());
    return MFB_OK;
}

extern "C" int mfb_f32_to_bf16(const float* x, long long n, void* out, void* stream) {
    MFB_RECORD(mfb_f32_to_bf16(x, n, out, stream));
    MFB_REQUIRE(x && out, "null pointer");
    f32_to_bf16_kernel<<<grid_for(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, n, static_cast<__nv_bfloat16*>(out));
    MFB_CUDA_OK(cudaGetLastError());
    return MFB_OK;
}

extern "C" int mfb_transpose_tokens(const void* x, int ld, int col0, int C, int B, int T, void* out, int ldt, void* stream) {
    MFB_RECORD(mfb_transpose_tokens(x, ld, col0, C, B, T, out, ldt, stream));
    MFB_REQUIRE(x && out && ldt >= T, "bad arguments");
    MFB_REQUIRE(ld % 8 == 0 && col0 % 8 == 0 && C % 8 == 0 && ldt % 8 == 0, "transpose_tokens needs ld, col0, C, ldt to be multiples of 8");
    dim3 grid((ldt + 63) / 64, (C + 63) / 64, B), block(256);
    MFB_CUDA_OK(launch_k(transpose_tokens_kernel, grid, block, 0, static_cast<cudaStream_t>(stream), 1,
                         static_cast<const __nv_bfloat16*>(x), ld, col0, C, T, static_cast<__nv_bfloat16*>(out), ldt));
    return MFB_OK;
}

extern "C" int mfb_timestep_sinusoid(const float* t, int M, int dim, float* out, void* stream) {
    MFB_RECORD(mfb_timestep_sinusoid(t, M, dim, out, stream));
    MFB_REQUIRE(t && out && dim % 2 == 0, "bad arguments");
    const int n = M * dim / 2;
    sinusoid_kernel<<<(n + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(t, M, dim, out);
    MFB_CUDA_OK(cudaGetLastError());
    return MFB_OK;
}

extern "C" int mfb_linear_small(const float* x, int M, int K, const void* w, const float* b, int N, int act_in, int act_out,
                                float* y, void* stream) {
    MFB_RECORD(mfb_linear_small(x, M, K, w, b, N, act_in, act_out, y, stream));
    MFB_REQUIRE(x && w && y && K % 8 == 0 && K <= 4096, "linear_small needs K %% 8 == 0 and K <= 4096");
    const size_t smem = static_cast<size_t>(8) * K * sizeof(float);
    static bool configured = false;
    if (!configured) {
        MFB_CUDA_OK(cudaFuncSetAttribute(linear_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 4096 * 4));
        configured = true;
    }
    dim3 grid((N + 63) / 64, (M + 7) / 8), block(256);
    linear_small_kernel<<<grid, block, smem, static_cast<cudaStream_t>(stream)>>>(x, M, K, static_cast<const __nv_bfloat16*>(w), b, N,
                                                                             act_in, act_out, y);
    MFB_CUDA_OK(cudaGetLastError());
    return MFB_OK;
}

// ---------------------------------------------------------------------------------------------- SynMirror input preprocessing
// The eval sweep's per-image host work of the reference (PIL + numpy + torchvision, E/test_brushnet.py:186-226,
// S/pipelines/brushnet/pipeline_brushnet.py:741-774,1116-1202, E/dataset/dataset.py:98-145) as three small kernels on
// uint8 / fp32 arrays that are already at the target resolution.
namespace mfb {
// uint8 HWC RGB -> fp32 NCHW in [-1, 1]  (VaeImageProcessor.preprocess: /255 then 2x - 1)
__global__ void prep_image_kernel(const uint8_t* __restrict__ in, int H, int W, float* __restrict__ out, long long npix) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;      // pixel index over N*H*W
    if (i >= npix) return;
    const long long n = i / (static_cast<long long>(H) * W), p = i % (static_cast<long long>(H) * W);
#pragma unroll
    for (int c = 0; c < 3; ++c) out[(n * 3 + c) * H * W + p] = 2.0f * (static_cast<float>(in[i * 3 + c]) / 255.0f) - 1.0f;
}
// per image: max of depth over the pixels with mask > 0 (non-negative depths: the int ordering of the bit patterns is the float one)
__global__ void depth_max_kernel(const float* __restrict__ depth, const uint8_t* __restrict__ mask, long long hw, int* __restrict__ maxbits) {
    const int n = blockIdx.y;
    float m = 0.f;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < hw; i += static_cast<long long>(gridDim.x) * blockDim.x)
        if (mask[n * hw + i] > 0) m = fmaxf(m, depth[n * hw + i]);
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(maxbits + n, __float_as_int(m));
}
// latent-resolution mask and depth: nearest sampling at (f*i, f*j) (F.interpolate default mode, pipeline_brushnet.py:1190-1202);
// mask_lat = 1 where the (3-channel, [-1,1]-normalised) mask sums below 0, i.e. where the uint8 mask is < 127.5 (:1139);
// depth_lat = 2 * clip(d, 0, dmax) / dmax - 1 with dmax = max depth over the mask + delta (dataset.py:131-145)
__global__ void prep_mask_depth_kernel(const uint8_t* __restrict__ mask, const float* __restrict__ depth, int H, int W, int f,
                                       const int* __restrict__ maxbits, float delta, float* __restrict__ mask_lat,
                                       float* __restrict__ depth_lat, long long nlat) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= nlat) return;
    const int h = H / f, w = W / f;
    const long long n = i / (h * w);
    const int y = (i / w) % h, x = i % w;
    const long long src = (n * H + static_cast<long long>(y) * f) * W + static_cast<long long>(x) * f;
    mask_lat[i] = mask[src] < 128 ? 1.0f : 0.0f;
    if (depth) {
        const float dmax = __int_as_float(maxbits[n]) + delta;
        depth_lat[i] = 2.0f * (fminf(fmaxf(depth[src], 0.0f), dmax) / dmax) - 1.0f;
    }
}
// fp32 NCHW in [-1, 1] -> uint8 HWC  (VaeImageProcessor.postprocess: (x / 2 + 0.5).clamp(0, 1), then (255 x).round())
__global__ void post_image_kernel(const float* __restrict__ in, int H, int W, uint8_t* __restrict__ out, long long npix) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= npix) return;
    const long long n = i / (static_cast<long long>(H) * W), p = i % (static_cast<long long>(H) * W);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float v = fminf(fmaxf(in[(n * 3 + c) * H * W + p] * 0.5f + 0.5f, 0.0f), 1.0f);
        out[i * 3 + c] = static_cast<uint8_t>(rintf(v * 255.0f));
    }
}
}  // namespace mfb

extern "C" int mfb_prep_image_u8(const void* rgb_hwc, int N, int H, int W, float* out_nchw, void* stream) {
    MFB_REQUIRE(rgb_hwc && out_nchw && N > 0 && H > 0 && W > 0, "bad arguments");
    const long long npix = static_cast<long long>(N) * H * W;
    prep_image_kernel<<<unsigned((npix + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const uint8_t*>(rgb_hwc), H, W,
                                                                                                  out_nchw, npix);
    MFB_CUDA_OK(cudaGetLastError());
    return MFB_OK;
}

extern "C" int mfb_prep_mask_depth(const void* mask_u8, const float* depth, int N, int H, int W, int factor, float delta,
                                   float* mask_lat, float* depth_lat, int* scratch_n_ints, void* stream) {
    MFB_REQUIRE(mask_u8 && mask_lat && N > 0 && factor > 0 && H % factor == 0 && W % factor == 0, "bad arguments");
    MFB_REQUIRE(!depth || (depth_lat && scratch_n_ints), "depth needs depth_lat and an N-int scratch buffer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const long long hw = static_cast<long long>(H) * W;
    if (depth) {
        MFB_CUDA_OK(cudaMemsetAsync(scratch_n_ints, 0, sizeof(int) * N, st));
        depth_max_kernel<<<dim3(64, N), 256, 0, st>>>(depth, static_cast<const uint8_t*>(mask_u8), hw, scratch_n_ints);
        MFB_CUDA_OK(cudaGetLastError());
    }
    const long long nlat = static_cast<long long>(N) * (H / factor) * (W / factor);
    prep_mask_depth_kernel<<<unsigned((nlat + 255) / 256), 256, 0, st>>>(static_cast<const uint8_t*>(mask_u8), depth, H, W, factor,
                                                                        scratch_n_ints, delta, mask_lat, depth_lat, nlat);
    MFB_CUDA_OK(cudaGetLastError());
    return MFB_OK;
}

extern "C" int mfb_post_image_u8(const float* img_nchw, int N, int H, int W, void* out_hwc, void* stream) {
    MFB_REQUIRE(img_nchw && out_hwc && N > 0, "bad arguments");
    const long long npix = static_cast<long long>(N) * H * W;
    post_image_kernel<<<unsigned((npix + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(img_nchw, H, W,
                                                                                                  static_cast<uint8_t*>(out_hwc), npix);
    MFB_CUDA_OK(cudaGetLastError());
    return MFB_OK;
}

// ------------------------------------------------------------------------------- dataset resize: bicubic (antialiased) + centre crop
// torchvision `transforms.Resize(res, BICUBIC)` + `CenterCrop(res)` on a float tensor (E/dataset/dataset.py:70-76,86-92,155-165:
// RGB / mask / normalised depth).  On tensors torchvision resizes with F.interpolate(mode="bicubic", align_corners=False,
// antialias=True) = ATen's separable `_upsample_bicubic2d_aa`: per output index the taps cover `center +- 2*max(scale,1)` input
// samples, weights = the Keys cubic with a = -0.5 evaluated at (tap - center + 0.5) / max(scale,1), normalised to sum 1; a horizontal
// pass, then a vertical one.  One thread per output element runs the same two passes over its own support (row sums first, then
// the column combination), so results agree with ATen to fp32 rounding.  `step` > 1 evaluates only the crop's pixels
// (step*i, step*j): the depth map is consumed at latent resolution by nearest sampling (pipeline_brushnet.py:1196-1199), so the
// resize, the crop and that sampling are one launch over 64 x 64 outputs per image.
namespace mfb {
__device__ __forceinline__ float cubic_aa(float x) {
    const float a = -0.5f;
    x = fabsf(x);
    if (x < 1.0f) return ((a + 2.0f) * x - (a + 3.0f)) * x * x + 1.0f;
    if (x < 2.0f) return (((x - 5.0f) * x + 8.0f) * x - 4.0f) * a;
    return 0.0f;
}
// taps of output index `o` of a resize in_size -> out_size: first input index, count, and the un-normalised weights' sum
struct AaTaps { int lo, n; float center, inv, total; };
__device__ __forceinline__ AaTaps aa_taps(int o, int in_size, float scale) {
    AaTaps t;
    const float support = scale >= 1.0f ? 2.0f * scale : 2.0f;
    t.inv = scale >= 1.0f ? 1.0f / scale : 1.0f;
    t.center = scale * (static_cast<float>(o) + 0.5f);
    t.lo = max(static_cast<int>(t.center - support + 0.5f), 0);
    t.n = min(static_cast<int>(t.center + support + 0.5f), in_size) - t.lo;
    t.total = 0.f;
    for (int j = 0; j < t.n; ++j) t.total += cubic_aa((static_cast<float>(j + t.lo) - t.center + 0.5f) * t.inv);
    return t;
}
__global__ void resize_crop_kernel(const float* __restrict__ in, int Hs, int Ws, int Hr, int Wr, int top, int left, int step,
                                   int Ho, int Wo, float* __restrict__ out, long long total) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int x = static_cast<int>(i % Wo), y = static_cast<int>((i / Wo) % Ho);
    const long long nc = i / (static_cast<long long>(Wo) * Ho);
    const float* src = in + nc * static_cast<long long>(Hs) * Ws;
    // area_pixel_compute_scale with align_corners = false: in / out
    const float sy = static_cast<float>(Hs) / static_cast<float>(Hr), sx = static_cast<float>(Ws) / static_cast<float>(Wr);
    const AaTaps ty = aa_taps(top + y * step, Hs, sy), tx = aa_taps(left + x * step, Ws, sx);
    float acc = 0.f;
    for (int r = 0; r < ty.n; ++r) {
        const float* row = src + static_cast<long long>(ty.lo + r) * Ws + tx.lo;
        float h = 0.f;
        for (int c = 0; c < tx.n; ++c)
            h += (cubic_aa((static_cast<float>(c + tx.lo) - tx.center + 0.5f) * tx.inv) / tx.total) * __ldg(row + c);
        acc += (cubic_aa((static_cast<float>(r + ty.lo) - ty.center + 0.5f) * ty.inv) / ty.total) * h;
    }
    out[i] = acc;
}
// depth -> 2 * clip(d, 0, dmax) / dmax - 1 at the map's own resolution (dataset.py:131-145), dmax = max over mask > 0 (+ delta)
__global__ void depth_normalize_kernel(const float* __restrict__ depth, const int* __restrict__ maxbits, float delta, long long hw,
                                       float* __restrict__ out, long long total) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const float dmax = __int_as_float(maxbits[i / hw]) + delta;
    out[i] = 2.0f * (fminf(fmaxf(depth[i], 0.0f), dmax) / dmax) - 1.0f;
}
}  // namespace mfb

extern "C" int mfb_resize_crop_bicubic(const float* in, int NC, int Hs, int Ws, int res, int step, float* out, void* stream) {
    MFB_REQUIRE(in && out && NC > 0 && Hs > 0 && Ws > 0 && res > 0 && step > 0 && res % step == 0, "bad arguments");
    // torchvision _compute_resized_output_size: the shorter side becomes `res`, the longer int(res * long / short)
    int Hr, Wr;
    if (Hs <= Ws) { Hr = res; Wr = static_cast<int>(static_cast<long long>(res) * Ws / Hs); }
    else { Wr = res; Hr = static_cast<int>(static_cast<long long>(res) * Hs / Ws); }
    // CenterCrop: int(round((size - crop) / 2.0)) with Python's round-half-to-even
    auto half_even = [](int d) { return (d % 2 == 0) ? d / 2 : ((d / 2) % 2 == 0 ? d / 2 : d / 2 + 1); };
    const int top = half_even(Hr - res), left = half_even(Wr - res);
    const int Ho = res / step, Wo = res / step;
    const long long total = static_cast<long long>(NC) * Ho * Wo;
    resize_crop_kernel<<<static_cast<unsigned>((total + 127) / 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(in, Hs, Ws, Hr, Wr, top, left,
                                                                                                              step, Ho, Wo, out, total);
    MFB_CUDA_OK(cudaGetLastError());
    return MFB_OK;
}

extern "C" int mfb_depth_normalize(const float* depth, const void* mask_u8, int N, int H, int W, float delta, float* out,
                                   int* scratch_n_ints, void* stream) {
    MFB_REQUIRE(depth && mask_u8 && out && scratch_n_ints && N > 0 && H > 0 && W > 0, "bad arguments");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const long long hw = static_cast<long long>(H) * W, total = hw * N;
    MFB_CUDA_OK(cudaMemsetAsync(scratch_n_ints, 0, sizeof(int) * N, st));
    depth_max_kernel<<<dim3(64, N), 256, 0, st>>>(depth, static_cast<const uint8_t*>(mask_u8), hw, scratch_n_ints);
    MFB_CUDA_OK(cudaGetLastError());
    depth_normalize_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(depth, scratch_n_ints, delta, hw, out, total);
    MFB_CUDA_OK(cudaGetLastError());
    return MFB_OK;
}

namespace mfb {
__global__ void latent_sample_kernel(const float* __restrict__ mean, const float* __restrict__ logvar, const float* __restrict__ noise,
                                     float scale, float* __restrict__ out, long long n) {
    pdl_trigger();
    pdl_wait();
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float v = mean[i];
    if (noise) v += expf(0.5f * fminf(fmaxf(logvar[i], -30.0f), 20.0f)) * noise[i];
    out[i] = scale * v;
}
}  // namespace mfb

extern "C" int mfb_latent_sample(const float* mean, const float* logvar, const float* noise, float scale, float* out, long long n,
                                 void* stream) {
    MFB_RECORD(mfb_latent_sample(mean, logvar, noise, scale, out, n, stream));
    MFB_REQUIRE(mean && out && (noise == nullptr || logvar != nullptr) && n > 0, "bad arguments");
    MFB_CUDA_OK(launch_k(latent_sample_kernel, dim3(static_cast<unsigned>((n + 255) / 256)), dim3(256), 0,
                         static_cast<cudaStream_t>(stream), 1, mean, logvar, noise, scale, out, n));
    return MFB_OK;
}

extern "C" int mfb_cfg_sched_step(const float* eps_uncond, const float* eps_cond, float* x, float* last, float* m0, float* m1,
                                  const float* coef, int Bimg, long long n, void* stream) {
    MFB_RECORD(mfb_cfg_sched_step(eps_uncond, eps_cond, x, last, m0, m1, coef, Bimg, n, stream));
    MFB_REQUIRE(eps_uncond && eps_cond && x && last && m0 && m1 && coef, "null pointer");
    const long long total = static_cast<long long>(Bimg) * n;
    MFB_CUDA_OK(launch_k(cfg_sched_kernel, dim3(grid_for(total, 256)), dim3(256), 0, static_cast<cudaStream_t>(stream), 1,
                         eps_uncond, eps_cond, x, last, m0, m1, coef, total));
    return MFB_OK;
}
