// Weight gradient of conv3x3 / conv1x1 / linear layers on tcgen05 — autograd's weight gradient of F.conv2d / F.linear behind
// LoRACompatibleConv / Linear (S/models/lora.py:363-377,445-451) for the TRAINABLE BrushNet branch of the fine-tune step
// (BASELINE config 4, E/train_brushnet_mirror.py:1459).
//
//   dW[co, tap, ci] = sum over pixels p  dy[p, co] * x[p + shift(tap), ci]
//
// One GEMM per filter tap with M = Cout, N = Cin and the PIXELS as the reduction dimension.  Both tensors are stored
// [pixel][channel], i.e. with the reduction index as the slow dimension — exactly the MN-major operand form of tcgen05.mma: a TMA
// box {64 channels, tw, th, tn} (64 pixels) lands as 64 dense 128-byte rows with the 128B swizzle and is consumed as stored, as the
// A operand (dy) and as the B operand (x); no transposed copy of either tensor exists (the legacy mma.sync version needed
// ldmatrix.trans for the same reason and ran at a quarter of this rate).  The tap shift is a coordinate offset of the x box and
// TMA's out-of-bounds zero fill is the conv padding, as in the forward implicit GEMM (igemm.cu); stride 2 reads the four parity
// views of x.  CTA = (128 output channels) x (BN input channels) of ONE tap over a contiguous range of 64-pixel tiles (split-K
// slices chosen against wave quantisation); the fp32 partial tiles are summed in slice order by wgrad_reduce_kernel ->
// deterministic.  Warps: 0-3 epilogue (TMEM -> fp32 partial tile), 4 TMA producer, 5 MMA issuer.
#include <stdlib.h>
#include <string.h>

#include "common.h"
#include "ptx.cuh"
#include "wgrad5.h"

namespace mfb {

constexpr int W5_BM = 128;
constexpr int W5_BK = 64;        // pixels per K block

struct Wgrad5Params {
    CUtensorMap tmA;             // dy [B, Ho, Wo, Cout]
    CUtensorMap tmB[4];          // x: stride 1 -> [0]; stride 2 -> the four parity views (element (h2, w2) = x[2 h2 + ph, 2 w2 + pw])
    int tap_map[9], tap_dh[9], tap_dw[9];
    int taps, Cin, Cout, ktot, mtiles;
    int tw, th, tn, tiles_w, tiles_h, ptiles;      // 64-pixel tile box and the pixel-tile grid
    int tiles_per_slice;
    float* part;                 // [slices][Cout][ktot]
    int accumulate;              // single slice writing straight into dw: add to what is there (one writer per element: deterministic)
};

template <int BN>
struct W5Cfg {
    static constexpr int NB = (BN + 63) / 64;                       // 64-channel boxes of the x tile
    static constexpr int A_BYTES = 2 * W5_BK * 128;                 // 128 channels x 64 pixels
    static constexpr int B_BYTES = NB * W5_BK * 128;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int STAGES = BN > 128 ? 4 : 6;
    static constexpr int TMEM_COLS = BN <= 64 ? 64 : BN <= 128 ? 128 : 256;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
    static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
};

template <int BN>
__global__ void __launch_bounds__(192, 1) wgrad5_kernel(const __grid_constant__ Wgrad5Params p) {
    using Cfg = W5Cfg<BN>;
    constexpr int STAGES = Cfg::STAGES, NB = Cfg::NB;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar = base + STAGES * Cfg::STAGE_BYTES;
    auto full_bar = [&](int s) { return bar + 8u * s; };
    auto empty_bar = [&](int s) { return bar + 8u * (STAGES + s); };
    const uint32_t acc_bar = bar + 8u * (2 * STAGES);
    const uint32_t tmem_slot = bar + 8u * (2 * STAGES + 1);
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
    const int warp = __shfl_sync(0xffffffffu, int(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const int tap = blockIdx.y / p.mtiles, mt = blockIdx.y % p.mtiles;
    const int co0 = mt * W5_BM, ci0 = blockIdx.x * BN;
    const int t_lo = blockIdx.z * p.tiles_per_slice;
    const int t_hi = min(t_lo + p.tiles_per_slice, p.ptiles);

    if (warp == 4 && lane == 0) {
        prefetch_tmap(&p.tmA);
        prefetch_tmap(&p.tmB[p.tap_map[tap]]);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        mbar_init(acc_bar, 1);
        fence_barrier_init();
    }
    if (warp == 5) {
        tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tacc = *tmem_slot_ptr;
    pdl_trigger();
    pdl_wait();

    if (warp == 4) {
        if (elect_one()) {
            const void* tmB = &p.tmB[p.tap_map[tap]];
            const int dh = p.tap_dh[tap], dw = p.tap_dw[tap];
            int stage = 0;
            uint32_t phase = 0;
            for (int t = t_lo; t < t_hi; ++t) {
                const int iw = t % p.tiles_w, ih = (t / p.tiles_w) % p.tiles_h, ig = t / (p.tiles_w * p.tiles_h);
                const int w0 = iw * p.tw, h0 = ih * p.th, n0 = ig * p.tn;
                mbar_wait(empty_bar(stage), phase ^ 1);
                mbar_expect_tx(full_bar(stage), Cfg::STAGE_BYTES);
                const uint32_t a_dst = base + stage * Cfg::STAGE_BYTES, b_dst = a_dst + Cfg::A_BYTES;
                tma_load_4d(a_dst, &p.tmA, full_bar(stage), co0, w0, h0, n0);
                tma_load_4d(a_dst + W5_BK * 128, &p.tmA, full_bar(stage), co0 + 64, w0, h0, n0);
#pragma unroll
                for (int j = 0; j < NB; ++j)
                    tma_load_4d(b_dst + j * W5_BK * 128, tmB, full_bar(stage), ci0 + 64 * j, w0 + dw, h0 + dh, n0);
                if (++stage == STAGES) { stage = 0; phase ^= 1u; }
            }
        }
        __syncwarp();
    } else if (warp == 5) {
        if (elect_one()) {
            constexpr uint32_t idesc = make_idesc_bf16(W5_BM, BN, 1, 1);        // both operands MN-major
            int stage = 0;
            uint32_t phase = 0;
            for (int t = t_lo; t < t_hi; ++t) {
                mbar_wait(full_bar(stage), phase);
                tc_fence_after();
                const uint32_t a_addr = base + stage * Cfg::STAGE_BYTES, b_addr = a_addr + Cfg::A_BYTES;
#pragma unroll
                for (int ks = 0; ks < W5_BK / 16; ++ks) {
                    // a K step of 16 pixels = 16 rows of 128 B; LBO = one 64-channel box (64 rows)
                    const uint64_t ad = make_desc_mn_sw128(a_addr + ks * 16 * 128, W5_BK * 128);
                    const uint64_t bd = make_desc_mn_sw128(b_addr + ks * 16 * 128, W5_BK * 128);
                    umma_bf16(tacc, ad, bd, idesc, (t != t_lo) || (ks != 0));
                }
                umma_commit(empty_bar(stage));
                if (++stage == STAGES) { stage = 0; phase ^= 1u; }
            }
            umma_commit(acc_bar);
        }
        __syncwarp();
    } else {
        // ===== epilogue: thread = output channel row; fp32 partial tile to the slice's buffer =====
        const int r = warp * 32 + lane;
        const int co = co0 + r;
        mbar_wait_relaxed(acc_bar, 0);
        tc_fence_after();
        float* dst = p.part + (static_cast<size_t>(blockIdx.z) * p.Cout + co) * p.ktot + static_cast<size_t>(tap) * p.Cin + ci0;
        const uint32_t trow = tacc + (uint32_t(warp * 32) << 16);
#pragma unroll
        for (int c = 0; c < BN / 16; ++c) {
            uint32_t v[16];
            tmem_ld16(trow + c * 16, v);
            tmem_wait_ld();
            if (co < p.Cout) {
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (ci0 + c * 16 + q * 4 < p.Cin) {    // Cin % 8 == 0: a float4 is valid as a whole
                        float4 o = make_float4(__uint_as_float(v[q * 4]), __uint_as_float(v[q * 4 + 1]), __uint_as_float(v[q * 4 + 2]),
                                               __uint_as_float(v[q * 4 + 3]));
                        float4* d4 = reinterpret_cast<float4*>(dst + c * 16 + q * 4);
                        if (p.accumulate) {
                            const float4 old = *d4;
                            o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
                        }
                        *d4 = o;
                    }
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 5) {
        tc_fence_after();
        tmem_dealloc(tacc, Cfg::TMEM_COLS);
    }
}

static void pick_tile64(int W, int H, int B, int* tw, int* th, int* tn) {
    long best = -1;
    for (int a = 64; a >= 1; a >>= 1)
        for (int b = 64 / a; b >= 1; b >>= 1) {
            const int c = 64 / (a * b);
            if (a * b * c != 64) continue;
            const long tiles = long((W + a - 1) / a) * ((H + b - 1) / b) * ((B + c - 1) / c);
            if (best < 0 || tiles < best) {
                best = tiles;
                *tw = a; *th = b; *tn = c;
            }
        }
}

static int pick_bn(int Cin) { return Cin % 160 == 0 ? 160 : (Cin >= 128 ? 128 : 64); }

bool wgrad5_supported(int Cin, int Cout) { return Cin % 8 == 0 && Cout % 8 == 0 && Cin >= 64 && Cout >= 64; }

void wgrad5_plan(int B, int Ho, int Wo, int Cin, int Cout, int ksize, int* slices, int* tiles_per_slice) {
    int tw, th, tn;
    pick_tile64(Wo, Ho, B, &tw, &th, &tn);
    const int ptiles = ((Wo + tw - 1) / tw) * ((Ho + th - 1) / th) * ((B + tn - 1) / tn);
    const int bn = pick_bn(Cin);
    const long tiles = long((Cin + bn - 1) / bn) * ((Cout + W5_BM - 1) / W5_BM) * ksize * ksize;
    const int wave = device_sm_count() > 0 ? device_sm_count() : 148;
    int smax = ptiles / 8;           // at least 8 K blocks per slice
    if (smax > 32) smax = 32;
    if (smax < 1) smax = 1;
    int best = 1;
    long long best_cost = -1;
    for (int s = 1; s <= smax; ++s) {
        const long long waves = (tiles * s + wave - 1) / wave;
        const long long cost = waves * ((ptiles + s - 1) / s + 6) + 2 * s;      // + fixed cost per CTA, + one partial tile per slice
        if (best_cost < 0 || cost < best_cost) {
            best_cost = cost;
            best = s;
        }
    }
    const int per = (ptiles + best - 1) / best;
    *tiles_per_slice = per;
    *slices = (ptiles + per - 1) / per;
}

template <int BN>
static int launch_w5(const Wgrad5Params& p, dim3 grid, cudaStream_t st) {
    using Cfg = W5Cfg<BN>;
    static bool configured = false;
    if (!configured) {
        MFB_CUDA_OK(cudaFuncSetAttribute(wgrad5_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
        configured = true;
    }
    MFB_CUDA_OK(launch_k(wgrad5_kernel<BN>, grid, dim3(192), Cfg::SMEM_BYTES, st, 1, p));
    return MFB_OK;
}

// tensor map over NHWC [B, Hf, Wf, C] restricted to pixels (step h + py, step w + px): dims (C, Wv, Hv, B)
static int encode_view(CUtensorMap* tm, const void* base, int C, int Wf, int Hf, int B, int step, int py, int px, const uint32_t* box) {
    const int Wv = (Wf - px + step - 1) / step, Hv = (Hf - py + step - 1) / step;
    const uint64_t dims[4] = {uint64_t(C), uint64_t(Wv), uint64_t(Hv), uint64_t(B)};
    const uint64_t str[3] = {uint64_t(C) * 2 * step, uint64_t(Wf) * C * 2 * step, uint64_t(Hf) * Wf * C * 2};
    const char* b = static_cast<const char*>(base) + (size_t(py) * Wf + px) * C * 2;
    return encode_tmap_bf16(tm, b, 4, dims, str, box, 128);
}

int wgrad5_run(const void* x, const void* dy, int B, int Ho, int Wo, int Cin, int Cout, int ksize, int stride, float* part, int accumulate,
               cudaStream_t st) {
    Wgrad5Params p;
    memset(&p, 0, sizeof(p));
    p.accumulate = accumulate;
    pick_tile64(Wo, Ho, B, &p.tw, &p.th, &p.tn);
    p.tiles_w = (Wo + p.tw - 1) / p.tw;
    p.tiles_h = (Ho + p.th - 1) / p.th;
    p.ptiles = p.tiles_w * p.tiles_h * ((B + p.tn - 1) / p.tn);
    int slices;
    wgrad5_plan(B, Ho, Wo, Cin, Cout, ksize, &slices, &p.tiles_per_slice);
    p.taps = ksize * ksize;
    p.Cin = Cin; p.Cout = Cout; p.ktot = p.taps * Cin;
    p.mtiles = (Cout + W5_BM - 1) / W5_BM;
    p.part = part;
    const uint32_t box[4] = {64u, uint32_t(p.tw), uint32_t(p.th), uint32_t(p.tn)};
    int rc = encode_view(&p.tmA, dy, Cout, Wo, Ho, B, 1, 0, 0, box);
    if (rc) return rc;
    const int H = Ho * stride, W = Wo * stride;
    if (stride == 1) {
        rc = encode_view(&p.tmB[0], x, Cin, W, H, B, 1, 0, 0, box);
        if (rc) return rc;
        const int r = ksize / 2;
        int t = 0;
        for (int kh = -r; kh <= r; ++kh)
            for (int kw = -r; kw <= r; ++kw, ++t) { p.tap_map[t] = 0; p.tap_dh[t] = kh; p.tap_dw[t] = kw; }
    } else {
        for (int ph = 0; ph < 2; ++ph)
            for (int pw = 0; pw < 2; ++pw) {
                rc = encode_view(&p.tmB[ph * 2 + pw], x, Cin, W, H, B, 2, ph, pw, box);
                if (rc) return rc;
            }
        // padding 1: input row = 2 oh + kh - 1:  kh = 0 -> (parity 1, h2 = oh - 1); kh = 1 -> (0, oh); kh = 2 -> (1, oh)
        const int par[3] = {1, 0, 1}, off[3] = {-1, 0, 0};
        for (int kh = 0; kh < 3; ++kh)
            for (int kw = 0; kw < 3; ++kw) {
                p.tap_map[kh * 3 + kw] = par[kh] * 2 + par[kw];
                p.tap_dh[kh * 3 + kw] = off[kh];
                p.tap_dw[kh * 3 + kw] = off[kw];
            }
    }
    const int bn = pick_bn(Cin);
    dim3 grid((Cin + bn - 1) / bn, p.mtiles * p.taps, slices);
    MFB_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "weight-gradient grid too large");
    switch (bn) {
        case 160: return launch_w5<160>(p, grid, st);
        case 128: return launch_w5<128>(p, grid, st);
        default: return launch_w5<64>(p, grid, st);
    }
}

}  // namespace mfb
