// Flash-style scaled-dot-product attention on tcgen05 (replaces F.scaled_dot_product_attention in
// AttnProcessor2_0.__call__, S/models/attention_processor.py:1266-1268; no mask, non-causal).
//
// One CTA = one 128-query tile of one (batch, head).  Per 128-key tile:
//   warp4 (TMA)     : K tile [128 keys x d] and V^T tile [d x 128 keys] -> swizzled smem ring
//   warp5 (MMA)     : S = Q K^T  (M128 N128, K = d in 16-steps)  -> TMEM cols [0,128)
//                     O += P V   (M128 N=dpad, K = 128 keys)     -> TMEM cols [128,128+dpad)
//   warps0-3 (128 t): thread = query row: tcgen05.ld S row, online softmax in the exp2 domain with lazy
//                     (threshold 2^8) rescaling of O in TMEM, P -> bf16 -> smem in the K-major 128B-swizzled
//                     layout the PV MMA reads, final O / l -> bf16 global.
// Head dims 40/80/160 are not multiples of 64: Q/K boxes are 64 columns wide starting at h*d, the MMA K extent
// is d rounded up to 16, and for d % 16 != 0 the softmax threads zero the Q columns [d, dpad) in smem once, so
// the neighbouring head's K columns that ride along contribute nothing.  V is consumed transposed
// ([B, heads*d, keys]) so that both MMAs use the same K-major operand form as the GEMM kernel.
#include <math.h>
#include <string.h>

#include "common.h"
#include "ptx.cuh"

namespace mfb {

constexpr int ATT_THREADS = 192;
constexpr int BQ = 128;   // query rows per CTA
constexpr int BKV = 128;  // keys per tile

template <int D>
struct AttCfg {
    static constexpr int DPAD = (D + 15) / 16 * 16;       // MMA K extent of QK^T and N extent of PV
    static constexpr int NKB = (D + 63) / 64;             // 64-column boxes per Q / K tile
    static constexpr int STAGES = D > 80 ? 1 : 2;         // K/V ring depth (smem-limited for d = 160)
    static constexpr int Q_BYTES = NKB * BQ * 128;
    static constexpr int K_BYTES = NKB * BKV * 128;
    static constexpr int V_BYTES = 2 * DPAD * 128;        // two 64-key boxes of DPAD rows
    static constexpr int KV_BYTES = K_BYTES + ((V_BYTES + 1023) / 1024) * 1024;
    static constexpr int P_BYTES = 2 * BQ * 128;          // 128 x 128 bf16 as two 64-key blocks
    static constexpr int TMEM_COLS = (128 + DPAD) <= 256 ? 256 : 512;
    static constexpr int O_COL = 128;
    static constexpr int SMEM_BYTES = Q_BYTES + STAGES * KV_BYTES + P_BYTES + 1024 + 256;
};

struct AttParams {
    CUtensorMap tmQ, tmK, tmV;
    __nv_bfloat16* out;
    int ldo;
    int Tq, Tk, heads;
    float scale_log2;  // head_dim^-0.5 * log2(e)
};

__device__ __forceinline__ float ex2f(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

template <int D>
__global__ void __launch_bounds__(ATT_THREADS, (D <= 40) ? 2 : 1) attention_kernel(const __grid_constant__ AttParams p) {
    using Cfg = AttCfg<D>;
    constexpr int DPAD = Cfg::DPAD, NKB = Cfg::NKB, STAGES = Cfg::STAGES;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_u32 = smem_u32(smem_raw);
    const uint32_t base = (raw_u32 + 1023u) & ~1023u;
    const uint32_t q_smem = base;
    const uint32_t kv_smem = q_smem + Cfg::Q_BYTES;
    const uint32_t p_smem = kv_smem + STAGES * Cfg::KV_BYTES;
    const uint32_t bar = p_smem + Cfg::P_BYTES;
    const uint32_t q_full = bar, q_ready = bar + 8, s_full = bar + 16, p_full = bar + 24, o_full = bar + 32;
    const uint32_t s_free = bar + 40;
    // K and V tiles have separate full/empty barriers: a K slot is free as soon as Q K^T has read it, long before
    // the P V product of the same tile — so the K tile two steps ahead is already in flight
    auto k_full = [&](int s) { return bar + 48u + 8u * s; };
    auto k_empty = [&](int s) { return bar + 48u + 8u * (STAGES + s); };
    auto v_full = [&](int s) { return bar + 48u + 8u * (2 * STAGES + s); };
    auto v_empty = [&](int s) { return bar + 48u + 8u * (3 * STAGES + s); };
    const uint32_t tmem_slot = bar + 48u + 8u * (4 * STAGES);
    uint8_t* gen_base = smem_raw + (base - raw_u32);  // generic pointer to `base`

    // warp index broadcast from lane 0: the role branches are provably warp-uniform; the issuing thread of the TMA /
    // MMA warps is picked with elect.sync so ptxas emits UTMALDG / UTCHMMA / UTCBAR without ELECT waterfall loops
    const int warp = __shfl_sync(0xffffffffu, int(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * BQ;
    const int h = blockIdx.y;
    const int b = blockIdx.z;
    const int ntiles = (p.Tk + BKV - 1) / BKV;
    constexpr bool kZeroPad = (D % 16) != 0;

    if (warp == 4 && lane == 0) {
        prefetch_tmap(&p.tmQ);
        prefetch_tmap(&p.tmK);
        prefetch_tmap(&p.tmV);
        mbar_init(q_full, 1);
        mbar_init(q_ready, 128);
        mbar_init(s_full, 1);
        mbar_init(p_full, 128);
        mbar_init(o_full, 1);
        mbar_init(s_free, 4);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(k_full(s), 1); mbar_init(k_empty(s), 1);
            mbar_init(v_full(s), 1); mbar_init(v_empty(s), 1);
        }
        fence_barrier_init();
    }
    if (warp == 5) {
        tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<uint32_t*>(gen_base + (tmem_slot - base));
    pdl_trigger();
    pdl_wait();

    // warps 0-3 = softmax (TMEM lane quarter = warp), warp 4 = TMA, warp 5 = MMA: the single-thread issuers get the
    // highest warp ids so the arbiter never lets softmax warps starve them
    if (warp == 4) {
        if (elect_one()) {
            // ===== TMA producer =====
            mbar_expect_tx(q_full, Cfg::Q_BYTES);
            for (int kb = 0; kb < NKB; ++kb) tma_load_3d(q_smem + kb * BQ * 128, &p.tmQ, q_full, h * D + kb * 64, q0, b);
            auto load_k = [&](int j) {
                const int s = j % STAGES;
                mbar_wait_relaxed(k_empty(s), ((j / STAGES) & 1) ^ 1);
                mbar_expect_tx(k_full(s), Cfg::K_BYTES);
                const uint32_t kd = kv_smem + s * Cfg::KV_BYTES;
                for (int kb = 0; kb < NKB; ++kb) tma_load_3d(kd + kb * BKV * 128, &p.tmK, k_full(s), h * D + kb * 64, j * BKV, b);
            };
            load_k(0);
            for (int j = 0; j < ntiles; ++j) {
                if (j + 1 < ntiles) load_k(j + 1);        // K runs one tile ahead of V
                const int s = j % STAGES;
                mbar_wait_relaxed(v_empty(s), ((j / STAGES) & 1) ^ 1);
                mbar_expect_tx(v_full(s), Cfg::V_BYTES);
                const uint32_t vd = kv_smem + s * Cfg::KV_BYTES + Cfg::K_BYTES;
                tma_load_3d(vd, &p.tmV, v_full(s), j * BKV, h * D, b);
                tma_load_3d(vd + DPAD * 128, &p.tmV, v_full(s), j * BKV + 64, h * D, b);
            }
        }
        __syncwarp();
    } else if (warp == 5) {
        if (elect_one()) {
            // ===== MMA issuer =====
            constexpr uint32_t idesc_qk = make_idesc_bf16(128, BKV);
            constexpr uint32_t idesc_pv = make_idesc_bf16(128, DPAD);
            const uint32_t tS = tmem_base, tO = tmem_base + Cfg::O_COL;
            auto issue_qk = [&](int j) {
                const int s = j % STAGES;
                mbar_wait_relaxed(k_full(s), (j / STAGES) & 1);
                tc_fence_after();
                const uint32_t kd = kv_smem + s * Cfg::KV_BYTES;
#pragma unroll
                for (int ks = 0; ks < DPAD / 16; ++ks) {
                    const uint64_t ad = make_desc_k_sw128(q_smem + (ks / 4) * BQ * 128) + uint64_t(2 * (ks % 4));
                    const uint64_t bd = make_desc_k_sw128(kd + (ks / 4) * BKV * 128) + uint64_t(2 * (ks % 4));
                    umma_bf16(tS, ad, bd, idesc_qk, ks != 0);
                }
                umma_commit(s_full);
                umma_commit(k_empty(s));
            };
            mbar_wait(q_full, 0);
            if (kZeroPad) mbar_wait(q_ready, 0);
            issue_qk(0);
            for (int j = 0; j < ntiles; ++j) {
                const int s = j % STAGES;
                // The softmax warps copy S(j) into registers and hand the TMEM buffer back at once (s_free), so
                // S(j+1) = Q K(j+1)^T is computed while they are still exponentiating S(j).  Needs the K/V tile of
                // j+1 resident next to tile j's (K slots are recycled independently of V slots).
                if (j + 1 < ntiles) {
                    mbar_wait_relaxed(s_free, j & 1);
                    tc_fence_after();
                    issue_qk(j + 1);
                }
                mbar_wait_relaxed(v_full(s), (j / STAGES) & 1);
                mbar_wait_relaxed(p_full, j & 1);  // P(j) in smem, O rescaled
                tc_fence_after();
                const uint32_t vd = kv_smem + s * Cfg::KV_BYTES + Cfg::K_BYTES;
#pragma unroll
                for (int ks = 0; ks < BKV / 16; ++ks) {
                    const uint64_t ad = make_desc_k_sw128(p_smem + (ks / 4) * BQ * 128) + uint64_t(2 * (ks % 4));
                    const uint64_t bd = make_desc_k_sw128(vd + (ks / 4) * DPAD * 128) + uint64_t(2 * (ks % 4));
                    umma_bf16(tO, ad, bd, idesc_pv, (j | ks) != 0);
                }
                umma_commit(o_full);
                umma_commit(v_empty(s));
            }
        }
        __syncwarp();
    } else {
        // ===== softmax / correction / epilogue: thread = query row =====
        const int qd = warp & 3;
        const int r = qd * 32 + lane;
        const uint32_t lane_off = uint32_t(qd * 32) << 16;
        const uint32_t tS = tmem_base + lane_off, tO = tmem_base + Cfg::O_COL + lane_off;
        if (kZeroPad) {
            // zero Q columns [D, DPAD) of this row (one 16-byte chunk: D % 8 == 0) in the swizzled tile
            mbar_wait(q_full, 0);
            constexpr int kb = D / 64, ch = (D % 64) / 8;
            uint8_t* qrow = gen_base + (q_smem - base) + kb * BQ * 128 + r * 128 + ((ch ^ (r & 7)) << 4);
            *reinterpret_cast<uint4*>(qrow) = make_uint4(0, 0, 0, 0);
            fence_proxy_async_smem();
            mbar_arrive(q_ready);
        }
        float m_used = -INFINITY, l = 0.f;
        for (int j = 0; j < ntiles; ++j) {
            mbar_wait_relaxed(s_full, j & 1);
            tc_fence_after();
            uint32_t sv[4][32];
            tmem_ld32(tS + 0, sv[0]);
            tmem_ld32(tS + 32, sv[1]);
            tmem_ld32(tS + 64, sv[2]);
            tmem_ld32(tS + 96, sv[3]);
            tmem_wait_ld();
            const int kbase = j * BKV;
            const bool tail = kbase + BKV > p.Tk;
            // S(j) now lives in registers: the MMA warp may overwrite the TMEM buffer with S(j+1)
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(s_free);
            float mx = -INFINITY;      // row max of the RAW scores (scale > 0 commutes with max)
#pragma unroll
            for (int c = 0; c < 4; ++c)
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    float s = __uint_as_float(sv[c][i]);
                    if (tail && kbase + c * 32 + i >= p.Tk) {
                        s = -INFINITY;
                        sv[c][i] = __float_as_uint(s);
                    }
                    mx = fmaxf(mx, s);
                }
            mx *= p.scale_log2;
            float alpha = 1.f;
            if (mx > m_used + 8.f) {  // lazy rescale: keep a stale max while the row max grew by < 2^8
                alpha = ex2f(m_used - mx);
                m_used = mx;
            }
            if (j > 0) {
                mbar_wait_relaxed(o_full, (j - 1) & 1);  // PV(j-1) finished: P buffer free, O stable
                tc_fence_after();
                if (__any_sync(0xffffffffu, alpha != 1.f)) {
#pragma unroll
                    for (int c = 0; c < DPAD / 16; ++c) {
                        uint32_t ov[16];
                        tmem_ld16(tO + c * 16, ov);
                        tmem_wait_ld();
#pragma unroll
                        for (int i = 0; i < 16; ++i) ov[i] = __float_as_uint(__uint_as_float(ov[i]) * alpha);
                        tmem_st16(tO + c * 16, ov);
                    }
                    tmem_wait_st();
                }
            }
            l *= alpha;
            float sum = 0.f;
            uint8_t* prow = gen_base + (p_smem - base) + r * 128;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
#pragma unroll
                for (int i8 = 0; i8 < 4; ++i8) {
                    float e[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        e[i] = ex2f(fmaf(__uint_as_float(sv[c][i8 * 8 + i]), p.scale_log2, -m_used));   // one FFMA + one MUFU
                        sum += e[i];
                    }
                    uint4 pk;
                    pk.x = pack_bf16x2(e[0], e[1]);
                    pk.y = pack_bf16x2(e[2], e[3]);
                    pk.z = pack_bf16x2(e[4], e[5]);
                    pk.w = pack_bf16x2(e[6], e[7]);
                    const int col = c * 32 + i8 * 8;           // key column inside the tile
                    const int blk = col >> 6, ch = (col & 63) >> 3;
                    *reinterpret_cast<uint4*>(prow + blk * BQ * 128 + ((ch ^ (r & 7)) << 4)) = pk;
                }
            }
            l += sum;
            fence_proxy_async_smem();
            tc_fence_before();
            mbar_arrive(p_full);
        }
        // ===== epilogue =====
        mbar_wait(o_full, (ntiles - 1) & 1);
        tc_fence_after();
        const float inv_l = 1.f / l;
        const int qrow = q0 + r;
        __nv_bfloat16* dst = p.out + (static_cast<size_t>(b) * p.Tq + qrow) * p.ldo + h * D;
#pragma unroll
        for (int c = 0; c < DPAD / 16; ++c) {
            uint32_t ov[16];
            tmem_ld16(tO + c * 16, ov);
            tmem_wait_ld();
            if (qrow < p.Tq) {
#pragma unroll
                for (int i8 = 0; i8 < 2; ++i8) {
                    if (c * 16 + i8 * 8 < D) {
                        uint4 pk;
                        pk.x = pack_bf16x2(__uint_as_float(ov[i8 * 8 + 0]) * inv_l, __uint_as_float(ov[i8 * 8 + 1]) * inv_l);
                        pk.y = pack_bf16x2(__uint_as_float(ov[i8 * 8 + 2]) * inv_l, __uint_as_float(ov[i8 * 8 + 3]) * inv_l);
                        pk.z = pack_bf16x2(__uint_as_float(ov[i8 * 8 + 4]) * inv_l, __uint_as_float(ov[i8 * 8 + 5]) * inv_l);
                        pk.w = pack_bf16x2(__uint_as_float(ov[i8 * 8 + 6]) * inv_l, __uint_as_float(ov[i8 * 8 + 7]) * inv_l);
                        *reinterpret_cast<uint4*>(dst + c * 16 + i8 * 8) = pk;
                    }
                }
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 5) {
        tc_fence_after();
        tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}

template <int D>
static int launch_attention(const AttParams& p, int B, cudaStream_t st) {
    using Cfg = AttCfg<D>;
    static bool configured = false;
    if (!configured) {
        MFB_CUDA_OK(cudaFuncSetAttribute(attention_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
        configured = true;
    }
    dim3 grid((p.Tq + BQ - 1) / BQ, p.heads, B);
    MFB_CUDA_OK(launch_k(attention_kernel<D>, grid, dim3(ATT_THREADS), Cfg::SMEM_BYTES, st, 1, p));
    return MFB_OK;
}

}  // namespace mfb

using namespace mfb;

extern "C" int mfb_attention(const void* q, int ldq, const void* k, int ldk, const void* vt, int ldvt, void* out, int ldo,
                             int B, int heads, int head_dim, int Tq, int Tk, void* stream) {
    MFB_REQUIRE(q && k && vt && out, "null pointer");
    MFB_REQUIRE(ldq % 8 == 0 && ldk % 8 == 0 && ldvt % 8 == 0 && ldo % 8 == 0, "leading dimensions must be multiples of 8");
    MFB_REQUIRE(Tq > 0 && Tk > 0 && ldvt >= Tk, "bad sequence lengths");
    AttParams p;
    memset(&p, 0, sizeof(p));
    int dpad = (head_dim + 15) / 16 * 16;
    {
        const uint64_t dims[3] = {uint64_t(ldq), uint64_t(Tq), uint64_t(B)};
        const uint64_t str[2] = {uint64_t(ldq) * 2, uint64_t(Tq) * ldq * 2};
        const uint32_t box[3] = {64, BQ, 1};
        int rc = encode_tmap_bf16(&p.tmQ, q, 3, dims, str, box, 128);
        if (rc) return rc;
    }
    {
        const uint64_t dims[3] = {uint64_t(ldk), uint64_t(Tk), uint64_t(B)};
        const uint64_t str[2] = {uint64_t(ldk) * 2, uint64_t(Tk) * ldk * 2};
        const uint32_t box[3] = {64, BKV, 1};
        int rc = encode_tmap_bf16(&p.tmK, k, 3, dims, str, box, 128);
        if (rc) return rc;
    }
    {
        const uint64_t rows = uint64_t(heads) * head_dim;
        const uint64_t dims[3] = {uint64_t(ldvt), rows, uint64_t(B)};
        const uint64_t str[2] = {uint64_t(ldvt) * 2, rows * ldvt * 2};
        const uint32_t box[3] = {64, uint32_t(dpad), 1};
        int rc = encode_tmap_bf16(&p.tmV, vt, 3, dims, str, box, 128);
        if (rc) return rc;
    }
    p.out = static_cast<__nv_bfloat16*>(out);
    p.ldo = ldo;
    p.Tq = Tq;
    p.Tk = Tk;
    p.heads = heads;
    p.scale_log2 = 1.4426950408889634f / sqrtf(static_cast<float>(head_dim));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    switch (head_dim) {
        case 32: return launch_attention<32>(p, B, st);
        case 40: return launch_attention<40>(p, B, st);
        case 64: return launch_attention<64>(p, B, st);
        case 80: return launch_attention<80>(p, B, st);
        case 160: return launch_attention<160>(p, B, st);
        default:
            set_error("head_dim %d is not instantiated (supported: 32, 40, 64, 80, 160)", head_dim);
            return MFB_EINVAL;
    }
}
