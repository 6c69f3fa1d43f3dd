// Flash-style scaled-dot-product attention on tcgen05 (replaces F.scaled_dot_product_attention in
// AttnProcessor2_0.__call__, S/models/attention_processor.py:1266-1268; no mask, non-causal).
//
// One CTA = NQ 128-query tiles of one (batch, head) sharing the K/V stream.  Per key tile of BKV keys:
//   TMA warp        : K tile [BKV keys x d] and V tile [BKV keys x d] -> swizzled smem ring
//   MMA warp        : S_t = Q_t K^T  (M128 N=BKV, K = d in 16-steps)  -> TMEM, one S buffer per query tile
//                     O_t += P_t V   (M128 N=dpad, K = BKV keys)      -> TMEM
//   softmax warpgroup t (128 threads, thread = query row): tcgen05.ld S row, online softmax in the exp2 domain with
//                     lazy (threshold 2^8) rescaling of O in TMEM, P -> bf16 -> smem in the K-major 128B-swizzled
//                     layout the PV MMA reads, final O / l -> bf16 global.
// The 4096-token layers (d = 40) are bound by the exponential unit (MUFU: 16 ex2 / clk / SM; the two MMAs need a
// third of that time), so their schedule is built around keeping it busy: NQ = 2 warpgroups take turns in the
// exponential phase (named-barrier ping-pong) — while one exponentiates tile j the other does everything else of its
// own tile (wait, TMEM load, row max, O rescale, fences).  Two independent CTAs per SM do NOT do this: their relative
// phase is arbitrary and neutral-stable, and 57 % of the time both sat in the exponential phase together
// (profiles/r01k, profiles/r01n).
// Head dims 40/80/160 are not multiples of 64: Q/K boxes are 64 columns wide starting at h*d, the MMA K extent
// is d rounded up to 16, and for d % 16 != 0 the softmax threads zero the Q columns [d, dpad) in smem once, so
// the neighbouring head's K columns that ride along contribute nothing.  V is consumed as stored ([B, keys, heads*d],
// e.g. straight out of the fused q|k|v projection): its tile is the MN-major B operand of the P V product, so no
// transposed copy of V exists anywhere (the neighbouring head's columns [d, dpad) only feed accumulator columns that
// are never stored).
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "common.h"
#include "ptx.cuh"

namespace mfb {

constexpr int BQ = 128;   // query rows per tile (= TMEM lanes)
// tuning knobs of the small-head-dim (d <= 40) configuration; A/B record in profiles/r01n_attention_pingpong.md
#ifndef MFB_ATT_NQ_SMALL        // query tiles (softmax warpgroups) per CTA
#define MFB_ATT_NQ_SMALL 2
#endif
#ifndef MFB_ATT_KV_SMALL        // keys per tile
#define MFB_ATT_KV_SMALL 64
#endif
#ifndef MFB_ATT_CTAS_SMALL      // CTAs per SM the kernel is compiled for
#define MFB_ATT_CTAS_SMALL 2
#endif
#ifndef MFB_ATT_PINGPONG        // 1: the two warpgroups alternate in the exponential phase
#define MFB_ATT_PINGPONG 1
#endif
// same knobs for the middle head dims (40 < d <= 80: the 1024-token layers), one CTA per SM
#ifndef MFB_ATT_NQ_MID
#define MFB_ATT_NQ_MID 2
#endif
#ifndef MFB_ATT_KV_MID
#define MFB_ATT_KV_MID 64
#endif
#ifndef MFB_ATT_PINGPONG_MID
#define MFB_ATT_PINGPONG_MID 1
#endif
#ifndef MFB_ATT_HANDOFF_EARLY   // the turn is handed over this many 8-column groups before the end of the exponential phase
#define MFB_ATT_HANDOFF_EARLY 0
#endif
#ifndef MFB_ATT_PACKED_MATH     // 1: FFMA2 / FADD2 in the exponential loop (measured: d = 40 0.744 vs 0.693 ms — register pairs under the 96-register budget; d = 80 0.0746 vs 0.0758)
#define MFB_ATT_PACKED_MATH 0
#endif
#ifndef MFB_ATT_POLY_EVERY      // every n-th exponential of a row on the FMA pipe (ex2_poly) instead of MUFU; 0 = none
#define MFB_ATT_POLY_EVERY 0
#endif

// SHORT = the whole key sequence fits ONE tile of 80 keys (the 77-token CLIP context of every cross attention): one query
// tile per CTA, no K/V ring, no ping-pong, a single softmax pass — instead of two 64-key iterations of which the second is
// 80 % padding.
template <int D, bool SHORT = false>
struct AttCfg {
    static constexpr bool SMALL = D <= 40;
    static constexpr int DPAD = (D + 15) / 16 * 16;       // MMA K extent of QK^T and N extent of PV
    static constexpr int NKB = (D + 63) / 64;             // 64-column boxes per Q / K tile
    static constexpr bool MID = D > 40 && D <= 80;
    static constexpr int NQ = SHORT ? 1 : SMALL ? MFB_ATT_NQ_SMALL : MID ? MFB_ATT_NQ_MID : 1;
    static constexpr int BKV = SHORT ? 80 : SMALL ? MFB_ATT_KV_SMALL : MID ? MFB_ATT_KV_MID : 128;
    static constexpr bool PINGPONG = NQ == 2 && (SMALL ? MFB_ATT_PINGPONG : MFB_ATT_PINGPONG_MID);
    static constexpr int CTAS_PER_SM = SHORT ? (D <= 80 ? 2 : 1) : SMALL ? MFB_ATT_CTAS_SMALL : 1;
    static constexpr int THREADS = (4 * NQ + 2) * 32;     // NQ softmax warpgroups + TMA warp + MMA warp
    static constexpr int STAGES = (SHORT || D > 80) ? 1 : 2;   // K/V ring depth (smem-limited for d = 160; one tile in all if SHORT)
    static constexpr int QT_BYTES = NKB * BQ * 128;       // one query tile
    static constexpr int Q_BYTES = NQ * QT_BYTES;
    static constexpr int K_BYTES = NKB * BKV * 128;
    static constexpr int V_BYTES = K_BYTES;                   // same box shape: BKV key rows x 64-column blocks
    static constexpr int KV_BYTES = K_BYTES + V_BYTES;
    static constexpr int P_BYTES = ((BKV + 63) / 64) * BQ * 128;     // 128 x BKV bf16 as 64-key blocks, per query tile
    static constexpr int TMEM_USED = NQ * (BKV + DPAD);       // S_0..S_{NQ-1} | O_0..O_{NQ-1}
    static constexpr int TMEM_COLS = TMEM_USED <= 128 ? 128 : TMEM_USED <= 256 ? 256 : 512;
    static constexpr int SMEM_BYTES = Q_BYTES + STAGES * KV_BYTES + NQ * P_BYTES + 1024 + 256;
    static_assert(TMEM_USED <= 512 && TMEM_COLS * CTAS_PER_SM <= 512, "TMEM over-subscribed");
};

struct AttParams {
    CUtensorMap tmQ, tmK, tmV;
    __nv_bfloat16* out;
    int ldo;
    int Tq, Tk, heads;
    float scale_log2;  // head_dim^-0.5 * log2(e)
    float* lse;        // optional [B, heads, Tq]: log2 sum_j 2^(s_ij) of every query row (kept for the backward pass, attn_bwd.cu)
};

__device__ __forceinline__ float ex2f(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// same, but ordered against the barrier instructions around the ping-pong turn (volatile asms keep their order)
__device__ __forceinline__ float ex2f_ordered(float x) {
    float y;
    asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

template <int D, bool SHORT>
__global__ void __launch_bounds__(AttCfg<D, SHORT>::THREADS, AttCfg<D, SHORT>::CTAS_PER_SM) attention_kernel(const __grid_constant__ AttParams p) {
    using Cfg = AttCfg<D, SHORT>;
    constexpr int BKV = Cfg::BKV, NQ = Cfg::NQ;
    constexpr int DPAD = Cfg::DPAD, NKB = Cfg::NKB, STAGES = Cfg::STAGES;
    constexpr int TMA_WARP = 4 * NQ, MMA_WARP = 4 * NQ + 1;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_u32 = smem_u32(smem_raw);
    const uint32_t base = (raw_u32 + 1023u) & ~1023u;
    const uint32_t q_smem = base;
    const uint32_t kv_smem = q_smem + Cfg::Q_BYTES;
    const uint32_t p_smem = kv_smem + STAGES * Cfg::KV_BYTES;
    const uint32_t bar = p_smem + NQ * Cfg::P_BYTES;
    auto bar_at = [&](int i) { return bar + 8u * i; };
    const uint32_t q_full = bar_at(0), q_ready = bar_at(1);
    auto s_full = [&](int t) { return bar_at(2 + t); };    // S_t(j) is in TMEM
    auto s_free = [&](int t) { return bar_at(4 + t); };    // S_t(j) has been copied to registers
    auto p_full = [&](int t) { return bar_at(6 + t); };    // P_t(j) is in smem, O_t rescaled
    auto pv_done = [&](int t) { return bar_at(8 + t); };   // O_t += P_t(j) V(j) finished
    // K and V tiles have separate full/empty barriers: a K slot is free as soon as Q K^T has read it, long before
    // the P V product of the same tile — so the K tile two steps ahead is already in flight
    auto k_full = [&](int s) { return bar_at(10 + s); };
    auto k_empty = [&](int s) { return bar_at(10 + STAGES + s); };
    auto v_full = [&](int s) { return bar_at(10 + 2 * STAGES + s); };
    auto v_empty = [&](int s) { return bar_at(10 + 3 * STAGES + s); };
    const uint32_t tmem_slot = bar_at(10 + 4 * STAGES);
    uint8_t* gen_base = smem_raw + (base - raw_u32);  // generic pointer to `base`

    // warp index broadcast from lane 0: the role branches are provably warp-uniform; the issuing thread of the TMA /
    // MMA warps is picked with elect.sync so ptxas emits UTMALDG / UTCHMMA / UTCBAR without ELECT waterfall loops
    const int warp = __shfl_sync(0xffffffffu, int(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * (NQ * BQ);
    const int h = blockIdx.y;
    const int b = blockIdx.z;
    const int ntiles = (p.Tk + BKV - 1) / BKV;
    constexpr bool kZeroPad = (D % 16) != 0;

    if (warp == TMA_WARP && lane == 0) {
        prefetch_tmap(&p.tmQ);
        prefetch_tmap(&p.tmK);
        prefetch_tmap(&p.tmV);
        mbar_init(q_full, 1);
        mbar_init(q_ready, NQ * 128);
        for (int t = 0; t < NQ; ++t) {
            mbar_init(s_full(t), 1);
            mbar_init(s_free(t), 4);
            mbar_init(p_full(t), 128);
            mbar_init(pv_done(t), 1);
        }
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(k_full(s), 1); mbar_init(k_empty(s), 1);
            mbar_init(v_full(s), 1); mbar_init(v_empty(s), 1);
        }
        fence_barrier_init();
    }
    if (warp == MMA_WARP) {
        tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<uint32_t*>(gen_base + (tmem_slot - base));
    pdl_trigger();
    pdl_wait();

    // warps [0, 4 NQ) = softmax (TMEM lane quarter = warp % 4), then the TMA warp, then the MMA warp: the
    // single-thread issuers get the highest warp ids so the arbiter never lets softmax warps starve them
    if (warp == TMA_WARP) {
        if (elect_one()) {
            // ===== TMA producer =====
            mbar_expect_tx(q_full, Cfg::Q_BYTES);
            for (int t = 0; t < NQ; ++t)
                for (int kb = 0; kb < NKB; ++kb)
                    tma_load_3d(q_smem + t * Cfg::QT_BYTES + kb * BQ * 128, &p.tmQ, q_full, h * D + kb * 64, q0 + t * BQ, b);
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
                for (int kb = 0; kb < NKB; ++kb) tma_load_3d(vd + kb * BKV * 128, &p.tmV, v_full(s), h * D + kb * 64, j * BKV, b);
            }
        }
        __syncwarp();
    } else if (warp == MMA_WARP) {
        if (elect_one()) {
            // ===== MMA issuer =====
            constexpr uint32_t idesc_qk = make_idesc_bf16(128, BKV);
            constexpr uint32_t idesc_pv = make_idesc_bf16(128, DPAD, 0, 1);     // B = V tile, MN-major
            // S_t(j) = Q_t K(j)^T for every query tile; the K slot is released after the last one
            auto issue_qk = [&](int j, bool wait_free) {
                const int s = j % STAGES;
                mbar_wait_relaxed(k_full(s), (j / STAGES) & 1);
                const uint32_t kd = kv_smem + s * Cfg::KV_BYTES;
#pragma unroll
                for (int t = 0; t < NQ; ++t) {
                    // The softmax warps copy S_t(j-1) into registers and hand the TMEM buffer back at once (s_free),
                    // so S_t(j) is computed while they are still exponentiating S_t(j-1).
                    if (wait_free) mbar_wait_relaxed(s_free(t), (j - 1) & 1);
                    tc_fence_after();
#pragma unroll
                    for (int ks = 0; ks < DPAD / 16; ++ks) {
                        const uint64_t ad = make_desc_k_sw128(q_smem + t * Cfg::QT_BYTES + (ks / 4) * BQ * 128) + uint64_t(2 * (ks % 4));
                        const uint64_t bd = make_desc_k_sw128(kd + (ks / 4) * BKV * 128) + uint64_t(2 * (ks % 4));
                        umma_bf16(tmem_base + t * BKV, ad, bd, idesc_qk, ks != 0);
                    }
                    umma_commit(s_full(t));
                }
                umma_commit(k_empty(s));
            };
            mbar_wait(q_full, 0);
            if (kZeroPad) mbar_wait(q_ready, 0);
            issue_qk(0, false);
            for (int j = 0; j < ntiles; ++j) {
                const int s = j % STAGES;
                if (j + 1 < ntiles) issue_qk(j + 1, true);      // needs the K tile of j+1 resident next to tile j's
                mbar_wait_relaxed(v_full(s), (j / STAGES) & 1);
                const uint32_t vd = kv_smem + s * Cfg::KV_BYTES + Cfg::K_BYTES;
#pragma unroll
                for (int t = 0; t < NQ; ++t) {
                    mbar_wait_relaxed(p_full(t), j & 1);  // P_t(j) in smem, O_t rescaled
                    tc_fence_after();
#pragma unroll
                    for (int ks = 0; ks < BKV / 16; ++ks) {
                        const uint64_t ad = make_desc_k_sw128(p_smem + t * Cfg::P_BYTES + (ks / 4) * BQ * 128) + uint64_t(2 * (ks % 4));
                        const uint64_t bd = make_desc_mn_sw128(vd + ks * 16 * 128, BKV * 128);      // 16 key rows per K step
                        umma_bf16(tmem_base + NQ * BKV + t * DPAD, ad, bd, idesc_pv, (j | ks) != 0);
                    }
                    umma_commit(pv_done(t));
                }
                umma_commit(v_empty(s));
            }
        }
        __syncwarp();
    } else {
        // ===== softmax / correction / epilogue: warpgroup t owns query tile t, thread = query row =====
        const int t = warp >> 2;
        const int qd = warp & 3;
        const int r = qd * 32 + lane;
        const uint32_t lane_off = uint32_t(qd * 32) << 16;
        const uint32_t tS = tmem_base + t * BKV + lane_off, tO = tmem_base + NQ * BKV + t * DPAD + lane_off;
        constexpr bool kPingPong = Cfg::PINGPONG;
        constexpr int TURN_BAR = 1;                      // named barriers 1, 2: "warpgroup 0 / 1 may exponentiate"
        if (kZeroPad) {
            // zero Q columns [D, DPAD) of this row (one 16-byte chunk: D % 8 == 0) in the swizzled tile
            mbar_wait(q_full, 0);
            constexpr int kb = D / 64, ch = (D % 64) / 8;
            uint8_t* qrow = gen_base + (q_smem - base) + t * Cfg::QT_BYTES + kb * BQ * 128 + r * 128 + ((ch ^ (r & 7)) << 4);
            *reinterpret_cast<uint4*>(qrow) = make_uint4(0, 0, 0, 0);
            fence_proxy_async_smem();
            mbar_arrive(q_ready);
        }
        if (kPingPong && t == 1) named_bar_arrive(TURN_BAR + 0, 256);       // warpgroup 0 goes first
        float m_used = -INFINITY, l = 0.f;
        uint8_t* prow = gen_base + (p_smem - base) + t * Cfg::P_BYTES + r * 128;
        for (int j = 0; j < ntiles; ++j) {
            mbar_wait_relaxed(s_full(t), j & 1);
            tc_fence_after();
            constexpr int CW = (BKV % 32 == 0) ? 32 : 16;      // TMEM load width: 32-column chunks (16 for the 80-key tile)
            constexpr int NC = BKV / CW;
            uint32_t sv[NC][CW];
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                if constexpr (CW == 32) tmem_ld32(tS + c * 32, sv[c]);
                else tmem_ld16(tS + c * 16, sv[c]);
            }
            tmem_wait_ld();
            const int kbase = j * BKV;
            const bool tail = kbase + BKV > p.Tk;
            // S(j) now lives in registers: the MMA warp may overwrite the TMEM buffer with S(j+1)
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(s_free(t));
            // key padding exists only in the last tile: mask it in a separate (warp-uniform) block so the main path
            // carries no per-element index arithmetic
            if (tail) {
#pragma unroll
                for (int c = 0; c < NC; ++c)
#pragma unroll
                    for (int i = 0; i < CW; ++i)
                        if (kbase + c * CW + i >= p.Tk) sv[c][i] = 0xff800000u;     // -inf
            }
            // row max of the RAW scores (scale > 0 commutes with max), four independent chains
            float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
            for (int c = 0; c < NC; ++c)
#pragma unroll
                for (int i = 0; i < CW; ++i) m4[i & 3] = fmaxf(m4[i & 3], __uint_as_float(sv[c][i]));
            const float mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])) * p.scale_log2;
            float alpha = 1.f;
            if (mx > m_used + 8.f) {  // lazy rescale: keep a stale max while the row max grew by < 2^8
                alpha = ex2f(m_used - mx);
                m_used = mx;
            }
            if (j > 0) {
                mbar_wait_relaxed(pv_done(t), (j - 1) & 1);  // PV(j-1) finished: P buffer free, O stable
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
            if (kPingPong) named_bar_sync(TURN_BAR + t, 256);               // my turn on the exponential unit
            float s2[2] = {0.f, 0.f};
#pragma unroll
            for (int c = 0; c < NC; ++c) {
#pragma unroll
                for (int i8 = 0; i8 < CW / 8; ++i8) {
                    float e[8];
#if MFB_ATT_PACKED_MATH
                    // pairwise: one FFMA2 (scale, subtract the max) and one FADD2 (row sums) per two scores
#pragma unroll
                    for (int i = 0; i < 8; i += 2) {
                        const float2 x = ffma2(make_float2(__uint_as_float(sv[c][i8 * 8 + i]), __uint_as_float(sv[c][i8 * 8 + i + 1])),
                                               make_float2(p.scale_log2, p.scale_log2), make_float2(-m_used, -m_used));
                        e[i] = kPingPong ? ex2f_ordered(x.x) : ex2f(x.x);
                        e[i + 1] = kPingPong ? ex2f_ordered(x.y) : ex2f(x.y);
                        const float2 acc = fadd2(make_float2(s2[0], s2[1]), make_float2(e[i], e[i + 1]));
                        s2[0] = acc.x;
                        s2[1] = acc.y;
                    }
#else
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float x = fmaf(__uint_as_float(sv[c][i8 * 8 + i]), p.scale_log2, -m_used);
                        // one FFMA + one MUFU; optionally every n-th on the FMA pipe instead
                        e[i] = (MFB_ATT_POLY_EVERY > 0 && Cfg::SMALL && (i % (MFB_ATT_POLY_EVERY > 0 ? MFB_ATT_POLY_EVERY : 1)) == 0)
                                   ? ex2_poly(x) : (kPingPong ? ex2f_ordered(x) : ex2f(x));
                        s2[i & 1] += e[i];
                    }
#endif
                    uint4 pk;
                    pk.x = pack_bf16x2(e[0], e[1]);
                    pk.y = pack_bf16x2(e[2], e[3]);
                    pk.z = pack_bf16x2(e[4], e[5]);
                    pk.w = pack_bf16x2(e[6], e[7]);
                    const int col = c * CW + i8 * 8;           // key column inside the tile
                    const int blk = col >> 6, ch = (col & 63) >> 3;
                    *reinterpret_cast<uint4*>(prow + blk * BQ * 128 + ((ch ^ (r & 7)) << 4)) = pk;
                    // hand the turn over, optionally a few groups early so the other warpgroup's wake-up latency
                    // overlaps the tail of this one (the very last hand-over of warpgroup 1 would have no taker)
                    if (kPingPong && c * (CW / 8) + i8 == NC * (CW / 8) - 1 - MFB_ATT_HANDOFF_EARLY && !(t == 1 && j == ntiles - 1))
                        named_bar_arrive(TURN_BAR + (t ^ 1), 256);
                }
            }
            l = l * alpha + (s2[0] + s2[1]);
            fence_proxy_async_smem();
            tc_fence_before();
            mbar_arrive(p_full(t));
        }
        // ===== epilogue =====
        mbar_wait(pv_done(t), (ntiles - 1) & 1);
        tc_fence_after();
        const float inv_l = 1.f / l;
        const int qrow = q0 + t * BQ + r;
        if (p.lse != nullptr && qrow < p.Tq) p.lse[(static_cast<size_t>(b) * p.heads + h) * p.Tq + qrow] = m_used + log2f(l);
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
    if (warp == MMA_WARP) {
        tc_fence_after();
        tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}

template <int D, bool SHORT>
static int launch_attention_v(const AttParams& p, int B, cudaStream_t st) {
    using Cfg = AttCfg<D, SHORT>;
    static bool configured = false;
    if (!configured) {
        MFB_CUDA_OK(cudaFuncSetAttribute(attention_kernel<D, SHORT>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
        configured = true;
    }
    dim3 grid((p.Tq + Cfg::NQ * BQ - 1) / (Cfg::NQ * BQ), p.heads, B);
    MFB_CUDA_OK(launch_k(attention_kernel<D, SHORT>, grid, dim3(Cfg::THREADS), Cfg::SMEM_BYTES, st, 1, p));
    return MFB_OK;
}
template <int D>
static int launch_attention(const AttParams& p, int B, bool short_kv, cudaStream_t st) {
    return short_kv ? launch_attention_v<D, true>(p, B, st) : launch_attention_v<D, false>(p, B, st);
}

}  // namespace mfb

using namespace mfb;

static int attention_impl(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, void* out, int ldo, int B, int heads,
                          int head_dim, int Tq, int Tk, float* lse, void* stream) {
    MFB_REQUIRE(q && k && v && out, "null pointer");
    MFB_REQUIRE(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 8 == 0, "leading dimensions must be multiples of 8");
    MFB_REQUIRE(Tq > 0 && Tk > 0, "bad sequence lengths");
    MFB_REQUIRE(ldq >= heads * head_dim && ldk >= heads * head_dim && ldv >= heads * head_dim && ldo >= heads * head_dim,
                "leading dimensions must cover heads * head_dim columns");
    AttParams p;
    memset(&p, 0, sizeof(p));
    // the maps expose exactly the heads*d columns of each tensor (row stride = ld): the 64-column boxes of the last
    // heads reach past them, and those columns must come back as TMA zero fill, not as whatever follows in memory
    // (for a v view into a fused q|k|v buffer that would be the next row — and past the allocation on the last row)
    const uint64_t width = uint64_t(heads) * head_dim;
    // the one-tile variant for short key sequences (cross attention to the 77-token context); MFB_ATT_SHORTKV=0 disables it
    static const bool short_on = [] { const char* e = getenv("MFB_ATT_SHORTKV"); return !e || atoi(e) != 0; }();
    // (measured, profiles/logs/ab_shortkv.log: d = 80 0.0264 -> 0.0212 ms, d = 160 0.0144 -> 0.0125 ms, d = 40 unchanged -> kept on the
    // two-query-tile kernel there)
    const bool short_kv = short_on && Tk <= AttCfg<40, true>::BKV && head_dim > 40;
    const uint32_t kv_box = short_kv ? uint32_t(AttCfg<40, true>::BKV)
                                     : uint32_t(head_dim <= 40 ? AttCfg<40>::BKV : head_dim <= 80 ? AttCfg<80>::BKV : 128);   // keys per tile
    {
        const uint64_t dims[3] = {width, uint64_t(Tq), uint64_t(B)};
        const uint64_t str[2] = {uint64_t(ldq) * 2, uint64_t(Tq) * ldq * 2};
        const uint32_t box[3] = {64, BQ, 1};
        int rc = encode_tmap_bf16(&p.tmQ, q, 3, dims, str, box, 128);
        if (rc) return rc;
    }
    {
        const uint64_t dims[3] = {width, uint64_t(Tk), uint64_t(B)};
        const uint64_t str[2] = {uint64_t(ldk) * 2, uint64_t(Tk) * ldk * 2};
        const uint32_t box[3] = {64, kv_box, 1};
        int rc = encode_tmap_bf16(&p.tmK, k, 3, dims, str, box, 128);
        if (rc) return rc;
    }
    {
        const uint64_t dims[3] = {width, uint64_t(Tk), uint64_t(B)};
        const uint64_t str[2] = {uint64_t(ldv) * 2, uint64_t(Tk) * ldv * 2};
        const uint32_t box[3] = {64, kv_box, 1};
        int rc = encode_tmap_bf16(&p.tmV, v, 3, dims, str, box, 128);
        if (rc) return rc;
    }
    p.out = static_cast<__nv_bfloat16*>(out);
    p.ldo = ldo;
    p.Tq = Tq;
    p.Tk = Tk;
    p.heads = heads;
    p.scale_log2 = 1.4426950408889634f / sqrtf(static_cast<float>(head_dim));
    p.lse = lse;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    switch (head_dim) {
        case 32: return launch_attention<32>(p, B, short_kv, st);
        case 40: return launch_attention<40>(p, B, short_kv, st);
        case 64: return launch_attention<64>(p, B, short_kv, st);
        case 80: return launch_attention<80>(p, B, short_kv, st);
        case 160: return launch_attention<160>(p, B, short_kv, st);
        default:
            set_error("head_dim %d is not instantiated (supported: 32, 40, 64, 80, 160)", head_dim);
            return MFB_EINVAL;
    }
}

extern "C" int mfb_attention(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, void* out, int ldo, int B, int heads,
                             int head_dim, int Tq, int Tk, void* stream) {
    MFB_RECORD(mfb_attention(q, ldq, k, ldk, v, ldv, out, ldo, B, heads, head_dim, Tq, Tk, stream));
    return attention_impl(q, ldq, k, ldk, v, ldv, out, ldo, B, heads, head_dim, Tq, Tk, nullptr, stream);
}

// Forward for training: same kernel, and the per-row log-sum-exp (log2 domain, scale folded in) the backward needs.
extern "C" int mfb_attention_lse(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, void* out, int ldo, int B,
                                 int heads, int head_dim, int Tq, int Tk, float* lse, void* stream) {
    MFB_REQUIRE(lse, "null lse");
    return attention_impl(q, ldq, k, ldk, v, ldv, out, ldo, B, heads, head_dim, Tq, Tk, lse, stream);
}
