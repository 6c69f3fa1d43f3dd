// Backward of scaled-dot-product attention on tcgen05 — the data-gradient chain of the FROZEN UNet in the BrushNet fine-tune step
// (BASELINE config 4): autograd of F.scaled_dot_product_attention in AttnProcessor2_0.__call__
// (S/models/attention_processor.py:1266-1268; no mask, non-causal) under accelerator.backward (E/train_brushnet_mirror.py:1459).
//
// Flash-style, nothing of size Tq x Tk is ever stored, and every output element has ONE writer (deterministic, no atomics):
//   forward (attn.cu) also emits   L_i = log2 sum_j 2^(s_ij)            s_ij = q_i.k_j * d^-1/2 * log2(e)
//   kernel A, CTA = 128 queries, loop over 64-key tiles:                 D_i = dO_i . O_i
//       S = Q K^T, dP = dO V^T   (tcgen05, TMEM)    P = 2^(s - L),  dS = P (dP - D) d^-1/2  -> bf16 smem   dQ += dS K
//   kernel B, CTA = 128 keys, loop over 64-query tiles (the transposed problem, rows = keys):
//       S^T = K Q^T, dP^T = V dO^T                  P^T, dS^T as above with L, D per COLUMN           dV += P^T dO,  dK += dS^T Q
// Cross attention to the (non-trainable) text context only needs kernel A.
// Operand layouts follow attn.cu: 64-column TMA boxes at column h*d land as 128B-swizzled K-major tiles; the SAME smem tile
// is also read as the MN-major B operand of the second product (K in dQ += dS K, dO in dV += P^T dO, Q in dK += dS^T Q), so no
// transposed copy of anything exists.  Head dims that are not multiples of 16 (d = 40): the pad columns [d, dpad) that ride along
// from the neighbouring head are zeroed in smem on ONE operand of every product whose reduction runs over d.
#include <math.h>
#include <string.h>

#include "common.h"
#include "ptx.cuh"

namespace mfb {

template <int D>
struct BwdCfg {
    static constexpr int DPAD = (D + 15) / 16 * 16;
    static constexpr int NKB = (D + 63) / 64;               // 64-column boxes per row tile
    static constexpr int STAGES = D > 80 ? 1 : 2;           // ring depth of the streamed tiles (smem-limited for d = 160)
    static constexpr int T128_BYTES = NKB * 128 * 128;      // 128-row tile
    static constexpr int T64_BYTES = NKB * 64 * 128;        // 64-row tile
    static constexpr int X_BYTES = 128 * 128;               // 128 x 64 bf16 (P / dS)
    static constexpr int CTAS = D <= 40 ? 2 : 1;
    // kernel A: Q, dO resident (128 rows), K / V streamed (64 rows), dS
    static constexpr int A_SMEM = 2 * T128_BYTES + STAGES * 2 * T64_BYTES + X_BYTES + 1024 + 256;
    static constexpr int A_TMEM_USED = 128 + DPAD;          // S | dP | dQ
    static constexpr int A_TMEM = A_TMEM_USED <= 256 ? 256 : 512;
    // kernel B: K, V resident (128 rows), Q / dO streamed (64 rows), P^T, dS^T, per-column L and D of the current query tile
    static constexpr int B_SMEM = 2 * T128_BYTES + STAGES * 2 * T64_BYTES + 2 * X_BYTES + STAGES * 2 * 64 * 4 + 1024 + 256;
    static constexpr int B_TMEM_USED = 128 + 2 * DPAD;      // S^T | dP^T | dV | dK
    static constexpr int B_TMEM = B_TMEM_USED <= 256 ? 256 : 512;
    static_assert(A_SMEM <= 227 * 1024 && B_SMEM <= 227 * 1024, "shared memory budget");
    static_assert(A_TMEM * CTAS <= 512 && B_TMEM * CTAS <= 512, "TMEM over-subscribed");
};

struct AttBwdParams {
    CUtensorMap tmRow128a, tmRow128b;   // the two RESIDENT tensors (A: Q, dO; B: K, V), box {64, 128, 1}
    CUtensorMap tmRow64a, tmRow64b;     // the two STREAMED tensors (A: K, V; B: Q, dO), box {64, 64, 1}
    const __nv_bfloat16* o;             // A: attention output O [B, Tq, ldo] (for D = dO . O)
    const __nv_bfloat16* d_o;           // A: dO, same layout
    int ldo, lddo;
    float* lse;                         // [B, heads, Tq] log2-domain log-sum-exp from the forward (read)
    float* dvec;                        // [B, heads, Tq] D (A writes, B reads)
    __nv_bfloat16* out0;                // A: dQ; B: dK
    __nv_bfloat16* out1;                // B: dV
    int ld0, ld1;
    int Tq, Tk, heads;
    float scale, scale_log2;
};

__device__ __forceinline__ float ex2_fast(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// zero the 16-byte chunk holding columns [D, D+8) of row r of a swizzled row tile (the only pad chunk: D % 8 == 0, DPAD - D = 8)
template <int D, int ROWS>
__device__ __forceinline__ void zero_pad_chunk(uint8_t* tile, int r) {
    constexpr int kb = D / 64, ch = (D % 64) / 8;
    *reinterpret_cast<uint4*>(tile + kb * ROWS * 128 + r * 128 + ((ch ^ (r & 7)) << 4)) = make_uint4(0, 0, 0, 0);
}

// bf16 store of 32 consecutive columns [c0, c0+32) of row r into a 128-row x 64-column K-major swizzled tile
__device__ __forceinline__ void store_row32(uint8_t* tile_row, int r, int c0, const float (&v)[32]) {
#pragma unroll
    for (int i8 = 0; i8 < 4; ++i8) {
        uint4 pk;
        pk.x = pack_bf16x2(v[i8 * 8 + 0], v[i8 * 8 + 1]);
        pk.y = pack_bf16x2(v[i8 * 8 + 2], v[i8 * 8 + 3]);
        pk.z = pack_bf16x2(v[i8 * 8 + 4], v[i8 * 8 + 5]);
        pk.w = pack_bf16x2(v[i8 * 8 + 6], v[i8 * 8 + 7]);
        const int ch = (c0 >> 3) + i8;
        *reinterpret_cast<uint4*>(tile_row + ((ch ^ (r & 7)) << 4)) = pk;
    }
}

// ================================================================================================ kernel A: dQ (and D)
template <int D>
__global__ void __launch_bounds__(192, BwdCfg<D>::CTAS) attn_bwd_dq_kernel(const __grid_constant__ AttBwdParams p) {
    using Cfg = BwdCfg<D>;
    constexpr int DPAD = Cfg::DPAD, NKB = Cfg::NKB, STAGES = Cfg::STAGES, BKV = 64;
    constexpr int TMA_WARP = 4, MMA_WARP = 5;
    constexpr bool kZeroPad = (D % 16) != 0;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_u32 = smem_u32(smem_raw);
    const uint32_t base = (raw_u32 + 1023u) & ~1023u;
    const uint32_t q_smem = base, do_smem = q_smem + Cfg::T128_BYTES;
    const uint32_t kv_smem = do_smem + Cfg::T128_BYTES;
    const uint32_t ds_smem = kv_smem + STAGES * 2 * Cfg::T64_BYTES;
    const uint32_t bar = ds_smem + Cfg::X_BYTES;
    auto bar_at = [&](int i) { return bar + 8u * i; };
    const uint32_t q_full = bar_at(0), q_ready = bar_at(1), s_full = bar_at(2), s_free = bar_at(3), p_full = bar_at(4), pv_done = bar_at(5);
    auto k_full = [&](int s) { return bar_at(6 + s); };
    auto k_empty = [&](int s) { return bar_at(6 + STAGES + s); };
    auto v_full = [&](int s) { return bar_at(6 + 2 * STAGES + s); };
    auto v_empty = [&](int s) { return bar_at(6 + 3 * STAGES + s); };
    const uint32_t tmem_slot = bar_at(6 + 4 * STAGES);
    uint8_t* gen_base = smem_raw + (base - raw_u32);

    const int warp = __shfl_sync(0xffffffffu, int(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;
    const int ntiles = (p.Tk + BKV - 1) / BKV;

    if (warp == TMA_WARP && lane == 0) {
        prefetch_tmap(&p.tmRow128a); prefetch_tmap(&p.tmRow128b); prefetch_tmap(&p.tmRow64a); prefetch_tmap(&p.tmRow64b);
        mbar_init(q_full, 1);
        mbar_init(q_ready, 128);
        mbar_init(s_full, 1);
        mbar_init(s_free, 4);
        mbar_init(p_full, 128);
        mbar_init(pv_done, 1);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(k_full(s), 1); mbar_init(k_empty(s), 1);
            mbar_init(v_full(s), 1); mbar_init(v_empty(s), 1);
        }
        fence_barrier_init();
    }
    if (warp == MMA_WARP) {
        tmem_alloc(tmem_slot, Cfg::A_TMEM);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<uint32_t*>(gen_base + (tmem_slot - base));
    const uint32_t tS = tmem_base, tDP = tmem_base + 64, tDQ = tmem_base + 128;
    pdl_trigger();
    pdl_wait();

    if (warp == TMA_WARP) {
        if (elect_one()) {
            mbar_expect_tx(q_full, 2 * Cfg::T128_BYTES);
            for (int kb = 0; kb < NKB; ++kb) {
                tma_load_3d(q_smem + kb * 128 * 128, &p.tmRow128a, q_full, h * D + kb * 64, q0, b);
                tma_load_3d(do_smem + kb * 128 * 128, &p.tmRow128b, q_full, h * D + kb * 64, q0, b);
            }
            for (int j = 0; j < ntiles; ++j) {
                const int s = j % STAGES;
                const uint32_t ph = ((j / STAGES) & 1) ^ 1;
                const uint32_t kd = kv_smem + s * 2 * Cfg::T64_BYTES, vd = kd + Cfg::T64_BYTES;
                mbar_wait_relaxed(k_empty(s), ph);
                mbar_expect_tx(k_full(s), Cfg::T64_BYTES);
                for (int kb = 0; kb < NKB; ++kb) tma_load_3d(kd + kb * BKV * 128, &p.tmRow64a, k_full(s), h * D + kb * 64, j * BKV, b);
                mbar_wait_relaxed(v_empty(s), ph);
                mbar_expect_tx(v_full(s), Cfg::T64_BYTES);
                for (int kb = 0; kb < NKB; ++kb) tma_load_3d(vd + kb * BKV * 128, &p.tmRow64b, v_full(s), h * D + kb * 64, j * BKV, b);
            }
        }
        __syncwarp();
    } else if (warp == MMA_WARP) {
        if (elect_one()) {
            constexpr uint32_t idesc_s = make_idesc_bf16(128, BKV);
            constexpr uint32_t idesc_dq = make_idesc_bf16(128, DPAD, 0, 1);       // B = K tile, MN-major
            // S(j) = Q K(j)^T and dP(j) = dO V(j)^T
            auto issue_sdp = [&](int j) {
                const int s = j % STAGES;
                const uint32_t ph = (j / STAGES) & 1;
                const uint32_t kd = kv_smem + s * 2 * Cfg::T64_BYTES, vd = kd + Cfg::T64_BYTES;
                mbar_wait_relaxed(k_full(s), ph);
                mbar_wait_relaxed(v_full(s), ph);
                if (j > 0) mbar_wait_relaxed(s_free, (j - 1) & 1);       // S / dP of tile j-1 are in registers
                tc_fence_after();
#pragma unroll
                for (int ks = 0; ks < DPAD / 16; ++ks) {
                    const uint64_t aq = make_desc_k_sw128(q_smem + (ks / 4) * 128 * 128) + uint64_t(2 * (ks % 4));
                    const uint64_t bk = make_desc_k_sw128(kd + (ks / 4) * BKV * 128) + uint64_t(2 * (ks % 4));
                    umma_bf16(tS, aq, bk, idesc_s, ks != 0);
                }
#pragma unroll
                for (int ks = 0; ks < DPAD / 16; ++ks) {
                    const uint64_t ad = make_desc_k_sw128(do_smem + (ks / 4) * 128 * 128) + uint64_t(2 * (ks % 4));
                    const uint64_t bv = make_desc_k_sw128(vd + (ks / 4) * BKV * 128) + uint64_t(2 * (ks % 4));
                    umma_bf16(tDP, ad, bv, idesc_s, ks != 0);
                }
                umma_commit(s_full);
                umma_commit(v_empty(s));
            };
            mbar_wait(q_full, 0);
            mbar_wait(q_ready, 0);
            if (STAGES > 1) issue_sdp(0);
            for (int j = 0; j < ntiles; ++j) {
                const int s = j % STAGES;
                if (STAGES > 1) {
                    if (j + 1 < ntiles) issue_sdp(j + 1);     // runs under the exponentials of tile j
                } else {
                    issue_sdp(j);                             // one stage: K(j) must stay until dQ(j) has read it
                }
                const uint32_t kd = kv_smem + s * 2 * Cfg::T64_BYTES;
                mbar_wait_relaxed(p_full, j & 1);             // dS(j) is in smem
                tc_fence_after();
#pragma unroll
                for (int ks = 0; ks < BKV / 16; ++ks) {
                    const uint64_t ad = make_desc_k_sw128(ds_smem) + uint64_t(2 * ks);
                    const uint64_t bd = make_desc_mn_sw128(kd + ks * 16 * 128, BKV * 128);
                    umma_bf16(tDQ, ad, bd, idesc_dq, (j | ks) != 0);
                }
                umma_commit(pv_done);
                umma_commit(k_empty(s));
            }
        }
        __syncwarp();
    } else {
        // ===== compute warps: thread = query row =====
        const int r = warp * 32 + lane;
        const uint32_t lane_off = uint32_t(warp * 32) << 16;
        const int qrow = q0 + r;
        const bool valid = qrow < p.Tq;
        // D_r = dO_r . O_r from global (one row of d elements each), L_r from the forward
        float Dr = 0.f, Lr = 0.f;
        if (valid) {
            const __nv_bfloat16* orow = p.o + (static_cast<size_t>(b) * p.Tq + qrow) * p.ldo + h * D;
            const __nv_bfloat16* drow = p.d_o + (static_cast<size_t>(b) * p.Tq + qrow) * p.lddo + h * D;
#pragma unroll
            for (int c = 0; c < D; c += 8) {
                const uint4 uo = __ldg(reinterpret_cast<const uint4*>(orow + c)), ud = __ldg(reinterpret_cast<const uint4*>(drow + c));
                float2 a, g;
                a = unpack_bf16x2(uo.x); g = unpack_bf16x2(ud.x); Dr = fmaf(a.x, g.x, Dr); Dr = fmaf(a.y, g.y, Dr);
                a = unpack_bf16x2(uo.y); g = unpack_bf16x2(ud.y); Dr = fmaf(a.x, g.x, Dr); Dr = fmaf(a.y, g.y, Dr);
                a = unpack_bf16x2(uo.z); g = unpack_bf16x2(ud.z); Dr = fmaf(a.x, g.x, Dr); Dr = fmaf(a.y, g.y, Dr);
                a = unpack_bf16x2(uo.w); g = unpack_bf16x2(ud.w); Dr = fmaf(a.x, g.x, Dr); Dr = fmaf(a.y, g.y, Dr);
            }
            const size_t si = (static_cast<size_t>(b) * p.heads + h) * p.Tq + qrow;
            Lr = __ldg(p.lse + si);
            p.dvec[si] = Dr;
        }
        mbar_wait(q_full, 0);
        if (kZeroPad) {
            zero_pad_chunk<D, 128>(gen_base + (q_smem - base), r);
            zero_pad_chunk<D, 128>(gen_base + (do_smem - base), r);
            fence_proxy_async_smem();
        }
        mbar_arrive(q_ready);
        uint8_t* dsrow = gen_base + (ds_smem - base) + r * 128;
        for (int j = 0; j < ntiles; ++j) {
            mbar_wait_relaxed(s_full, j & 1);
            tc_fence_after();
            const int kbase = j * BKV;
            const bool tail = kbase + BKV > p.Tk;
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                uint32_t sv[32], dp[32];
                tmem_ld32(tS + lane_off + c * 32, sv);
                tmem_ld32(tDP + lane_off + c * 32, dp);
                tmem_wait_ld();
                if (c == 1) {                               // both halves are in registers: the MMA warp may start tile j+1
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(s_free);
                }
                float ds[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const float pr = ex2_fast(fmaf(__uint_as_float(sv[i]), p.scale_log2, -Lr));
                    ds[i] = pr * (__uint_as_float(dp[i]) - Dr) * p.scale;
                }
                if (tail) {
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        if (kbase + c * 32 + i >= p.Tk) ds[i] = 0.f;
                }
                if (c == 0 && j > 0) {
                    mbar_wait_relaxed(pv_done, (j - 1) & 1);      // dQ += dS(j-1) K(j-1) has read the dS buffer
                }
                store_row32(dsrow, r, c * 32, ds);
            }
            fence_proxy_async_smem();
            tc_fence_before();
            mbar_arrive(p_full);
        }
        mbar_wait(pv_done, (ntiles - 1) & 1);
        tc_fence_after();
        __nv_bfloat16* dst = p.out0 + (static_cast<size_t>(b) * p.Tq + qrow) * p.ld0 + h * D;
#pragma unroll
        for (int c = 0; c < DPAD / 16; ++c) {
            uint32_t ov[16];
            tmem_ld16(tDQ + lane_off + c * 16, ov);
            tmem_wait_ld();
            if (valid) {
#pragma unroll
                for (int i8 = 0; i8 < 2; ++i8) {
                    if (c * 16 + i8 * 8 < D) {
                        uint4 pk;
                        pk.x = pack_bf16x2(__uint_as_float(ov[i8 * 8 + 0]), __uint_as_float(ov[i8 * 8 + 1]));
                        pk.y = pack_bf16x2(__uint_as_float(ov[i8 * 8 + 2]), __uint_as_float(ov[i8 * 8 + 3]));
                        pk.z = pack_bf16x2(__uint_as_float(ov[i8 * 8 + 4]), __uint_as_float(ov[i8 * 8 + 5]));
                        pk.w = pack_bf16x2(__uint_as_float(ov[i8 * 8 + 6]), __uint_as_float(ov[i8 * 8 + 7]));
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
        tmem_dealloc(tmem_base, Cfg::A_TMEM);
    }
}

// ================================================================================================ kernel B: dK, dV
template <int D>
__global__ void __launch_bounds__(192, BwdCfg<D>::CTAS) attn_bwd_dkv_kernel(const __grid_constant__ AttBwdParams p) {
    using Cfg = BwdCfg<D>;
    constexpr int DPAD = Cfg::DPAD, NKB = Cfg::NKB, STAGES = Cfg::STAGES, BQT = 64;
    constexpr int TMA_WARP = 4, MMA_WARP = 5;
    constexpr bool kZeroPad = (D % 16) != 0;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_u32 = smem_u32(smem_raw);
    const uint32_t base = (raw_u32 + 1023u) & ~1023u;
    const uint32_t k_smem = base, v_smem = k_smem + Cfg::T128_BYTES;
    const uint32_t qd_smem = v_smem + Cfg::T128_BYTES;                 // ring of {Q tile, dO tile}
    const uint32_t pt_smem = qd_smem + STAGES * 2 * Cfg::T64_BYTES;    // P^T
    const uint32_t dst_smem = pt_smem + Cfg::X_BYTES;                  // dS^T
    const uint32_t ld_smem = dst_smem + Cfg::X_BYTES;                  // [STAGES][2][64] floats: L | D of the query tile
    const uint32_t bar = ld_smem + STAGES * 2 * 64 * 4;
    auto bar_at = [&](int i) { return bar + 8u * i; };
    const uint32_t kv_full = bar_at(0), kv_ready = bar_at(1), s_full = bar_at(2), s_free = bar_at(3), p_full = bar_at(4), pv_done = bar_at(5);
    auto q_full = [&](int s) { return bar_at(6 + s); };
    auto q_empty = [&](int s) { return bar_at(6 + STAGES + s); };
    const uint32_t tmem_slot = bar_at(6 + 2 * STAGES);
    uint8_t* gen_base = smem_raw + (base - raw_u32);

    const int warp = __shfl_sync(0xffffffffu, int(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const int k0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;
    const int ntiles = (p.Tq + BQT - 1) / BQT;

    if (warp == TMA_WARP && lane == 0) {
        prefetch_tmap(&p.tmRow128a); prefetch_tmap(&p.tmRow128b); prefetch_tmap(&p.tmRow64a); prefetch_tmap(&p.tmRow64b);
        mbar_init(kv_full, 1);
        mbar_init(kv_ready, 128);
        mbar_init(s_full, 1);
        mbar_init(s_free, 4);
        mbar_init(p_full, 128);
        mbar_init(pv_done, 1);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(q_full(s), 1);
            mbar_init(q_empty(s), 1);
        }
        fence_barrier_init();
    }
    if (warp == MMA_WARP) {
        tmem_alloc(tmem_slot, Cfg::B_TMEM);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<uint32_t*>(gen_base + (tmem_slot - base));
    const uint32_t tS = tmem_base, tDP = tmem_base + 64, tDV = tmem_base + 128, tDK = tmem_base + 128 + DPAD;
    pdl_trigger();
    pdl_wait();

    if (warp == TMA_WARP) {
        if (elect_one()) {
            mbar_expect_tx(kv_full, 2 * Cfg::T128_BYTES);
            for (int kb = 0; kb < NKB; ++kb) {
                tma_load_3d(k_smem + kb * 128 * 128, &p.tmRow128a, kv_full, h * D + kb * 64, k0, b);
                tma_load_3d(v_smem + kb * 128 * 128, &p.tmRow128b, kv_full, h * D + kb * 64, k0, b);
            }
            for (int i = 0; i < ntiles; ++i) {
                const int s = i % STAGES;
                const uint32_t qd = qd_smem + s * 2 * Cfg::T64_BYTES, dd = qd + Cfg::T64_BYTES;
                mbar_wait_relaxed(q_empty(s), ((i / STAGES) & 1) ^ 1);
                mbar_expect_tx(q_full(s), 2 * Cfg::T64_BYTES);
                for (int kb = 0; kb < NKB; ++kb) {
                    tma_load_3d(qd + kb * BQT * 128, &p.tmRow64a, q_full(s), h * D + kb * 64, i * BQT, b);
                    tma_load_3d(dd + kb * BQT * 128, &p.tmRow64b, q_full(s), h * D + kb * 64, i * BQT, b);
                }
            }
        }
        __syncwarp();
    } else if (warp == MMA_WARP) {
        if (elect_one()) {
            constexpr uint32_t idesc_s = make_idesc_bf16(128, BQT);
            constexpr uint32_t idesc_o = make_idesc_bf16(128, DPAD, 0, 1);        // B = dO / Q tile, MN-major
            auto issue_sdp = [&](int i) {
                const int s = i % STAGES;
                const uint32_t qd = qd_smem + s * 2 * Cfg::T64_BYTES, dd = qd + Cfg::T64_BYTES;
                mbar_wait_relaxed(q_full(s), (i / STAGES) & 1);
                if (i > 0) mbar_wait_relaxed(s_free, (i - 1) & 1);
                tc_fence_after();
#pragma unroll
                for (int ks = 0; ks < DPAD / 16; ++ks) {
                    const uint64_t ak = make_desc_k_sw128(k_smem + (ks / 4) * 128 * 128) + uint64_t(2 * (ks % 4));
                    const uint64_t bq = make_desc_k_sw128(qd + (ks / 4) * BQT * 128) + uint64_t(2 * (ks % 4));
                    umma_bf16(tS, ak, bq, idesc_s, ks != 0);
                }
#pragma unroll
                for (int ks = 0; ks < DPAD / 16; ++ks) {
                    const uint64_t av = make_desc_k_sw128(v_smem + (ks / 4) * 128 * 128) + uint64_t(2 * (ks % 4));
                    const uint64_t bd = make_desc_k_sw128(dd + (ks / 4) * BQT * 128) + uint64_t(2 * (ks % 4));
                    umma_bf16(tDP, av, bd, idesc_s, ks != 0);
                }
                umma_commit(s_full);
            };
            mbar_wait(kv_full, 0);
            mbar_wait(kv_ready, 0);
            if (STAGES > 1) issue_sdp(0);
            for (int i = 0; i < ntiles; ++i) {
                const int s = i % STAGES;
                if (STAGES > 1) {
                    if (i + 1 < ntiles) issue_sdp(i + 1);
                } else {
                    issue_sdp(i);
                }
                const uint32_t qd = qd_smem + s * 2 * Cfg::T64_BYTES, dd = qd + Cfg::T64_BYTES;
                mbar_wait_relaxed(p_full, i & 1);             // P^T(i), dS^T(i) are in smem
                tc_fence_after();
#pragma unroll
                for (int ks = 0; ks < BQT / 16; ++ks) {
                    const uint64_t ap = make_desc_k_sw128(pt_smem) + uint64_t(2 * ks);
                    const uint64_t bo = make_desc_mn_sw128(dd + ks * 16 * 128, BQT * 128);
                    umma_bf16(tDV, ap, bo, idesc_o, (i | ks) != 0);
                }
#pragma unroll
                for (int ks = 0; ks < BQT / 16; ++ks) {
                    const uint64_t as_ = make_desc_k_sw128(dst_smem) + uint64_t(2 * ks);
                    const uint64_t bq = make_desc_mn_sw128(qd + ks * 16 * 128, BQT * 128);
                    umma_bf16(tDK, as_, bq, idesc_o, (i | ks) != 0);
                }
                umma_commit(pv_done);
                umma_commit(q_empty(s));
            }
        }
        __syncwarp();
    } else {
        // ===== compute warps: thread = key row =====
        const int r = warp * 32 + lane;
        const uint32_t lane_off = uint32_t(warp * 32) << 16;
        const int krow = k0 + r;
        const bool valid = krow < p.Tk;
        mbar_wait(kv_full, 0);
        if (kZeroPad) {
            zero_pad_chunk<D, 128>(gen_base + (k_smem - base), r);
            zero_pad_chunk<D, 128>(gen_base + (v_smem - base), r);
            fence_proxy_async_smem();
        }
        mbar_arrive(kv_ready);
        uint8_t* ptrow = gen_base + (pt_smem - base) + r * 128;
        uint8_t* dsrow = gen_base + (dst_smem - base) + r * 128;
        float* ld_gen = reinterpret_cast<float*>(gen_base + (ld_smem - base));
        const size_t sbase = (static_cast<size_t>(b) * p.heads + h) * p.Tq;
        // per-column L and D of a query tile (threads 0..63: L, 64..127: D); columns past Tq get L = +inf (P = 0).  The value of
        // tile i+1 is fetched while tile i is being processed, so the global-load latency is off the per-tile critical path.
        auto fetch_ld = [&](int i) {
            const int qi = i * BQT + (r & 63);
            float v = r < 64 ? INFINITY : 0.f;
            if (i < ntiles && qi < p.Tq) v = __ldg((r < 64 ? p.lse : p.dvec) + sbase + qi);
            return v;
        };
        float ld_next = fetch_ld(0);
        for (int i = 0; i < ntiles; ++i) {
            float* Ls = ld_gen + (i % STAGES) * 128;
            if (STAGES == 1 && i > 0) named_bar_sync(2, 128);          // one buffer: everybody has finished reading tile i-1's values
            Ls[r] = ld_next;
            named_bar_sync(1, 128);
            ld_next = fetch_ld(i + 1);
            mbar_wait_relaxed(s_full, i & 1);
            tc_fence_after();
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                uint32_t sv[32], dp[32];
                tmem_ld32(tS + lane_off + c * 32, sv);
                tmem_ld32(tDP + lane_off + c * 32, dp);
                tmem_wait_ld();
                if (c == 1) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(s_free);
                }
                float pt[32], ds[32];
#pragma unroll
                for (int i4 = 0; i4 < 8; ++i4) {
                    const float4 L4 = *reinterpret_cast<const float4*>(Ls + c * 32 + i4 * 4);
                    const float4 D4 = *reinterpret_cast<const float4*>(Ls + 64 + c * 32 + i4 * 4);
                    const float Lv[4] = {L4.x, L4.y, L4.z, L4.w}, Dv[4] = {D4.x, D4.y, D4.z, D4.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int ii = i4 * 4 + e;
                        pt[ii] = ex2_fast(fmaf(__uint_as_float(sv[ii]), p.scale_log2, -Lv[e]));
                        ds[ii] = pt[ii] * (__uint_as_float(dp[ii]) - Dv[e]) * p.scale;
                    }
                }
                if (c == 0 && i > 0) mbar_wait_relaxed(pv_done, (i - 1) & 1);      // the MMAs of tile i-1 have read both buffers
                store_row32(ptrow, r, c * 32, pt);
                store_row32(dsrow, r, c * 32, ds);
            }
            fence_proxy_async_smem();
            tc_fence_before();
            mbar_arrive(p_full);
        }
        mbar_wait(pv_done, (ntiles - 1) & 1);
        tc_fence_after();
        __nv_bfloat16* dkp = p.out0 + (static_cast<size_t>(b) * p.Tk + krow) * p.ld0 + h * D;
        __nv_bfloat16* dvp = p.out1 + (static_cast<size_t>(b) * p.Tk + krow) * p.ld1 + h * D;
#pragma unroll
        for (int which = 0; which < 2; ++which) {
            const uint32_t tsrc = (which == 0 ? tDV : tDK) + lane_off;
            __nv_bfloat16* dst = which == 0 ? dvp : dkp;
#pragma unroll
            for (int c = 0; c < DPAD / 16; ++c) {
                uint32_t ov[16];
                tmem_ld16(tsrc + c * 16, ov);
                tmem_wait_ld();
                if (valid) {
#pragma unroll
                    for (int i8 = 0; i8 < 2; ++i8) {
                        if (c * 16 + i8 * 8 < D) {
                            uint4 pk;
                            pk.x = pack_bf16x2(__uint_as_float(ov[i8 * 8 + 0]), __uint_as_float(ov[i8 * 8 + 1]));
                            pk.y = pack_bf16x2(__uint_as_float(ov[i8 * 8 + 2]), __uint_as_float(ov[i8 * 8 + 3]));
                            pk.z = pack_bf16x2(__uint_as_float(ov[i8 * 8 + 4]), __uint_as_float(ov[i8 * 8 + 5]));
                            pk.w = pack_bf16x2(__uint_as_float(ov[i8 * 8 + 6]), __uint_as_float(ov[i8 * 8 + 7]));
                            *reinterpret_cast<uint4*>(dst + c * 16 + i8 * 8) = pk;
                        }
                    }
                }
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == MMA_WARP) {
        tc_fence_after();
        tmem_dealloc(tmem_base, Cfg::B_TMEM);
    }
}

static int encode_rows(CUtensorMap* tm, const void* base, int ld, int width, int T, int B, int box_rows) {
    const uint64_t dims[3] = {uint64_t(width), uint64_t(T), uint64_t(B)};
    const uint64_t str[2] = {uint64_t(ld) * 2, uint64_t(T) * ld * 2};
    const uint32_t box[3] = {64, uint32_t(box_rows), 1};
    return encode_tmap_bf16(tm, base, 3, dims, str, box, 128);
}

template <int D>
static int launch_bwd(AttBwdParams& pa, AttBwdParams& pb, bool need_kv, int B, cudaStream_t st) {
    using Cfg = BwdCfg<D>;
    static bool configured = false;
    if (!configured) {
        MFB_CUDA_OK(cudaFuncSetAttribute(attn_bwd_dq_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::A_SMEM));
        MFB_CUDA_OK(cudaFuncSetAttribute(attn_bwd_dkv_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::B_SMEM));
        configured = true;
    }
    MFB_CUDA_OK(launch_k(attn_bwd_dq_kernel<D>, dim3((pa.Tq + 127) / 128, pa.heads, B), dim3(192), Cfg::A_SMEM, st, 1, pa));
    if (need_kv)
        MFB_CUDA_OK(launch_k(attn_bwd_dkv_kernel<D>, dim3((pb.Tk + 127) / 128, pb.heads, B), dim3(192), Cfg::B_SMEM, st, 1, pb));
    return MFB_OK;
}

}  // namespace mfb

using namespace mfb;

extern "C" int mfb_attention_bwd(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, const void* o, int ldo,
                                 const void* d_o, int lddo, const float* lse, float* dvec, void* dq, int lddq, void* dk, int lddk,
                                 void* dv, int lddv, int B, int heads, int head_dim, int Tq, int Tk, void* stream) {
    MFB_REQUIRE(q && k && v && o && d_o && lse && dvec && dq, "null pointer");
    MFB_REQUIRE((dk == nullptr) == (dv == nullptr), "dk and dv go together (both NULL: cross attention to a frozen context)");
    const int width = heads * head_dim;
    MFB_REQUIRE(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 8 == 0 && lddo % 8 == 0 && lddq % 8 == 0 && (!dk || (lddk % 8 == 0 && lddv % 8 == 0)),
                "leading dimensions must be multiples of 8");
    MFB_REQUIRE(ldq >= width && ldk >= width && ldv >= width && ldo >= width && lddo >= width && lddq >= width, "leading dimensions must cover heads * head_dim");
    MFB_REQUIRE(Tq > 0 && Tk > 0 && head_dim % 8 == 0, "bad geometry");
    AttBwdParams pa, pb;
    memset(&pa, 0, sizeof(pa));
    int rc;
    if ((rc = encode_rows(&pa.tmRow128a, q, ldq, width, Tq, B, 128))) return rc;
    if ((rc = encode_rows(&pa.tmRow128b, d_o, lddo, width, Tq, B, 128))) return rc;
    if ((rc = encode_rows(&pa.tmRow64a, k, ldk, width, Tk, B, 64))) return rc;
    if ((rc = encode_rows(&pa.tmRow64b, v, ldv, width, Tk, B, 64))) return rc;
    pa.o = static_cast<const __nv_bfloat16*>(o);
    pa.d_o = static_cast<const __nv_bfloat16*>(d_o);
    pa.ldo = ldo; pa.lddo = lddo;
    pa.lse = const_cast<float*>(lse);
    pa.dvec = dvec;
    pa.out0 = static_cast<__nv_bfloat16*>(dq);
    pa.ld0 = lddq;
    pa.Tq = Tq; pa.Tk = Tk; pa.heads = heads;
    pa.scale = 1.0f / sqrtf(static_cast<float>(head_dim));
    pa.scale_log2 = 1.4426950408889634f * pa.scale;
    pb = pa;
    if (dk) {
        if ((rc = encode_rows(&pb.tmRow128a, k, ldk, width, Tk, B, 128))) return rc;
        if ((rc = encode_rows(&pb.tmRow128b, v, ldv, width, Tk, B, 128))) return rc;
        if ((rc = encode_rows(&pb.tmRow64a, q, ldq, width, Tq, B, 64))) return rc;
        if ((rc = encode_rows(&pb.tmRow64b, d_o, lddo, width, Tq, B, 64))) return rc;
        pb.out0 = static_cast<__nv_bfloat16*>(dk);
        pb.out1 = static_cast<__nv_bfloat16*>(dv);
        pb.ld0 = lddk; pb.ld1 = lddv;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    switch (head_dim) {
        case 32: return launch_bwd<32>(pa, pb, dk != nullptr, B, st);
        case 40: return launch_bwd<40>(pa, pb, dk != nullptr, B, st);
        case 64: return launch_bwd<64>(pa, pb, dk != nullptr, B, st);
        case 80: return launch_bwd<80>(pa, pb, dk != nullptr, B, st);
        case 160: return launch_bwd<160>(pa, pb, dk != nullptr, B, st);
        default:
            set_error("head_dim %d is not instantiated (supported: 32, 40, 64, 80, 160)", head_dim);
            return MFB_EINVAL;
    }
}
