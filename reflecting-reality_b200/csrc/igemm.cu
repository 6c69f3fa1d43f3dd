// Implicit-GEMM convolution / linear layer on tcgen05 tensor cores.
//
// Replaces every F.conv2d (3x3 s1/s2, 1x1) and F.linear the reference issues on the hot path
// (S/models/lora.py:363-377,445-451 called from S/models/resnet.py:367,396-401, downsampling.py:146-152,
// upsampling.py:179-184, transformer_2d.py:340-344,417-421, attention_processor.py:1246-1274, attention.py:668-675,
// brushnet.py:832-834,851,891-893).
//
//   D[M = B*Ho*Wo pixels, N = Cout] = sum over K-segments  A_seg[M, 64*cblocks] * W[N, kofs_seg ...]^T
//
// Activations are NHWC bf16.  A K-segment is (tensor map, dh, dw, channel range): the TMA producer fetches the
// 128-pixel x 64-channel A tile for filter tap (dh,dw) as ONE 4-D box {64, tw, th, tn} at coordinates
// (c, w0+dw, h0+dh, n0); out-of-bounds rows/columns are zero-filled by TMA, which is exactly the conv padding.
// A box lands in shared memory as 128 dense 128-byte rows with the 128B swizzle = the canonical K-major UMMA
// operand layout, so no im2col buffer and no register staging exists anywhere.  Extra 1x1 segments let one
// accumulator also absorb the ResnetBlock2D 1x1 shortcut over the (concatenated) block input — the skip concat
// (unet_2d_blocks.py:2586,2728) is just two segments — and stride-2 convs read four parity views of the input.
// Warp roles: warps 0-7 = epilogue, warp 8 = TMA producer, warp 9 = TMEM allocator + single-thread tcgen05.mma issuer
// (epilogue (tcgen05.ld -> +bias +timestep-embedding row bias, xalpha, +residual(s) / GEGLU -> bf16 NHWC store).
// Two CTAs are co-resident per SM (<=113 KB smem, <=256 TMEM columns each), so one CTA's epilogue overlaps the
// other's main loop.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "common.h"
#include <type_traits>
#include "fp32mode.h"
#include "ptx.cuh"

namespace mfb {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int MAX_SEG = 16;
// Epilogue warps per CTA (kernel template parameter EPIW): 8 or 16 = 4 TMEM lane quarters x 2 or 4 column groups.  16 halves
// the per-tile epilogue latency (short-K and 8x8-level launches: 5-11 % faster) at the price of 96 registers per thread
// (long-K convs: 0.5-1.5 % slower), so the plan picks it per launch (profiles/r01o_igemm_double_staging.md).

struct IgemmSeg {
    int map;      // which A tensor map
    int dh, dw;   // spatial offset of this filter tap (in the coordinates of that map)
    int c0;       // first channel inside that source
    int cblocks;  // number of 64-channel K blocks
};

struct IgemmParams {
    CUtensorMap tmA[7];    // stride 1: [0] = conv input, [1..3] = extra 1x1 sources; stride 2: [0..3] = parity views, [4..6] = extras
    CUtensorMap tmB;
    CUtensorMap tmBh;      // weights with a BN/2-row box (pair mode: each CTA fetches half and multicasts it)
    CUtensorMap tmOut;     // output tile store  (box {EPI box cols, tw, th, tn}, 64B / 32B swizzle)
    CUtensorMap tmRes;     // res1 tile load, same geometry
    IgemmSeg seg[MAX_SEG];
    int nseg;
    int M, N;              // GEMM rows (valid output pixels) and columns (rows of the packed weight)
    int Wo, Ho, Bn;        // output geometry
    int tw, th, tn;        // tile box (tw*th*tn == 128)
    int tiles_w, tiles_h;  // tiles per row / column
    int tiles_m, tiles_nn; // total M tiles, N tiles
    const float* bias;     // [N] (packed order) or null
    const float* rowbias;  // [Bn, rowbias_ld] or null (timestep-embedding projection)
    int rowbias_ld;
    const float* alpha;    // device scalar or null (1.0)
    const __nv_bfloat16* res1;  // [M, out_ld] or null
    const __nv_bfloat16* res2;
    __nv_bfloat16* out;    // [M, out_ld]
    int out_ld;
    int geglu;             // BN == 128 only: cols [0,64) value, [64,128) gate -> out[:, nt*64 + j]
    float* stats;          // optional per-(image, slot, channel) {sum, sumsq} of the bf16 OUTPUT (GroupNorm statistics of the
                           // consumer, fused here): [Bn][stats_tiles][out_ld][2]; a slot = one epilogue warp's 32 rows of a tile; null = off
    int stats_tiles, stats_tile_base;   // slots per image; first slot of this launch (sub-pixel phases)
    int stats_rows_img;                 // tn > 1 (a tile spans whole images): rows per image (tw * th, a multiple of 32)
    int o_step, o_py, o_px, o_Hf, o_Wf;   // rows map to output pixel (o_step*oh + o_py, o_step*ow + o_px) of [Bn, o_Hf, o_Wf]
    int stages;            // operand ring depth actually used (<= IgemmCfg::STAGES)
    int nstg;              // staging tiles: 1, or 2 = double-buffered epilogue (short-K launches, see igemm_kernel)
};

template <int BN>
struct IgemmCfg {
    static constexpr int A_BYTES = BM * BK * 2;
    static constexpr int B_BYTES = BN * BK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    // epilogue staging tile: the bf16 output tile (and, before that, the res1 tile) as column blocks of BOXC
    // columns x 128 rows, each block written/read by one TMA box with the 64B (BOXC 32) or 32B (BOXC 16) swizzle
    static constexpr int BOXC = (BN % 32 == 0) ? 32 : 16;
    static constexpr int STG_BYTES = BM * BN * 2;
    // persistent CTA, one per SM: as many operand stages as fit under the 227 KB limit (operand fetch is
    // L2-latency/bandwidth bound, so depth matters more than anything else here)
    static constexpr int MAX_STAGES = 8;
    static constexpr int STAGES = (224 * 1024 - STG_BYTES) / STAGE_BYTES > MAX_STAGES ? MAX_STAGES : (224 * 1024 - STG_BYTES) / STAGE_BYTES;
    // with a second staging tile (double-buffered epilogue) the ring gives up STG_BYTES
    static constexpr int STAGES_2STG = (224 * 1024 - 2 * STG_BYTES) / STAGE_BYTES > MAX_STAGES ? MAX_STAGES : (224 * 1024 - 2 * STG_BYTES) / STAGE_BYTES;
    static constexpr int ACC_COLS = BN <= 64 ? 64 : BN <= 128 ? 128 : 256;  // TMEM columns per accumulator slot
    static constexpr int TMEM_COLS = 2 * ACC_COLS;                           // double-buffered accumulator
    static constexpr int SMEM_BYTES = 224 * 1024 /*ring + staging tile(s)*/ + 1024 /*align slack*/ + 256 /*barriers*/ +
                                      2 * BN * 4 /*column bias of the current / next tile*/;
    static_assert(STAGES >= 2 && STAGES_2STG >= 2 && SMEM_BYTES <= 227 * 1024, "shared memory budget");
};

// byte offset of the 16-byte chunk holding columns [col, col+8) of row r inside the swizzled staging tile
template <int BOXC>
__device__ __forceinline__ uint32_t stg_off(int r, int col) {
    const int blk = col / BOXC;
    const int j = (col % BOXC) >> 3;
    const int sw = BOXC == 32 ? ((r >> 1) & 3) : ((r >> 2) & 1);
    return uint32_t(blk * (BM * BOXC * 2) + r * (BOXC * 2) + ((j ^ sw) << 4));
}

__device__ __forceinline__ void add_bf16x8(float (&f)[8], const uint4& rv) {
    float2 t;
    t = unpack_bf16x2(rv.x); f[0] += t.x; f[1] += t.y;
    t = unpack_bf16x2(rv.y); f[2] += t.x; f[3] += t.y;
    t = unpack_bf16x2(rv.z); f[4] += t.x; f[5] += t.y;
    t = unpack_bf16x2(rv.w); f[6] += t.x; f[7] += t.y;
}
__device__ __forceinline__ void add_f32x8(float (&f)[8], const float* __restrict__ p) {
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(p));
    const float4 b1 = __ldg(reinterpret_cast<const float4*>(p + 4));
    f[0] += b0.x; f[1] += b0.y; f[2] += b0.z; f[3] += b0.w;
    f[4] += b1.x; f[5] += b1.y; f[6] += b1.z; f[7] += b1.w;
}
__device__ __forceinline__ uint4 pack_bf16x8(const float (&f)[8]) {
    uint4 o;
    o.x = pack_bf16x2(f[0], f[1]);
    o.y = pack_bf16x2(f[2], f[3]);
    o.z = pack_bf16x2(f[4], f[5]);
    o.w = pack_bf16x2(f[6], f[7]);
    return o;
}

// Column sums of a 32-row x 16-column register block (one row per lane): a transposing butterfly — at every level a lane keeps
// half of its columns and hands the other half to its partner, so 8 + 4 + 2 + 1 + 1 = 16 shuffles reduce 16 columns over 32 rows
// (a plain per-column xor-reduction would take 80).  On return every lane pair (l, l^1) holds the sum of column l >> 1.
__device__ __forceinline__ float colsum16(const float (&x)[16], int lane) {
    const unsigned full = 0xffffffffu;
    float a8[8], a4[4], a2[2];
    const bool h4 = (lane & 16) != 0, h3 = (lane & 8) != 0, h2 = (lane & 4) != 0, h1 = (lane & 2) != 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) a8[i] = (h4 ? x[8 + i] : x[i]) + __shfl_xor_sync(full, h4 ? x[i] : x[8 + i], 16);
#pragma unroll
    for (int i = 0; i < 4; ++i) a4[i] = (h3 ? a8[4 + i] : a8[i]) + __shfl_xor_sync(full, h3 ? a8[i] : a8[4 + i], 8);
#pragma unroll
    for (int i = 0; i < 2; ++i) a2[i] = (h2 ? a4[2 + i] : a4[i]) + __shfl_xor_sync(full, h2 ? a4[i] : a4[2 + i], 4);
    const float a1 = (h1 ? a2[1] : a2[0]) + __shfl_xor_sync(full, h1 ? a2[0] : a2[1], 2);
    return a1 + __shfl_xor_sync(full, a1, 1);
}

// Walks the tile sequence tile0, tile0 + step, ... of one CTA without per-tile integer divisions: the position is kept
// as mixed-radix digits (N tile, tile column, tile row, image group) and advanced by the pre-decomposed step.  (The
// K=320 projections have only 5 K blocks per tile: a handful of runtime divisions per tile in every thread showed up
// as ~14 % of the epilogue's samples.)
struct TileIter {
    int nt, iw, ih, ig;          // current digits
    int dn, dw, dh, dg;          // step digits
    int rn, rw, rh;              // radices
    __device__ __forceinline__ void init(int tile0, int step, int tiles_n, int tiles_w, int tiles_h) {
        rn = tiles_n; rw = tiles_w; rh = tiles_h;
        nt = tile0 % rn; int m = tile0 / rn;
        iw = m % rw; m /= rw; ih = m % rh; ig = m / rh;
        dn = step % rn; m = step / rn;
        dw = m % rw; m /= rw; dh = m % rh; dg = m / rh;
    }
    __device__ __forceinline__ void next() {
        nt += dn;
        int c = nt >= rn; nt -= c ? rn : 0;
        iw += dw + c;
        c = iw >= rw; iw -= c ? rw : 0;
        ih += dh + c;
        c = ih >= rh; ih -= c ? rh : 0;
        ig += dg + c;
    }
    __device__ __forceinline__ int mt() const { return (ig * rh + ih) * rw + iw; }
};

// Persistent kernel: grid = min(#tiles, #SMs); CTA c processes tiles c, c+grid, ... (N-tile fastest, so the CTAs
// running at the same time share A tiles through L2).  The accumulator is double-buffered in TMEM: while the 8
// epilogue warps drain tile i, the MMA warp already accumulates tile i+1.
// PAIR = true: launched as clusters of 2 CTAs that work on two consecutive M tiles of the same N tile.  The weight
// tile is the same for both, so each CTA fetches half of it and TMA-multicasts it into both shared memories: the
// L2 -> SM operand traffic per CTA drops from 16+20 KB to 16+10 KB per K block (the kernel is bound by that
// traffic, not by the tensor pipe).  A slot is recycled only after BOTH CTAs' MMAs have read it (multicast commit).
// MODE 2 (cta_group::2): the pair computes ONE 256 x BN tile per step: each CTA stages its own 128 activation rows and
// HALF of the weight tile (BN/2 rows); the leader CTA's single MMA thread issues tcgen05.mma.cta_group::2 (M = 256), which
// reads A from each CTA's own shared memory and the two halves of B from both, and accumulates into both CTAs' TMEM.
// Shared-memory operand traffic per SM per K step drops from (128 + BN) to (128 + BN/2) rows — the resource this
// kernel is bound by.  TMA completions of both CTAs are counted on the LEADER's full barrier; the leader's commits are
// multicast to both CTAs' empty / accumulator-full barriers; both epilogues arrive on the leader's accumulator-empty one.
// p.nstg == 2 (opt-in) double-buffers the epilogue's staging tile: the TMA store of tile i drains while tile i+1 is
// assembled in the other buffer, and the residual tile of tile i+1 is prefetched into that buffer one whole tile ahead, so
// the per-tile chain "previous store drained -> residual loaded -> TMEM load -> math -> store" loses its two TMA latencies.
// The ring gives up one staging tile's worth of stages for it.
template <int BN, int MODE, int EPIW>
__global__ void __launch_bounds__(64 + 32 * EPIW, 1) igemm_kernel(const __grid_constant__ IgemmParams p) {
    using Cfg = IgemmCfg<BN>;
    static_assert(EPIW == 8 || EPIW == 16, "8 or 16 epilogue warps");
    constexpr int IGEMM_EPI_WARPS = EPIW, IGEMM_COL_GROUPS = EPIW / 4;
    constexpr bool PAIR = MODE == 1 || MODE == 2;   // a pair works on two M tiles
    constexpr bool TWOSM = MODE == 2;
    // MODE 3: split-K — BOTH CTAs of a pair work on the SAME tile, each on half of the K blocks; rank 1 hands its fp32 partial
    // accumulator to rank 0 through distributed shared memory (into rank 0's operand ring, which is idle once its MMAs have
    // completed) and rank 0 runs the epilogue.  One tile per pair (host: tiles <= SMs / 2): the few-tile launches of the 8x8
    // level keep the wide 128 x 160 tile (shared-memory operand bandwidth is what bounds 128 x 80 tiles) on all SMs.
    constexpr bool SPLITK = MODE == 3;
    constexpr bool CLUSTER2 = MODE != 0;
    constexpr int STAGE_BYTES = TWOSM ? Cfg::A_BYTES + Cfg::B_BYTES / 2 : Cfg::STAGE_BYTES;
    constexpr int MAXS = Cfg::MAX_STAGES;
    const int STAGES = p.stages;
    const int nstg = CLUSTER2 ? 1 : p.nstg;
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte alignment is required by the 128B swizzle atoms
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t stg_base = smem_base + STAGES * Cfg::STAGE_BYTES;   // (MODE 2 uses smaller stages inside the same budget)
    const uint32_t bar_base = stg_base + nstg * Cfg::STG_BYTES;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (MAXS + s); };
    auto tmem_full_bar = [&](int a) { return bar_base + 8u * (2 * MAXS + a); };
    auto tmem_empty_bar = [&](int a) { return bar_base + 8u * (2 * MAXS + 2 + a); };
    const uint32_t res_full_bar = bar_base + 8u * (2 * MAXS + 4);
    const uint32_t tmem_slot = bar_base + 8u * (2 * MAXS + 5);
    const uint32_t xfree_bar = bar_base + 8u * (2 * MAXS + 6);   // MODE 3, in rank 1: rank 0's MMAs have completed (its ring is idle)
    const uint32_t xfull_bar = bar_base + 8u * (2 * MAXS + 7);   // MODE 3, in rank 0: rank 1's partial accumulator has landed
    uint8_t* stg_gen = smem_raw + (stg_base - smem_u32(smem_raw));   // generic pointer to the staging tile
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
    float* sbias = reinterpret_cast<float*>(smem_raw + (bar_base + 256 - smem_u32(smem_raw)));   // [2][BN]

    // warp index broadcast from lane 0 so the compiler knows the role branches below are warp-uniform
    const int warp = __shfl_sync(0xffffffffu, int(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;
    const int tiles_n = p.tiles_nn;
    // work items: PAIR -> (pair of M tiles, N tile); this CTA takes M tile 2*mp + rank (a tile past the end is all
    // out-of-bounds: zero-filled loads, clipped stores)
    const uint32_t crank = CLUSTER2 ? cluster_ctarank() : 0u;
    const int num_tiles = (PAIR ? (p.tiles_m + 1) / 2 : p.tiles_m) * tiles_n;
    const int wi0 = CLUSTER2 ? int(blockIdx.x >> 1) : int(blockIdx.x);
    const int wstep = CLUSTER2 ? int(gridDim.x >> 1) : int(gridDim.x);

    int total_kb = 0;
    for (int s = 0; s < p.nseg; ++s) total_kb += p.seg[s].cblocks;
    // K blocks of this CTA: all of them, or (MODE 3) the first / second half
    const int kb_lo = (SPLITK && crank != 0) ? (total_kb + 1) / 2 : 0;
    const int kb_hi = (SPLITK && crank == 0) ? (total_kb + 1) / 2 : total_kb;

    if (warp == IGEMM_EPI_WARPS && lane == 0) {
        prefetch_tmap(&p.tmB);
        prefetch_tmap(&p.tmA[0]);
        prefetch_tmap(&p.tmOut);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), MODE == 1 ? 2 : 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(tmem_full_bar(a), 1);
            mbar_init(tmem_empty_bar(a), TWOSM ? 2 * IGEMM_EPI_WARPS : IGEMM_EPI_WARPS);
        }
        mbar_init(res_full_bar, 1);
        if constexpr (SPLITK) {
            mbar_init(xfree_bar, 1);
            mbar_init(xfull_bar, IGEMM_EPI_WARPS * 32);
        }
        fence_barrier_init();
    }
    if (warp == IGEMM_EPI_WARPS + 1) {
        if constexpr (TWOSM) {
            tmem_alloc_2sm(tmem_slot, Cfg::TMEM_COLS);
            tmem_relinquish_2sm();
        } else {
            tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
            tmem_relinquish();
        }
    }
    tc_fence_before();
    __syncthreads();
    if constexpr (CLUSTER2) cluster_sync_all();     // the peer's barriers must exist before anything is multicast at them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    pdl_trigger();   // the next kernel may start its own prologue
    pdl_wait();      // everything above overlapped the predecessor's tail; its outputs are visible from here on

    // Warp roles: the two single-thread issuers sit in the HIGHEST warp ids (8 = TMA, 9 = MMA): the sub-partition
    // arbiter favours higher warp ids, so epilogue warps can never starve them.
    // The issuing thread of each is chosen with elect.sync, not `lane == 0`: only then does ptxas know that exactly
    // one thread executes the UTMALDG / UTCHMMA / UTCBAR instructions and emit them straight; behind a lane test
    // it wraps every one of them in an ELECT / BRA.U.ANY "waterfall" loop (~44 issue cycles per MMA, measured in
    // tools/microbench/mma_rate.cu), which made the MMA thread, not the tensor pipe, the pace-setter of the main loop.
    if (warp == IGEMM_EPI_WARPS) {
        if (elect_one()) {
            // ===== TMA producer =====
            int stage = 0;
            uint32_t phase = 0;
            TileIter ti;
            if constexpr (!PAIR) ti.init(wi0, wstep, tiles_n, p.tiles_w, p.tiles_h);
            for (int tile = wi0; tile < num_tiles; tile += wstep) {
                int nt, w0, h0, n0;
                if constexpr (PAIR) {
                    nt = tile % tiles_n;
                    const int mt = 2 * (tile / tiles_n) + int(crank);
                    w0 = (mt % p.tiles_w) * p.tw;
                    h0 = ((mt / p.tiles_w) % p.tiles_h) * p.th;
                    n0 = (mt / (p.tiles_w * p.tiles_h)) * p.tn;
                } else {
                    nt = ti.nt; w0 = ti.iw * p.tw; h0 = ti.ih * p.th; n0 = ti.ig * p.tn;
                    ti.next();
                }
                int kcol = 0, kbi = 0;
                for (int s = 0; s < p.nseg; ++s) {
                    const IgemmSeg sg = p.seg[s];
                    const void* tm = &p.tmA[sg.map];
                    for (int cb = 0; cb < sg.cblocks; ++cb, kcol += BK) {
                        if constexpr (SPLITK) {
                            const int k_ = kbi++;
                            if (k_ < kb_lo || k_ >= kb_hi) continue;       // the peer's half of the K blocks
                        }
                        mbar_wait(empty_bar(stage), phase ^ 1);
                        const uint32_t a_dst = smem_base + stage * STAGE_BYTES;
                        if constexpr (TWOSM) {
                            // both CTAs' bytes are counted on the leader's barrier
                            if (crank == 0) mbar_expect_tx(full_bar(stage), 2 * STAGE_BYTES);
                            const uint32_t lbar = mapa_shared(full_bar(stage), 0);
                            tma_load_4d_2sm(a_dst, tm, lbar, sg.c0 + cb * BK, w0 + sg.dw, h0 + sg.dh, n0);
                            tma_load_2d_2sm(a_dst + Cfg::A_BYTES, &p.tmBh, lbar, kcol, nt * BN + int(crank) * (BN / 2));
                        } else {
                            mbar_expect_tx(full_bar(stage), Cfg::STAGE_BYTES);
                            tma_load_4d(a_dst, tm, full_bar(stage), sg.c0 + cb * BK, w0 + sg.dw, h0 + sg.dh, n0);
                            if constexpr (MODE == 1) {
                                constexpr int HB = (BN / 2) * BK * 2;    // bytes of half a weight tile
                                tma_load_2d_mc(a_dst + Cfg::A_BYTES + crank * HB, &p.tmBh, full_bar(stage), kcol,
                                               nt * BN + int(crank) * (BN / 2), uint16_t(3));
                            } else {
                                tma_load_2d(a_dst + Cfg::A_BYTES, &p.tmB, full_bar(stage), kcol, nt * BN);
                            }
                        }
                        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == IGEMM_EPI_WARPS + 1) {
        if ((!TWOSM || crank == 0) && elect_one()) {
            // ===== MMA issuer (MODE 2: leader CTA only) =====
            constexpr uint32_t idesc = make_idesc_bf16(TWOSM ? 2 * BM : BM, BN);
            int li = 0, stage = 0;
            uint32_t phase = 0;
            for (int tile = wi0; tile < num_tiles; tile += wstep, ++li) {
                const int as = li & 1;
                mbar_wait(tmem_empty_bar(as), ((li >> 1) & 1) ^ 1);   // epilogue has drained this accumulator slot
                tc_fence_after();
                const uint32_t tacc = tmem_base + as * Cfg::ACC_COLS;
                const int my_kb = kb_hi - kb_lo;
                for (int kb = 0; kb < my_kb; ++kb) {
                    mbar_wait(full_bar(stage), phase);
                    tc_fence_after();
                    const uint32_t a_addr = smem_base + stage * STAGE_BYTES;
                    const uint64_t adesc = make_desc_k_sw128(a_addr);
                    const uint64_t bdesc = make_desc_k_sw128(a_addr + Cfg::A_BYTES);
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {
                        // advancing 16 bf16 (32 B) along K inside the 128B swizzle row: +2 in the (addr >> 4) field
                        if constexpr (TWOSM) umma_bf16_2sm(tacc, adesc + uint64_t(2 * k), bdesc + uint64_t(2 * k), idesc, (kb | k) != 0);
                        else umma_bf16(tacc, adesc + uint64_t(2 * k), bdesc + uint64_t(2 * k), idesc, (kb | k) != 0);
                    }
                    // smem slot reusable once these MMAs have read it (pair modes: tell the peer too)
                    if constexpr (TWOSM) umma_commit_2sm_mc(empty_bar(stage), uint16_t(3));
                    else if constexpr (MODE == 1) umma_commit_mc(empty_bar(stage), uint16_t(3));
                    else umma_commit(empty_bar(stage));
                    if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                }
                // accumulator complete
                if constexpr (TWOSM) umma_commit_2sm_mc(tmem_full_bar(as), uint16_t(3));
                else umma_commit(tmem_full_bar(as));
                if constexpr (SPLITK) {
                    // rank 0's operand ring is idle from here on: tell rank 1 it may park its partial accumulator there
                    if (crank == 0) umma_commit_mc(xfree_bar, uint16_t(2));
                }
            }
        }
        __syncwarp();
    } else {
        // ===== epilogue: warps 0..7.  TMEM lane quarter = warp % 4; the two warps of a quarter split the columns.
        // Per tile: (leader) wait until the previous tile's TMA store has drained the staging tile, TMA-load the res1
        // tile into it -> every thread: accumulator row chunk from TMEM, + bias / row bias, x alpha, + res1 (read from
        // the staging tile) + res2 -> bf16 back into the staging tile -> (leader) TMA store.  All global traffic of
        // the epilogue except the rare res2 is therefore full-line TMA traffic, and M/N tails are clipped by TMA.
        constexpr int BOXC = Cfg::BOXC;
        const int q = warp & 3;
        const int half = warp >> 2;          // column group of this warp: 0 .. IGEMM_COL_GROUPS-1 (named for the 2-group layout)
        const bool leader = (warp == 0 && lane == 0);
        const int r = q * 32 + lane;  // row of the tile == TMEM lane
        const int rw = r % p.tw;
        const int rh = (r / p.tw) % p.th;
        const int rn = r / (p.tw * p.th);
        const float alpha = p.alpha ? __ldg(p.alpha) : 1.0f;
        const int bn_out = p.geglu ? BN / 2 : BN;      // output columns per tile
        const int nbox = bn_out / BOXC;
        int li = 0;
        TileIter ti;
        if constexpr (!PAIR) ti.init(wi0, wstep, tiles_n, p.tiles_w, p.tiles_h);
        // which per-element extras this launch needs (uniform for the whole launch): 0 = bias only, 1 = + res1 tile,
        // 2 = anything else (per-row row bias, alpha, res2).  Each flavour gets its own straight-line code.
        const int flavour = (p.alpha || p.res2 || (p.rowbias && p.tn != 1)) ? 2 : (p.res1 ? 1 : 0);
        // residual tile of (nt, w0, h0, n0) -> staging buffer `buf` (leader only)
        auto load_res = [&](int buf, int nt_, int w0_, int h0_, int n0_) {
            mbar_expect_tx(res_full_bar, uint32_t(nbox) * BM * BOXC * 2);
            for (int bx = 0; bx < nbox; ++bx)
                tma_load_4d(stg_base + buf * Cfg::STG_BYTES + bx * (BM * BOXC * 2), &p.tmRes, res_full_bar, nt_ * bn_out + bx * BOXC,
                            w0_, h0_, n0_);
        };
        if constexpr (!PAIR) {
            if (nstg == 2 && p.res1 != nullptr && leader && wi0 < num_tiles)
                load_res(0, ti.nt, ti.iw * p.tw, ti.ih * p.th, ti.ig * p.tn);      // first tile; later ones are prefetched
        }
        // MODE 3, rank 1: no epilogue of its own — its slice of the partial accumulator goes, thread for thread, into rank 0's
        // (idle) operand ring: 16-byte chunk c of epilogue thread t at ((c * threads + t) * 16), so that consecutive threads
        // write consecutive chunks; rank 0's thread t adds exactly what rank 1's thread t wrote.
        constexpr int XTHREADS = IGEMM_EPI_WARPS * 32;
        bool handed_over = false;
        if constexpr (SPLITK) {
            if (crank != 0 && wi0 < num_tiles) {
                constexpr int NCH = BN / 16, NG = IGEMM_COL_GROUPS;
                constexpr int MAXC = (NCH + NG - 1) / NG;
                constexpr int CBASE = NCH / NG, CREM = NCH % NG;
                const int c_begin = half * CBASE + (half < CREM ? half : CREM);
                const int c_cnt = CBASE + (half < CREM ? 1 : 0);
                mbar_wait_relaxed(tmem_full_bar(0), 0);
                tc_fence_after();
                uint32_t v[MAXC][16];
                const uint32_t trow = tmem_base + (uint32_t(q * 32) << 16);
#pragma unroll
                for (int i = 0; i < MAXC; ++i)
                    if (i < c_cnt) tmem_ld16(trow + (c_begin + i) * 16, v[i]);
                tmem_wait_ld();
                mbar_wait_relaxed(xfree_bar, 0);                 // rank 0's MMAs no longer read its ring
                const uint32_t xdst = mapa_shared(smem_base, 0) + uint32_t(threadIdx.x) * 16u;
#pragma unroll
                for (int i = 0; i < MAXC; ++i) {
                    if (i < c_cnt) {
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            st_cluster_v4(xdst + uint32_t((i * 4 + j) * XTHREADS) * 16u, v[i][4 * j], v[i][4 * j + 1], v[i][4 * j + 2],
                                          v[i][4 * j + 3]);
                    }
                }
                mbar_arrive_cluster(mapa_shared(xfull_bar, 0));  // release.cluster: this thread's stores above are ordered before it
            }
            handed_over = crank != 0;
        }
        for (int tile = wi0; tile < num_tiles && !handed_over; tile += wstep, ++li) {
            int nt, mt, w0, h0, n0;
            if constexpr (PAIR) {
                nt = tile % tiles_n;
                mt = 2 * (tile / tiles_n) + int(crank);
                w0 = (mt % p.tiles_w) * p.tw;
                h0 = ((mt / p.tiles_w) % p.tiles_h) * p.th;
                n0 = (mt / (p.tiles_w * p.tiles_h)) * p.tn;
            } else {
                nt = ti.nt; mt = ti.mt(); w0 = ti.iw * p.tw; h0 = ti.ih * p.th; n0 = ti.ig * p.tn;
                ti.next();
            }
            const int ow = w0 + rw, oh = h0 + rh, on = n0 + rn;
            const bool valid = (ow < p.Wo) && (oh < p.Ho) && (on < p.Bn);
            // Column bias of this tile (bias + the row bias when the whole tile lies in one image), fetched into shared
            // memory while the MMAs of the tile are still running; the per-element code below then only does
            // broadcast LDS instead of dependent global loads.  A tile spanning several images (8x8 level) keeps
            // the per-row global read of the row bias.
            const int as = li & 1;
            const int sbuf = nstg == 2 ? as : 0;                       // staging buffer of this tile
            const uint32_t stg_b = stg_base + sbuf * Cfg::STG_BYTES;
            uint8_t* const stg_g = stg_gen + sbuf * Cfg::STG_BYTES;
            float* sb = sbias + as * BN;
            const bool rb_folded = p.rowbias != nullptr && p.tn == 1;
            const float* rb = (p.rowbias && !rb_folded) ? p.rowbias + static_cast<long long>(valid ? on : 0) * p.rowbias_ld : nullptr;
            if (int(threadIdx.x) < BN) {
                const int n = nt * BN + int(threadIdx.x);
                float bv = 0.f;
                if (n < p.N) {
                    if (p.bias) bv = __ldg(p.bias + n);
                    if (rb_folded) bv += __ldg(p.rowbias + static_cast<long long>(min(n0, p.Bn - 1)) * p.rowbias_ld + n);
                }
                sb[threadIdx.x] = bv;
            }
            const uint32_t trow = tmem_base + as * Cfg::ACC_COLS + (uint32_t(q * 32) << 16);
            // The staging tile is needed (a) for the residual tile, which TMA loads into it as soon as the previous
            // tile's store has drained it — at the tile start — or (b) without a residual only once the accumulator
            // slice sits in registers: then the leader's wait for the previous store and the barrier that publishes it
            // (and the column bias) move behind the TMEM load and the store drains under the accumulator wait.
            // (Reading the residual straight from global into registers was measured slower: row-strided 16-byte loads.)
            const bool early_stage = p.res1 != nullptr;
            if (early_stage) {
                if (leader && nstg == 1) {
                    tma_store_wait_read();
                    load_res(0, nt, w0, h0, n0);
                }
                __syncwarp();
                named_bar_sync(1, IGEMM_EPI_WARPS * 32);   // staging tile is being filled with res1; sb is visible
            }
            // once this warp's accumulator slice sits in registers the slot goes back to the MMA warp
            auto release_acc = [&]() {
                tmem_wait_ld();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if constexpr (TWOSM) mbar_arrive_cluster(mapa_shared(tmem_empty_bar(as), 0));   // the leader issues the MMAs
                    else mbar_arrive(tmem_empty_bar(as));
                }
                if (!early_stage) {
                    if (leader) {                          // the last store out of THIS staging buffer no longer reads it
                        if (nstg == 2) tma_store_wait_read_but1();
                        else tma_store_wait_read();
                    }
                    __syncwarp();
                    named_bar_sync(1, IGEMM_EPI_WARPS * 32);   // staging tile is free; sb is visible
                }
            };

            if (!p.geglu) {
                constexpr int NCH = BN / 16;                 // 16-column chunks, dealt to the column groups as evenly as possible
                constexpr int NG = IGEMM_COL_GROUPS;
                constexpr int MAXC = (NCH + NG - 1) / NG;
                constexpr int CBASE = NCH / NG, CREM = NCH % NG;
                const int c_begin = half * CBASE + (half < CREM ? half : CREM);
                const int c_cnt = CBASE + (half < CREM ? 1 : 0);
                mbar_wait_relaxed(tmem_full_bar(as), (li >> 1) & 1);
                tc_fence_after();
                // fetch this warp's whole accumulator slice with back-to-back tcgen05.ld and ONE wait
                uint32_t v[MAXC][16];
#pragma unroll
                for (int i = 0; i < MAXC; ++i)
                    if (i < c_cnt) tmem_ld16(trow + (c_begin + i) * 16, v[i]);
                release_acc();
                if constexpr (SPLITK) {
                    // the peer's half of the K sum: added in a fixed order (first half + second half) -> deterministic
                    mbar_wait_cluster(xfull_bar, 0);
                    const uint8_t* xsrc = smem_raw + (smem_base - smem_u32(smem_raw)) + threadIdx.x * 16;
#pragma unroll
                    for (int i = 0; i < MAXC; ++i) {
                        if (i < c_cnt) {
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const float4 t = *reinterpret_cast<const float4*>(xsrc + static_cast<size_t>((i * 4 + j) * XTHREADS) * 16);
                                v[i][4 * j + 0] = __float_as_uint(__uint_as_float(v[i][4 * j + 0]) + t.x);
                                v[i][4 * j + 1] = __float_as_uint(__uint_as_float(v[i][4 * j + 1]) + t.y);
                                v[i][4 * j + 2] = __float_as_uint(__uint_as_float(v[i][4 * j + 2]) + t.z);
                                v[i][4 * j + 3] = __float_as_uint(__uint_as_float(v[i][4 * j + 3]) + t.w);
                            }
                        }
                    }
                }
                if (p.res1) mbar_wait_relaxed(res_full_bar, li & 1);
                // fused output statistics: which (image, slot) this warp's 32 rows belong to
                const bool do_stats = p.stats != nullptr;
                const int stat_img = p.tn == 1 ? n0 : n0 + (q * 32) / p.stats_rows_img;
                const int stat_slot = p.stats_tile_base + (p.tn == 1 ? (mt % (p.tiles_w * p.tiles_h)) * 4 + q : ((q * 32) % p.stats_rows_img) >> 5);
                auto body = [&](auto res1_c, auto extra_c) {
                    constexpr bool RES1 = decltype(res1_c)::value, EXTRA = decltype(extra_c)::value;
#pragma unroll
                    for (int i = 0; i < MAXC; ++i) {
                        if (i < c_cnt) {
                            const int col = (c_begin + i) * 16;
                            float xs[16];
#pragma unroll
                            for (int j = 0; j < 2; ++j) {
                                uint4* sp = reinterpret_cast<uint4*>(stg_g + stg_off<BOXC>(r, col + j * 8));
                                const float4 b0 = *reinterpret_cast<const float4*>(sb + col + j * 8);
                                const float4 b1 = *reinterpret_cast<const float4*>(sb + col + j * 8 + 4);
                                float f[8];
                                f[0] = __uint_as_float(v[i][j * 8 + 0]) + b0.x; f[1] = __uint_as_float(v[i][j * 8 + 1]) + b0.y;
                                f[2] = __uint_as_float(v[i][j * 8 + 2]) + b0.z; f[3] = __uint_as_float(v[i][j * 8 + 3]) + b0.w;
                                f[4] = __uint_as_float(v[i][j * 8 + 4]) + b1.x; f[5] = __uint_as_float(v[i][j * 8 + 5]) + b1.y;
                                f[6] = __uint_as_float(v[i][j * 8 + 6]) + b1.z; f[7] = __uint_as_float(v[i][j * 8 + 7]) + b1.w;
                                if constexpr (EXTRA) {
                                    const int nn = nt * BN + col + j * 8;
                                    if (rb && nn < p.N) add_f32x8(f, rb + nn);
#pragma unroll
                                    for (int e = 0; e < 8; ++e) f[e] *= alpha;
                                    if (p.res1) add_bf16x8(f, *sp);
                                    if (p.res2 && valid && nn < p.N) {  // rare (tap sites): straight from global
                                        // linear pixel index inside the full output tensor
                                        const long long m = (static_cast<long long>(on) * p.o_Hf + oh * p.o_step + p.o_py) * p.o_Wf +
                                                            ow * p.o_step + p.o_px;
                                        add_bf16x8(f, __ldg(reinterpret_cast<const uint4*>(p.res2 + m * p.out_ld + nn)));
                                    }
                                } else if constexpr (RES1) {
                                    add_bf16x8(f, *sp);
                                }
                                const uint4 pk = pack_bf16x8(f);
                                *sp = pk;
                                if (do_stats) {        // the ROUNDED values, as a separate statistics pass over the tensor would see them
                                    float2 t2;
                                    t2 = unpack_bf16x2(pk.x); xs[j * 8 + 0] = t2.x; xs[j * 8 + 1] = t2.y;
                                    t2 = unpack_bf16x2(pk.y); xs[j * 8 + 2] = t2.x; xs[j * 8 + 3] = t2.y;
                                    t2 = unpack_bf16x2(pk.z); xs[j * 8 + 4] = t2.x; xs[j * 8 + 5] = t2.y;
                                    t2 = unpack_bf16x2(pk.w); xs[j * 8 + 6] = t2.x; xs[j * 8 + 7] = t2.y;
                                }
                            }
                            if (do_stats) {
                                // Fused GroupNorm statistics from the epilogue REGISTERS: per-column sum / sum of squares over this
                                // warp's 32 rows (rows outside the tensor contribute zero), written — not accumulated — to the warp's
                                // own slot, so the consumer's reduction order is fixed (deterministic) and no barrier or shared-memory
                                // round trip is added to the tile.
                                float sq16[16];
#pragma unroll
                                for (int e = 0; e < 16; ++e) {
                                    xs[e] = valid ? xs[e] : 0.f;
                                    sq16[e] = xs[e] * xs[e];
                                }
                                const float csum = colsum16(xs, lane), csq = colsum16(sq16, lane);
                                const int nn = nt * BN + col + (lane >> 1);
                                if ((lane & 1) == 0 && stat_img < p.Bn && nn < p.out_ld)
                                    reinterpret_cast<float2*>(p.stats)[(static_cast<size_t>(stat_img) * p.stats_tiles + stat_slot) * p.out_ld + nn] =
                                        make_float2(csum, csq);
                            }
                        }
                    }
                };
                if (flavour == 0) body(std::false_type{}, std::false_type{});
                else if (flavour == 1) body(std::true_type{}, std::false_type{});
                else body(std::false_type{}, std::true_type{});
            } else {
                // GEGLU (S/models/activations.py:100-103): out = value * gelu_erf(gate); tile = [64 value | 64 gate]
                if constexpr (BN == 128) {
                    mbar_wait_relaxed(tmem_full_bar(as), (li >> 1) & 1);
                    tc_fence_after();
                    constexpr int GC = 4 / IGEMM_COL_GROUPS;                 // 16-column value chunks per column group
                    uint32_t v[GC][16], g[GC][16];
#pragma unroll
                    for (int c = 0; c < GC; ++c) {
                        tmem_ld16(trow + half * (GC * 16) + c * 16, v[c]);        // value columns; gate = +64
                        tmem_ld16(trow + 64 + half * (GC * 16) + c * 16, g[c]);
                    }
                    release_acc();
#pragma unroll
                    for (int c = 0; c < GC; ++c) {
                        const int col = half * (GC * 16) + c * 16;
#pragma unroll
                        for (int j = 0; j < 2; ++j) {
                            float f[8], bv[8], bg[8];
                            *reinterpret_cast<float4*>(bv) = *reinterpret_cast<const float4*>(sb + col + j * 8);
                            *reinterpret_cast<float4*>(bv + 4) = *reinterpret_cast<const float4*>(sb + col + j * 8 + 4);
                            *reinterpret_cast<float4*>(bg) = *reinterpret_cast<const float4*>(sb + 64 + col + j * 8);
                            *reinterpret_cast<float4*>(bg + 4) = *reinterpret_cast<const float4*>(sb + 64 + col + j * 8 + 4);
#pragma unroll
                            for (int e = 0; e < 8; e += 2) {       // pairwise: FADD2 / FFMA2 / FMUL2 (half the issue slots)
                                const float2 val = fadd2(make_float2(__uint_as_float(v[c][j * 8 + e]), __uint_as_float(v[c][j * 8 + e + 1])),
                                                         make_float2(bv[e], bv[e + 1]));
                                const float2 gate = fadd2(make_float2(__uint_as_float(g[c][j * 8 + e]), __uint_as_float(g[c][j * 8 + e + 1])),
                                                          make_float2(bg[e], bg[e + 1]));
                                const float2 o = fmul2(val, gelu_erf_f2(gate));
                                f[e] = o.x;
                                f[e + 1] = o.y;
                            }
                            *reinterpret_cast<uint4*>(stg_g + stg_off<BOXC>(r, col + j * 8)) = pack_bf16x8(f);
                        }
                    }
                }
            }
            fence_proxy_async_smem();                  // generic-proxy writes of the staging tile -> visible to TMA
            named_bar_sync(1, IGEMM_EPI_WARPS * 32);
            if (leader) {
                for (int bx = 0; bx < nbox; ++bx)
                    tma_store_4d(&p.tmOut, stg_b + bx * (BM * BOXC * 2), nt * bn_out + bx * BOXC, w0, h0, n0);
                tma_store_commit();
                if constexpr (!PAIR) {
                    if (nstg == 2 && p.res1 != nullptr && tile + wstep < num_tiles) {
                        // prefetch the NEXT tile's residual into the other buffer (ti already points at that tile) as
                        // soon as the store issued one tile ago has finished reading it
                        tma_store_wait_read_but1();
                        load_res(sbuf ^ 1, ti.nt, ti.iw * p.tw, ti.ih * p.th, ti.ig * p.tn);
                    }
                }
            }
            __syncwarp();
        }
        if (leader) tma_store_wait_all();              // the stores must have landed before the kernel ends
    }

    tc_fence_before();
    __syncthreads();
    if constexpr (CLUSTER2) cluster_sync_all();     // no CTA may exit while its peer can still multicast / write into it
    if (warp == IGEMM_EPI_WARPS + 1) {
        tc_fence_after();
        if constexpr (TWOSM) tmem_dealloc_2sm(tmem_base, Cfg::TMEM_COLS);
        else tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}

// ------------------------------------------------------------------------------------------------ host side
struct Plan {
    IgemmParams p[4];   // 1 launch, or the 4 sub-pixel phases of an upsample-fused conv
    int nlaunch;
    dim3 grid;
    int bn;
    int mode;   // 0 independent CTAs, 1 pair + weight multicast, 2 pair + cta_group::2 UMMA, 3 split-K pair (DSMEM hand-off)
    int epiw;   // epilogue warps of the kernel variant: 8, or 16 for short-K / few-tile launches (mode 0 only)
    int ktot;
    int stats_tiles;    // per-image tile slots of the fused GroupNorm statistics (0 = not supported for this plan)
    int stats_C;
    int stats_B;
    double flops;
    bool f32;           // fp32 parity mode (mfb_conv_desc.dtype == 1): CUDA-core launches described by q[]
    Conv32Params q[4];
};

static void pick_tile(int W, int H, int B, int* tw, int* th, int* tn) {
    long best = -1;
    for (int a = 128; a >= 1; a >>= 1) {
        for (int b = 128 / a; b >= 1; b >>= 1) {
            const int c = 128 / (a * b);
            if (a * b * c != 128) continue;
            const long tiles = long((W + a - 1) / a) * ((H + b - 1) / b) * ((B + c - 1) / c);
            // fewest tiles wins; ties prefer wide boxes (longer contiguous runs)
            if (best < 0 || tiles < best) {
                best = tiles;
                *tw = a; *th = b; *tn = c;
            }
        }
    }
}

template <int BN, int MODE, int EPIW = 8>
static int launch_igemm(const Plan& pl, const IgemmParams& prm, cudaStream_t st) {
    using Cfg = IgemmCfg<BN>;
    static bool configured = false;
    if (!configured) {
        MFB_CUDA_OK(cudaFuncSetAttribute(igemm_kernel<BN, MODE, EPIW>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
        configured = true;
    }
    MFB_CUDA_OK(launch_k(igemm_kernel<BN, MODE, EPIW>, pl.grid, dim3(64 + 32 * EPIW), Cfg::SMEM_BYTES, st, MODE != 0 ? 2 : 1, prm));
    return MFB_OK;
}

}  // namespace mfb

using namespace mfb;

// Tensor map over an NHWC tensor [B, Hf, Wf, C], optionally restricted to the pixels (step*h + py, step*w + px):
// the view has dims (C, Wv, Hv, B).
static int encode_nhwc(CUtensorMap* tm, const void* base, int C, int Wf, int Hf, int B, int step, int py, int px,
                       const uint32_t* box, int swz) {
    const int Wv = (Wf - px + step - 1) / step, Hv = (Hf - py + step - 1) / step;
    const uint64_t dims[4] = {uint64_t(C), uint64_t(Wv), uint64_t(Hv), uint64_t(B)};
    const uint64_t str[3] = {uint64_t(C) * 2 * step, uint64_t(Wf) * C * 2 * step, uint64_t(Hf) * Wf * C * 2};
    const char* b = static_cast<const char*>(base) + (size_t(py) * Wf + px) * C * 2;
    return encode_tmap_bf16(tm, b, 4, dims, str, box, swz);
}

// One kernel launch worth of parameters.  up_py/up_px >= 0: sub-pixel phase of a 3x3 conv over the nearest-2x
// upsample of x (Upsample2D, S/models/upsampling.py:167-184): output pixels (2h+py, 2w+px) only see a 2x2
// neighbourhood of the low-resolution input, with the 3x3 taps that fall on the same source pixel pre-summed
// into one weight, so the upsampled tensor never exists and 4/9 of the MMA work remains.
static int build_params(const mfb_conv_desc* d, int up_py, int up_px, const void* w, Plan* pl, IgemmParams& p) {
    memset(&p, 0, sizeof(p));
    const bool up = up_py >= 0;
    const int B = d->B, H = d->H, W = d->W;
    // GEMM-row geometry (one row per computed output pixel of this launch)
    const int Ho = up ? H : (d->stride == 1 ? H : (H - 1) / 2 + 1);
    const int Wo = up ? W : (d->stride == 1 ? W : (W - 1) / 2 + 1);
    // geometry of the full output tensor and of this launch's view of it
    const int ostep = up ? 2 : 1, opy = up ? up_py : 0, opx = up ? up_px : 0;
    const int Hf = Ho * ostep, Wf = Wo * ostep;
    p.Wo = Wo; p.Ho = Ho; p.Bn = B;
    p.M = B * Ho * Wo;
    p.N = d->Cout;
    p.o_step = ostep; p.o_py = opy; p.o_px = opx; p.o_Hf = Hf; p.o_Wf = Wf;
    pick_tile(Wo, Ho, B, &p.tw, &p.th, &p.tn);
    p.tiles_w = (Wo + p.tw - 1) / p.tw;
    p.tiles_h = (Ho + p.th - 1) / p.th;
    const int tiles_n = (B + p.tn - 1) / p.tn;
    const uint32_t box[4] = {64u, uint32_t(p.tw), uint32_t(p.th), uint32_t(p.tn)};

    int ktot = 0, nseg = 0;
    int rc = MFB_OK;
    if (up) {
        rc = encode_nhwc(&p.tmA[0], d->x, d->Cin, W, H, B, 1, 0, 0, box, 128);
        if (rc) return rc;
        for (int ty = 0; ty < 2; ++ty)
            for (int tx = 0; tx < 2; ++tx) {
                p.seg[nseg++] = IgemmSeg{0, up_py - 1 + ty, up_px - 1 + tx, 0, d->Cin / 64};
                ktot += d->Cin;
            }
    } else if (d->stride == 1) {
        rc = encode_nhwc(&p.tmA[0], d->x, d->Cin, W, H, B, 1, 0, 0, box, 128);
        if (rc) return rc;
        const int r = d->ksize / 2;
        for (int kh = -r; kh <= r; ++kh)
            for (int kw = -r; kw <= r; ++kw) {
                p.seg[nseg++] = IgemmSeg{0, kh, kw, 0, d->Cin / 64};
                ktot += d->Cin;
            }
    } else {
        // four parity views (ph, pw) of the input: element (h2, w2) of view = x[2*h2 + ph, 2*w2 + pw]
        for (int ph = 0; ph < 2; ++ph)
            for (int pw = 0; pw < 2; ++pw) {
                rc = encode_nhwc(&p.tmA[ph * 2 + pw], d->x, d->Cin, W, H, B, 2, ph, pw, box, 128);
                if (rc) return rc;
            }
        // padding 1: input row = 2*oh + kh - 1:  kh=0 -> (parity 1, h2 = oh-1); kh=1 -> (0, oh); kh=2 -> (1, oh)
        // pad0 (zero row/column appended at the bottom/right, conv padding 0): input row = 2*oh + kh:
        //                                          kh=0 -> (0, oh); kh=1 -> (1, oh); kh=2 -> (0, oh+1)
        const int par1[3] = {1, 0, 1}, off1[3] = {-1, 0, 0}, par0[3] = {0, 1, 0}, off0[3] = {0, 0, 1};
        const int* par = d->pad0 ? par0 : par1;
        const int* off = d->pad0 ? off0 : off1;
        for (int kh = 0; kh < 3; ++kh)
            for (int kw = 0; kw < 3; ++kw) {
                p.seg[nseg++] = IgemmSeg{par[kh] * 2 + par[kw], off[kh], off[kw], 0, d->Cin / 64};
                ktot += d->Cin;
            }
    }
    for (int e = 0; e < d->n_extra; ++e) {
        MFB_REQUIRE(d->extra_x[e] && d->extra_C[e] % 64 == 0, "extra segment %d: channels must be a multiple of 64", e);
        const int C = d->extra_C[e];
        const int mi = ((!up && d->stride == 2) ? 4 : 1) + e;
        rc = encode_nhwc(&p.tmA[mi], d->extra_x[e], C, Wf, Hf, B, ostep, opy, opx, box, 128);   // extras live at OUTPUT resolution
        if (rc) return rc;
        p.seg[nseg++] = IgemmSeg{mi, 0, 0, 0, C / 64};
        ktot += C;
    }
    p.nseg = nseg;
    pl->ktot = ktot;

    // BN: 160 divides every conv width of the SD1.5 family (320/640/1280/960); GEGLU needs the 64|64 split.
    // When the wide tile leaves most SMs without work (8x8 latents: M = 64*B), halve it.
    const int tiles_m = p.tiles_w * p.tiles_h * tiles_n;
    int bn = d->geglu ? 128 : (d->Cout % 160 == 0 ? 160 : 128);
    if (!d->geglu) {
        const int sms = device_sm_count() > 0 ? device_sm_count() : 148;
        const long t_wide = long(tiles_m) * ((d->Cout + bn - 1) / bn);
        if (t_wide * 10 < long(sms) * 7 && d->Cout > bn / 2) bn /= 2;
    }
    // Split-K CTA pairs (MODE 3, opt-in: env MFB_IGEMM_SPLITK=1 or desc.igemm_mode 4): when even the WIDE tile leaves at least half
    // of the SMs idle, keep it and give every tile two CTAs, each half of the K blocks (the halved tile is bound by shared-memory
    // operand bandwidth: 26 KB of operands per 128 x 80 x 64 MMA block against 36 KB per 128 x 160 x 64).
    bool splitk = false;
    if (!d->geglu && d->block_n == 0) {
        static const int env_sk = [] { const char* e = getenv("MFB_IGEMM_SPLITK"); return e ? atoi(e) : 0; }();
        const int sms = device_sm_count() > 0 ? device_sm_count() : 148;
        const int bw = d->Cout % 160 == 0 ? 160 : 128;
        const long t_wide = long(tiles_m) * ((d->Cout + bw - 1) / bw);
        if ((env_sk == 1 || d->igemm_mode == 4) && 2 * t_wide <= sms && ktot / BK >= 16) {
            splitk = true;
            bn = bw;
        }
    }
    if (d->block_n == 64 || d->block_n == 80 || d->block_n == 128 || d->block_n == 160) bn = d->block_n;
    MFB_REQUIRE(!d->geglu || bn == 128, "geglu requires block_n 128");
    pl->bn = bn;
    p.tiles_m = tiles_m;
    p.tiles_nn = (d->Cout + bn - 1) / bn;
    {
        const uint64_t dims[2] = {uint64_t(ktot), uint64_t(d->Cout)};
        const uint64_t str[1] = {uint64_t(ktot) * 2};
        const uint32_t bbox[2] = {64u, uint32_t(bn)};
        rc = encode_tmap_bf16(&p.tmB, w, 2, dims, str, bbox, 128);
        if (rc) return rc;
        const uint32_t hbox[2] = {64u, uint32_t(bn / 2)};
        rc = encode_tmap_bf16(&p.tmBh, w, 2, dims, str, hbox, 128);
        if (rc) return rc;
    }
    // CTA-pair modes are opt-in (env MFB_IGEMM_MODE = 1 | 2 or desc.igemm_mode): measured slightly SLOWER than
    // independent CTAs on B200 for these tile shapes (profiles/r01d_pair_multicast_ab.md).
    {
        const char* np = getenv("MFB_IGEMM_MODE");
        pl->mode = (tiles_m >= 2 && np) ? atoi(np) : 0;
        if (pl->mode < 0 || pl->mode > 2) pl->mode = 0;
        if (d->igemm_mode >= 1 && d->igemm_mode <= 3) pl->mode = tiles_m >= 2 ? d->igemm_mode - 1 : 0;
        if (d->igemm_mode == 4) pl->mode = 0;        // split-K where the geometry qualifies, independent CTAs otherwise
        if (splitk) pl->mode = 3;
    }
    {
        // MFB_IGEMM_EPIW = 16: all launches on the 16-epilogue-warp variant; = 1: only short K (<= 2560) and the few-tile 8x8
        // level.  Default 8 everywhere: per shape the 16-warp variant wins 5-11 % on those launches, per STEP neither choice
        // is measurable (24.04-24.21 vs 24.08-24.11 vs 24.21-24.29 ms, profiles/logs/ab_epiw_auto.log)
        const char* ew = getenv("MFB_IGEMM_EPIW");
        const int knob = ew ? atoi(ew) : 8;
        pl->epiw = knob == 16 ? 16 : (knob == 1 && (ktot <= 2560 || tiles_m <= 8)) ? 16 : 8;
        if (pl->mode != 0) pl->epiw = 8;
    }
    {
        const int boxc = (bn % 32 == 0) ? 32 : 16;
        const int out_ld = d->geglu ? d->Cout / 2 : d->Cout;
        const uint32_t obox[4] = {uint32_t(boxc), uint32_t(p.tw), uint32_t(p.th), uint32_t(p.tn)};
        rc = encode_nhwc(&p.tmOut, d->out, out_ld, Wf, Hf, B, ostep, opy, opx, obox, boxc * 2);
        if (rc) return rc;
        if (d->res1) {
            rc = encode_nhwc(&p.tmRes, d->res1, out_ld, Wf, Hf, B, ostep, opy, opx, obox, boxc * 2);
            if (rc) return rc;
        }
    }
    {
        // fused output statistics: every epilogue warp (32 rows of a tile) must lie in ONE image: tiles inside an image
        // (tn == 1), or tiles of whole images whose pixel count is a multiple of 32.  A slot = one warp's rows: 4 per tile.
        const int per_img = p.tiles_w * p.tiles_h;
        const int rows_img = p.tw * p.th;
        const bool ok = !d->geglu && (p.tn == 1 || (per_img == 1 && rows_img % 32 == 0));
        const int nphase = up ? 4 : 1;
        const int slots = p.tn == 1 ? per_img * 4 : rows_img / 32;
        pl->stats_tiles = ok ? slots * nphase : 0;
        pl->stats_C = d->Cout;
        pl->stats_B = B;
        p.stats = nullptr;
        p.stats_tiles = pl->stats_tiles;
        p.stats_tile_base = up ? (up_py * 2 + up_px) * slots : 0;
        p.stats_rows_img = rows_img;
    }
    p.bias = d->bias;
    p.rowbias = d->rowbias;
    p.rowbias_ld = d->rowbias_ld;
    p.alpha = d->alpha;
    p.res1 = static_cast<const __nv_bfloat16*>(d->res1);
    p.res2 = static_cast<const __nv_bfloat16*>(d->res2);
    p.out = static_cast<__nv_bfloat16*>(d->out);
    p.geglu = d->geglu;
    p.out_ld = d->geglu ? d->Cout / 2 : d->Cout;
    {
        // double-buffered epilogue: opt-in (env MFB_IGEMM_NSTG = 2, or MFB_IGEMM_NSTG_KMAX = largest K that gets it).
        // Measured neutral-to-slower on B200 (profiles/r01o_igemm_double_staging.md): these launches are paced by operand
        // fetch and, for N = K = 320, by HBM bytes — not by the epilogue chain — and the ring stage it costs hurts.
        const int sms = device_sm_count() > 0 ? device_sm_count() : 148;
        const char* e1 = getenv("MFB_IGEMM_NSTG");
        const char* e2 = getenv("MFB_IGEMM_NSTG_KMAX");
        const int kmax = e2 ? atoi(e2) : 0;
        const long total = long(p.tiles_m) * p.tiles_nn;
        int nstg = (pl->mode == 0 && ktot <= kmax && total > sms) ? 2 : 1;
        if (e1 && pl->mode == 0) nstg = atoi(e1) == 2 ? 2 : 1;
        const int stg_bytes = BM * bn * 2, stage_bytes = BM * BK * 2 + bn * BK * 2;
        int stages = (224 * 1024 - nstg * stg_bytes) / stage_bytes;
        p.nstg = nstg;
        p.stages = stages > 8 ? 8 : stages;
    }
    {
        const int sms = device_sm_count() > 0 ? device_sm_count() : 148;
        if (pl->mode == 3) {
            pl->grid = dim3(unsigned(2 * long(p.tiles_m) * p.tiles_nn), 1, 1);     // one tile per pair (2 * tiles <= SMs)
        } else if (pl->mode != 0) {
            const long pairs = long((p.tiles_m + 1) / 2) * p.tiles_nn;
            const long maxp = sms / 2;
            pl->grid = dim3(unsigned(2 * (pairs < maxp ? pairs : maxp)), 1, 1);
        } else {
            const long total = long(p.tiles_m) * p.tiles_nn;
            // A/B knob (MFB_IGEMM_MAX_CTAS): cap the persistent grid so that two launch streams (BrushNet || UNet, StepEngine
            // two_streams) share the SMs spatially instead of time-slicing them — the tile loop is the same, results bit-identical
            static const int cap = [] { const char* e = getenv("MFB_IGEMM_MAX_CTAS"); return e ? atoi(e) : 0; }();
            const long lim = (cap > 0 && cap < sms) ? cap : sms;
            pl->grid = dim3(unsigned(total < lim ? total : lim), 1, 1);
        }
    }
    return MFB_OK;
}

extern "C" int mfb_conv_plan_create(const mfb_conv_desc* d, mfb_plan** out) {
    MFB_REQUIRE(d && out, "null argument");
    MFB_REQUIRE(d->ksize == 3 || d->ksize == 1, "ksize must be 1 or 3 (got %d)", d->ksize);
    MFB_REQUIRE(d->stride == 1 || d->stride == 2, "stride must be 1 or 2");
    MFB_REQUIRE(d->stride == 1 || d->ksize == 3, "stride 2 is only supported for 3x3");
    MFB_REQUIRE(d->Cin > 0 && d->Cin % 64 == 0, "Cin must be a positive multiple of 64 (got %d)", d->Cin);
    MFB_REQUIRE(d->Cout > 0 && d->Cout % 8 == 0, "Cout must be a multiple of 8 (got %d)", d->Cout);
    MFB_REQUIRE(d->n_extra >= 0 && d->n_extra <= 3, "at most 3 extra 1x1 segments");
    MFB_REQUIRE(d->x && d->w && d->out, "x / w / out must be device pointers");
    MFB_REQUIRE(!d->geglu || (d->Cout % 128 == 0 && !d->res1 && !d->res2 && !d->rowbias), "geglu needs Cout %% 128 == 0 and no residuals");
    MFB_REQUIRE(!d->up2x || (d->ksize == 3 && d->stride == 1 && !d->geglu), "up2x needs a 3x3 stride-1 conv");
    MFB_REQUIRE(!d->pad0 || d->stride == 2, "pad0 is a stride-2 mode");

    Plan* pl = new Plan();
    int rc;
    pl->f32 = d->dtype == 1;
    if (pl->f32) {
        pl->mode = 0; pl->bn = 0; pl->stats_tiles = 0; pl->stats_C = d->Cout; pl->stats_B = d->B;
        int kext = 0;
        for (int e = 0; e < d->n_extra; ++e) kext += d->extra_C[e];
        pl->nlaunch = d->up2x ? 4 : 1;
        for (int ph = 0; ph < pl->nlaunch; ++ph) {
            const size_t kt = size_t(d->up2x ? 4 : d->ksize * d->ksize) * d->Cin + kext;
            const char* wp = static_cast<const char*>(d->w) + size_t(ph) * d->Cout * kt * 4;
            rc = conv32_build(d, d->up2x ? ph >> 1 : -1, d->up2x ? ph & 1 : -1, wp, pl->q[ph]);
            if (rc) { delete pl; return rc; }
        }
        pl->ktot = pl->q[0].ktot;
        const double Ho = d->up2x ? 2.0 * d->H : double((d->H + d->stride - 1) / d->stride);
        const double Wo = d->up2x ? 2.0 * d->W : double((d->W + d->stride - 1) / d->stride);
        pl->flops = 2.0 * d->B * Ho * Wo * d->Cout * (double(d->ksize) * d->ksize * d->Cin + kext);
        *out = reinterpret_cast<mfb_plan*>(pl);
        return MFB_OK;
    }
    if (d->up2x) {
        // w = [4 phases][Cout][4*Cin + extras], phase index = py*2 + px
        pl->nlaunch = 4;
        for (int ph = 0; ph < 4; ++ph) {
            const int kt = 4 * d->Cin + (d->n_extra > 0 ? d->extra_C[0] : 0) + (d->n_extra > 1 ? d->extra_C[1] : 0) +
                           (d->n_extra > 2 ? d->extra_C[2] : 0);
            const char* wp = static_cast<const char*>(d->w) + size_t(ph) * d->Cout * kt * 2;
            rc = build_params(d, ph >> 1, ph & 1, wp, pl, pl->p[ph]);
            if (rc) { delete pl; return rc; }
        }
        int kext = 0;
        for (int e = 0; e < d->n_extra; ++e) kext += d->extra_C[e];
        pl->flops = 2.0 * double(d->B) * (2 * d->H) * (2 * d->W) * d->Cout * (9.0 * d->Cin + kext);   // algorithmic
    } else {
        pl->nlaunch = 1;
        rc = build_params(d, -1, -1, d->w, pl, pl->p[0]);
        if (rc) { delete pl; return rc; }
        pl->flops = 2.0 * double(pl->p[0].M) * d->Cout * pl->ktot;
    }
    *out = reinterpret_cast<mfb_plan*>(pl);
    return MFB_OK;
}

template <int MODE>
static int run_one(const Plan& pl, const IgemmParams& prm, cudaStream_t st) {
    if constexpr (MODE == 0) {
        if (pl.epiw == 16) {
            switch (pl.bn) {
                case 160: return launch_igemm<160, 0, 16>(pl, prm, st);
                case 128: return launch_igemm<128, 0, 16>(pl, prm, st);
                case 80: return launch_igemm<80, 0, 16>(pl, prm, st);
                default: return launch_igemm<64, 0, 16>(pl, prm, st);
            }
        }
    }
    if constexpr (MODE == 3) {       // split-K pairs exist for the wide tiles only
        return pl.bn == 160 ? launch_igemm<160, 3>(pl, prm, st) : launch_igemm<128, 3>(pl, prm, st);
    } else {
        switch (pl.bn) {
            case 160: return launch_igemm<160, MODE>(pl, prm, st);
            case 128: return launch_igemm<128, MODE>(pl, prm, st);
            case 80: return launch_igemm<80, MODE>(pl, prm, st);
            default: return launch_igemm<64, MODE>(pl, prm, st);
        }
    }
}

extern "C" int mfb_plan_run(mfb_plan* plan, void* stream) {
    MFB_RECORD(mfb_plan_run(plan, stream));
    MFB_REQUIRE(plan, "null plan");
    Plan* pl = reinterpret_cast<Plan*>(plan);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (pl->f32) {
        for (int i = 0; i < pl->nlaunch; ++i) {
            const int rc = conv32_launch(pl->q[i], st);
            if (rc) return rc;
        }
        return MFB_OK;
    }
    for (int i = 0; i < pl->nlaunch; ++i) {
        const int rc = pl->mode == 3 ? run_one<3>(*pl, pl->p[i], st) : pl->mode == 2 ? run_one<2>(*pl, pl->p[i], st)
                       : pl->mode == 1 ? run_one<1>(*pl, pl->p[i], st) : run_one<0>(*pl, pl->p[i], st);
        if (rc) return rc;
    }
    return MFB_OK;
}

extern "C" int mfb_plan_launches(const mfb_plan* plan) { return plan ? reinterpret_cast<const Plan*>(plan)->nlaunch : 0; }

// Fused GroupNorm statistics of the plan's output: number of floats the caller must provide (0 = unsupported)...
extern "C" long long mfb_plan_stats_floats(const mfb_plan* plan) {
    if (!plan) return 0;
    const Plan* pl = reinterpret_cast<const Plan*>(plan);
    return 2LL * pl->stats_B * pl->stats_tiles * pl->stats_C;
}
extern "C" int mfb_plan_stats_tiles(const mfb_plan* plan) { return plan ? reinterpret_cast<const Plan*>(plan)->stats_tiles : 0; }
// ... and the buffer [B][tiles][C][2] fp32 the epilogue writes them to (NULL switches the fusion off).
extern "C" int mfb_plan_set_stats(mfb_plan* plan, float* buf) {
    MFB_REQUIRE(plan, "null plan");
    Plan* pl = reinterpret_cast<Plan*>(plan);
    MFB_REQUIRE(pl->stats_tiles > 0 || !buf, "this plan cannot produce output statistics");
    for (int i = 0; i < pl->nlaunch; ++i) pl->p[i].stats = buf;
    return MFB_OK;
}

extern "C" int mfb_plan_destroy(mfb_plan* plan) {
    delete reinterpret_cast<Plan*>(plan);
    return MFB_OK;
}

extern "C" double mfb_plan_flops(const mfb_plan* plan) { return plan ? reinterpret_cast<const Plan*>(plan)->flops : 0.0; }

extern "C" int mfb_plan_igemm_mode(const mfb_plan* plan) { return plan ? (reinterpret_cast<const Plan*>(plan)->f32 ? 0 : reinterpret_cast<const Plan*>(plan)->mode) : 0; }

extern "C" int mfb_plan_ktotal(const mfb_plan* plan) {
    if (!plan) return 0;
    return reinterpret_cast<const Plan*>(plan)->ktot;
}
