// Thin inline-PTX wrappers for the sm_100a features the kernels use:
// mbarrier, TMA (cp.async.bulk.tensor), tcgen05 MMA / TMEM alloc / TMEM load, proxies and fences.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace mfb {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n"
        ".reg .b32 %%rx;\n"
        ".reg .pred %%px;\n"
        "elect.sync %%rx|%%px, 0xffffffff;\n"
        "selp.u32 %0, 1, 0, %%px;\n"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}

// ------------------------------------------------------------------ mbarrier
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded waits: a protocol bug must surface as a trapped kernel (an error the host sees), never as a hung GPU.  The
// bound is a spin count, not a clock read: the poll loop is try_wait + counter + branch, nothing else — polling warps
// share issue slots with the warps doing the math (40 % of the attention kernel's instructions were poll-loop
// bookkeeping when every iteration also read the clock).
static __device__ __noinline__ void mbar_timeout_trap(uint32_t bar, uint32_t parity) {
    printf("mfb200: mbarrier wait timed out (block %d,%d thread %d bar 0x%x parity %u)\n", blockIdx.x, blockIdx.y, threadIdx.x, bar,
           parity);
    __trap();
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    for (uint32_t spins = 0; !mbar_try_wait(bar, parity); ++spins)
        if (spins > (1u << 28)) mbar_timeout_trap(bar, parity);      // >= seconds
}

// Wait used by warps that have slack (epilogue / softmax / producers): the try_wait carries a suspend-time hint so the
// warp sleeps in hardware instead of hot-spinning.
__device__ __forceinline__ void mbar_wait_relaxed(uint32_t bar, uint32_t parity) {
    for (uint32_t spins = 0;; ++spins) {
        uint32_t ok;
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(ok)
            : "r"(bar), "r"(parity), "r"(2000u)
            : "memory");
        if (ok) return;
        if (spins > (1u << 26)) mbar_timeout_trap(bar, parity);
    }
}

// ------------------------------------------------------------------ TMA
__device__ __forceinline__ void prefetch_tmap(const void* tmap) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tmap)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
        "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
            dst),
        "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
            dst),
        "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}

// L2 prefetch of a 4-D box (no shared-memory destination, no completion tracking)
__device__ __forceinline__ void tma_prefetch_4d(const void* tmap, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global [%0, {%1, %2, %3, %4}];" ::"l"(reinterpret_cast<uint64_t>(tmap)),
                 "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}

// multicast: the box is written at the same CTA-relative smem offset of every CTA in `mask`, and complete_tx is
// signalled on the mbarrier at the same offset in each of them
__device__ __forceinline__ void tma_load_2d_mc(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1, uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], "
        "[%2], %5;" ::"r"(dst),
        "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1), "h"(mask)
        : "memory");
}
// ---- cta_group::2 (two SMs cooperate on one 256-row UMMA tile)
// cluster-shared address of `local_addr` in CTA `rank` of this cluster
__device__ __forceinline__ uint32_t mapa_shared(uint32_t local_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA loads whose completion is signalled on an mbarrier that may live in the peer CTA (cluster-shared address)
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const void* tmap, uint32_t bar_cluster, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            dst),
        "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar_cluster), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_load_4d_2sm(uint32_t dst, const void* tmap, uint32_t bar_cluster, int c0, int c1, int c2,
                                                int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], "
        "[%2];" ::"r"(dst),
        "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t smem_dst, uint32_t ncols) {  // one warp in EACH CTA of the pair
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish_2sm() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem of both CTAs] (+)= A[256 rows: 128 per CTA] * B[N rows: N/2 per CTA]; issued by ONE thread of the leader CTA
__device__ __forceinline__ void umma_bf16_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit_2sm_mc(uint32_t bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                 "h"(mask)
                 : "memory");
}

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// ---- split-K CTA pair (igemm MODE 3): partial accumulators handed over through distributed shared memory
// 16-byte store into a peer CTA's shared memory (cluster-shared address from mapa_shared)
__device__ __forceinline__ void st_cluster_v4(uint32_t cluster_addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(cluster_addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// mbarrier wait whose acquire covers writes made by the peer CTA before its (release.cluster) arrive; suspend-time hint like
// mbar_wait_relaxed, bounded like every wait of this library
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
    for (uint32_t spins = 0;; ++spins) {
        uint32_t ok;
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2, %3;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(ok)
            : "r"(bar), "r"(parity), "r"(2000u)
            : "memory");
        if (ok) return;
        if (spins > (1u << 21)) mbar_timeout_trap(bar, parity);      // seconds: a protocol bug traps, never hangs
    }
}
// commit that arrives on the mbarrier at the same offset in every CTA of `mask`
__device__ __forceinline__ void umma_commit_mc(uint32_t bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                 "h"(mask)
                 : "memory");
}
__device__ __forceinline__ void tma_store_4d(const void* tmap, uint32_t src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                     reinterpret_cast<uint64_t>(tmap)),
                 "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// wait until the bulk stores of this thread have finished READING shared memory (the buffer may be reused)
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// ... all but the most recent bulk store group (double-buffered staging: the previous tile's store may still be draining)
__device__ __forceinline__ void tma_store_wait_read_but1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

__device__ __forceinline__ void named_bar_arrive(int id, int nthreads) {      // non-blocking arrival on a named barrier
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ------------------------------------------------------------------ tcgen05 / TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {  // whole warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // whole warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc]; bf16 x bf16 -> fp32; issued by ONE thread.
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Arrive on an mbarrier once all previously issued MMAs of this thread have completed
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// Instruction descriptor for kind::f16, A/B bf16 K-major, D fp32 (cute/arch/mma_sm100_desc.hpp InstrDescriptor):
// [4,6) c_format=1(F32) | [7,10) a_format=1(BF16) | [10,13) b_format=1(BF16) | bit15 a_major | bit16 b_major |
// [17,23) N>>3 | [24,29) M>>4
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N, int a_mn_major = 0, int b_mn_major = 0) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (uint32_t(a_mn_major) << 15) | (uint32_t(b_mn_major) << 16) |
           (uint32_t(N >> 3) << 17) | (uint32_t(M >> 4) << 24);
}

// Shared-memory matrix descriptor for a K-major operand tile stored as dense 128-byte rows with the
// 128B swizzle (exactly what a TMA box with inner extent 64 bf16 and CU_TENSOR_MAP_SWIZZLE_128B writes):
// start>>4 | LBO(enc 1, unused for swizzled K-major) | SBO = 1024 B (8 rows x 128 B) | version 1 | layout SWIZZLE_128B (2).
// Same physical tile (dense 128-byte rows, 128B swizzle) read as an MN-major operand: a row is one K index holding 64
// contiguous M/N elements.  Canonical form ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units (cute/atom/mma_traits_sm100.hpp
// make_umma_desc<Major::MN>): SBO = 1024 B between groups of 8 K rows, LBO = bytes between 64-element M/N blocks.  A K step
// of 16 advances the start address by 16 rows = 2048 B.  Used for V in attention (P V needs V[key][d] with d contiguous).
__device__ __forceinline__ uint64_t make_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= uint64_t((smem_addr & 0x3FFFF) >> 4);
    d |= uint64_t((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= uint64_t(1024 >> 4) << 32;
    d |= uint64_t(1) << 46;
    d |= uint64_t(2) << 61;
    return d;
}
__device__ __forceinline__ uint64_t make_desc_k_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= uint64_t((smem_addr & 0x3FFFF) >> 4);
    d |= uint64_t(1) << 16;
    d |= uint64_t(1024 >> 4) << 32;
    d |= uint64_t(1) << 46;
    d |= uint64_t(2) << 61;
    return d;
}

// 32 lanes x 32 consecutive fp32 columns: thread i of the warp receives row (lane base + i).
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
        "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ------------------------------------------------------------------ programmatic dependent launch (PDL)
// Every kernel of the step is launched with programmatic stream serialization: it may start (prologue: barrier
// init, TMEM alloc, descriptor prefetch) while its predecessor drains, and must call pdl_wait() before touching
// anything the predecessor wrote.  pdl_trigger() lets the NEXT kernel start early in the same way.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ------------------------------------------------------------------ small helpers
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
    __nv_bfloat162 h = *reinterpret_cast<__nv_bfloat162*>(&u);
    return __bfloat1622float2(h);
}
// x * sigmoid(x) with one ex2 and one approximate reciprocal (2 MUFU + 3 FP32 ops, ~2 ulp): the IEEE division of the
// textbook form costs ~10 more instructions per element and made the GroupNorm+SiLU pass issue-bound, not HBM-bound.
__device__ __forceinline__ float silu_f(float x) {
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * -1.4426950408889634f));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
    return x * r;
}
// 2^x on the FMA / ALU pipes (no MUFU): x = n + f with n = round(x), f in [-0.5, 0.5]; 2^f by a cubic (max relative
// error 1.0e-4), 2^n added into the exponent field.  x is clamped at -126 (result ~1e-38 instead of 0).
__device__ __forceinline__ float ex2_poly(float x) {
    const float xc = fmaxf(x, -126.0f);
    const float t = xc + 12582912.0f;               // 1.5 * 2^23: the low mantissa bits now hold round(xc)
    const float f = xc - (t - 12582912.0f);
    float p = fmaf(5.592203513e-02f, f, 2.426400781e-01f);
    p = fmaf(p, f, 6.931210160e-01f);
    p = fmaf(p, f, 9.999244809e-01f);
    return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}

// exact (erf) GELU, F.gelu default (S/models/activations.py:94-98): gelu(x) = x * Phi(x).  The normal CDF tail is
// evaluated as Phi(-a) = 2^-L(a), a = min(|x|, 6), with L a degree-7 polynomial (Chebyshev fit of -log2(erfc(a/sqrt2)/2)
// on [0, 6]; |gelu error| <= 7e-7 over all x, far below bf16 resolution): 7 FMA + one ex2, branch-free.  (libdevice
// erff is ~40 instructions and the Abramowitz-Stegun form needs a reciprocal as well; the GEGLU epilogue is bound
// by issue slots, not by the tensor pipe, on the K=320 projections.)
__device__ __forceinline__ float gelu_erf_f(float x) {
    const float a = fminf(fabsf(x), 6.0f);
    float l = fmaf(1.889626219e-06f, a, -6.268139987e-05f);
    l = fmaf(l, a, 9.388679173e-04f);
    l = fmaf(l, a, -8.539461531e-03f);
    l = fmaf(l, a, 5.402068794e-02f);
    l = fmaf(l, a, 4.584097862e-01f);
    l = fmaf(l, a, 1.151269197e+00f);
    l = fmaf(l, a, 9.999943376e-01f);
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-l));
    const float phi = x < 0.f ? e : 1.0f - e;
    return x * phi;
}

// ------------------------------------------------------------------ packed fp32 pairs (FFMA2 / FADD2 / FMUL2 on sm_100)
// One instruction, two independent fp32 lanes (PTX fma.rn.f32x2 & co.): halves the issue slots of straight-line fp32
// math.  Used where an epilogue is bound by instruction issue / dependent-FMA latency with only two warps per scheduler
// (the GEGLU epilogue of the K = 320 projection: ncu r01p, issue-active 45 %, tensor pipe 41 %).
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    float2 d;
    asm("{ .reg .b64 ra, rb, rc, rd;\n mov.b64 ra, {%2, %3};\n mov.b64 rb, {%4, %5};\n mov.b64 rc, {%6, %7};\n"
        " fma.rn.f32x2 rd, ra, rb, rc;\n mov.b64 {%0, %1}, rd; }"
        : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
    return d;
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
    float2 d;
    asm("{ .reg .b64 ra, rb, rd;\n mov.b64 ra, {%2, %3};\n mov.b64 rb, {%4, %5};\n add.rn.f32x2 rd, ra, rb;\n mov.b64 {%0, %1}, rd; }"
        : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return d;
}
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
    float2 d;
    asm("{ .reg .b64 ra, rb, rd;\n mov.b64 ra, {%2, %3};\n mov.b64 rb, {%4, %5};\n mul.rn.f32x2 rd, ra, rb;\n mov.b64 {%0, %1}, rd; }"
        : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return d;
}
// gelu_erf_f on two values at once: same polynomial, same results bit for bit (every operation is the same IEEE fma /
// add / mul, only issued pairwise)
__device__ __forceinline__ float2 gelu_erf_f2(float2 x) {
    const float2 a = make_float2(fminf(fabsf(x.x), 6.0f), fminf(fabsf(x.y), 6.0f));
    auto bc = [](float c) { return make_float2(c, c); };
    float2 l = ffma2(bc(1.889626219e-06f), a, bc(-6.268139987e-05f));
    l = ffma2(l, a, bc(9.388679173e-04f));
    l = ffma2(l, a, bc(-8.539461531e-03f));
    l = ffma2(l, a, bc(5.402068794e-02f));
    l = ffma2(l, a, bc(4.584097862e-01f));
    l = ffma2(l, a, bc(1.151269197e+00f));
    l = ffma2(l, a, bc(9.999943376e-01f));
    float e0, e1;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(-l.x));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(-l.y));
    const float2 phi = make_float2(x.x < 0.f ? e0 : 1.0f - e0, x.y < 0.f ? e1 : 1.0f - e1);
    return fmul2(x, phi);
}

}  // namespace mfb
