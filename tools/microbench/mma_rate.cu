// Microbenchmark: issue rate of tcgen05.mma.cta_group::1.kind::f16 (M=128, K=16, bf16->fp32) as a function of N with
// operands resident in shared memory (no TMA traffic): cycles per MMA instruction.  Build:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../reflecting-reality_b200/csrc -o mma_rate mma_rate.cu
#include <cstdio>
#include <cstdlib>
#include "ptx.cuh"
using namespace mfb;

template <int N>
__global__ void __launch_bounds__(128, 1) mma_rate(int iters, long long* out_cycles, int nstages, int same_acc) {
    extern __shared__ uint8_t raw[];
    const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
    constexpr int STAGE = 128 * 128 + N * 128;            // A 128x64 bf16 + B Nx64 bf16
    const uint32_t bar = base + nstages * STAGE;
    const uint32_t slot = bar + 16;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // fill operand memory with small finite values
    for (uint32_t i = threadIdx.x; i < uint32_t(nstages * STAGE / 4); i += blockDim.x)
        reinterpret_cast<uint32_t*>(raw + (base - smem_u32(raw)))[i] = 0x3c003c00u;
    if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
    if (warp == 1) { tmem_alloc(slot, 512); tmem_relinquish(); }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<uint32_t*>(raw + (slot - smem_u32(raw)));
    if (warp == 0 && lane == 0) {
        constexpr uint32_t idesc = make_idesc_bf16(128, N);
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            const uint32_t a = base + (it % nstages) * STAGE;
            const uint64_t ad = make_desc_k_sw128(a), bd = make_desc_k_sw128(a + 128 * 128);
            const uint32_t acc = tmem + (same_acc ? 0 : (it & 1) * 256);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_bf16(acc, ad + 2 * k, bd + 2 * k, idesc, 1);
        }
        umma_commit(bar);
        mbar_wait(bar, 0);
        const long long t1 = clock64();
        if (blockIdx.x == 0) out_cycles[0] = t1 - t0;
    }
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, 512);
}

// Exposed cost of the barrier traffic an MMA-issuing thread executes between groups of 4 MMAs (one 64-wide K block):
// mode bit0 = mbarrier.try_wait on an already-completed phase, bit1 = tcgen05.commit to a barrier nobody waits on,
// bit2 = tcgen05.fence::after_thread_sync, bit3 = a second try_wait, bit4 = commit only every other iteration.
template <int N>
__global__ void __launch_bounds__(128, 1) issue_overhead(int iters, long long* out_cycles, int mode, int per_group) {
    extern __shared__ uint8_t raw[];
    const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
    constexpr int STAGE = 128 * 128 + N * 128;
    const uint32_t bar = base + 3 * STAGE;          // final
    const uint32_t bar_done = bar + 8;              // never arrived: waiting for parity 1 returns at once
    const uint32_t bar_sink = bar + 16;             // commit sink
    const uint32_t slot = bar + 32;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (uint32_t i = threadIdx.x; i < uint32_t(3 * STAGE / 4); i += blockDim.x)
        reinterpret_cast<uint32_t*>(raw + (base - smem_u32(raw)))[i] = 0x3c003c00u;
    if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(bar_done, 1); mbar_init(bar_sink, 1); fence_barrier_init(); }
    if (warp == 1) { tmem_alloc(slot, 512); tmem_relinquish(); }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<uint32_t*>(raw + (slot - smem_u32(raw)));
    if (warp == 0 && lane == 0) {
        constexpr uint32_t idesc = make_idesc_bf16(128, N);
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            if (mode & 1) mbar_wait_relaxed(bar_done, 1);
            if (mode & 8) mbar_wait_relaxed(bar_done, 1);
            if (mode & 4) tc_fence_after();
            const uint32_t a = base + (it % 3) * STAGE;
            const uint64_t ad = make_desc_k_sw128(a), bd = make_desc_k_sw128(a + 128 * 128);
            for (int g = 0; g < per_group; ++g) {
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_bf16(tmem, ad + 2 * k, bd + 2 * k, idesc, 1);
            }
            if ((mode & 2) && (!(mode & 16) || (it & 1))) umma_commit(bar_sink);
        }
        umma_commit(bar);
        mbar_wait(bar, 0);
        const long long t1 = clock64();
        if (blockIdx.x == 0) out_cycles[0] = t1 - t0;
    }
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, 512);
}

template <int N>
void run_overhead(int mode, int per_group) {
    long long* d;
    cudaMalloc(&d, 8);
    const int smem = 3 * (128 * 128 + N * 128) + 2048;
    cudaFuncSetAttribute(issue_overhead<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int iters = 2000;
    issue_overhead<N><<<148, 128, smem>>>(iters, d, mode, per_group);
    issue_overhead<N><<<148, 128, smem>>>(iters, d, mode, per_group);
    cudaError_t e = cudaDeviceSynchronize();
    long long c = 0;
    cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
    printf("overhead N=%3d mode=%2d (wait=%d wait2=%d fence=%d commit=%d alt=%d) groups=%d : %7.1f cycles / 64-K block  %s\n", N, mode,
           mode & 1, (mode >> 3) & 1, (mode >> 2) & 1, (mode >> 1) & 1, (mode >> 4) & 1, per_group,
           double(c) / (double(iters) * per_group), e == cudaSuccess ? "" : cudaGetErrorString(e));
    cudaFree(d);
}

// How far does the issuing thread run ahead of the tensor pipe?  Per iteration: 4 MMAs, then `pad` dependent integer
// multiply-adds (~4-5 cycles each).  If issue is asynchronous with a deep queue the padding is hidden until it exceeds
// the MMA time; also reports the cycles the thread spends inside the 4 issue instructions.
template <int N>
__global__ void __launch_bounds__(128, 1) issue_probe(int iters, long long* out_cycles, int pad, int seed) {
    extern __shared__ uint8_t raw[];
    const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
    constexpr int STAGE = 128 * 128 + N * 128;
    const uint32_t bar = base + 3 * STAGE;
    const uint32_t slot = bar + 32;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (uint32_t i = threadIdx.x; i < uint32_t(3 * STAGE / 4); i += blockDim.x)
        reinterpret_cast<uint32_t*>(raw + (base - smem_u32(raw)))[i] = 0x3c003c00u;
    if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
    if (warp == 1) { tmem_alloc(slot, 512); tmem_relinquish(); }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<uint32_t*>(raw + (slot - smem_u32(raw)));
    if (warp == 0 && lane == 0) {
        constexpr uint32_t idesc = make_idesc_bf16(128, N);
        const uint64_t ad = make_desc_k_sw128(base), bd = make_desc_k_sw128(base + 128 * 128);
        long long issue = 0;
        int x = seed;
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            const long long a0 = clock64();
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_bf16(tmem, ad + 2 * k, bd + 2 * k, idesc, 1);
            issue += clock64() - a0;
            for (int j = 0; j < pad; ++j) x = x * 1664525 + 1013904223;
        }
        umma_commit(bar);
        mbar_wait(bar, 0);
        const long long t1 = clock64();
        if (blockIdx.x == 0) { out_cycles[0] = t1 - t0; out_cycles[1] = issue; out_cycles[2] = x; }
    }
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, 512);
}

template <int N>
void run_probe(int pad) {
    long long* d;
    cudaMalloc(&d, 24);
    const int smem = 3 * (128 * 128 + N * 128) + 2048;
    cudaFuncSetAttribute(issue_probe<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int iters = 2000;
    issue_probe<N><<<148, 128, smem>>>(iters, d, pad, 7);
    issue_probe<N><<<148, 128, smem>>>(iters, d, pad, 7);
    cudaError_t e = cudaDeviceSynchronize();
    long long c[3] = {0, 0, 0};
    cudaMemcpy(c, d, 24, cudaMemcpyDeviceToHost);
    printf("probe N=%3d pad=%3d : %7.1f cycles / 4-MMA group, of which %6.1f inside the 4 issue instructions  %s\n", N, pad,
           double(c[0]) / iters, double(c[1]) / iters, e == cudaSuccess ? "" : cudaGetErrorString(e));
    cudaFree(d);
}

template <int N>
void run(int grid, int nstages, int same_acc) {
    long long* d;
    cudaMalloc(&d, 8);
    const int smem = nstages * (128 * 128 + N * 128) + 2048;
    cudaFuncSetAttribute(mma_rate<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int iters = 2000;
    mma_rate<N><<<grid, 128, smem>>>(iters, d, nstages, same_acc);
    mma_rate<N><<<grid, 128, smem>>>(iters, d, nstages, same_acc);
    cudaError_t e = cudaDeviceSynchronize();
    long long c = 0;
    cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
    const double per = double(c) / (iters * 4.0);
    printf("N=%3d grid=%3d stages=%d same_acc=%d : %7.1f cycles/MMA  -> %6.0f MAC/clk/SM (%4.1f%% of 4096)  %s\n", N, grid, nstages,
           same_acc, per, 128.0 * N * 16 / per, 100.0 * 128.0 * N * 16 / per / 4096.0, e == cudaSuccess ? "" : cudaGetErrorString(e));
    cudaFree(d);
}

int main() {
    for (int grid : {1, 148}) {
        run<64>(grid, 3, 1); run<80>(grid, 3, 1); run<128>(grid, 3, 1); run<160>(grid, 3, 1); run<192>(grid, 3, 1);
        run<256>(grid, 3, 1);
    }
    run<160>(148, 1, 1); run<160>(148, 3, 0); run<256>(148, 3, 0);
    for (int pad : {0, 8, 16, 32, 48, 64, 96, 128, 192}) run_probe<160>(pad);
    for (int pad : {0, 32, 64}) run_probe<256>(pad);
    for (int mode : {0, 1, 2, 3, 7, 9, 11, 18, 19}) run_overhead<160>(mode, 1);
    for (int mode : {0, 3, 7}) run_overhead<160>(mode, 2);
    for (int mode : {0, 3}) run_overhead<128>(mode, 1);
    for (int mode : {0, 3}) run_overhead<256>(mode, 1);
    return 0;
}
