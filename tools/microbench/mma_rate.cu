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
    return 0;
}
