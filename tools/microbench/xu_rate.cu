// Microbenchmark: which pipe does the fp32 -> bf16x2 pack (F2FP.BF16.F32.PACK_AB) use, and does it share it with
// MUFU.EX2?  Per-SM lanes/clk of: (0) ex2 only, (1) F2FP only, (2) the softmax mix 2 x (ffma, ex2, fadd) + 1 F2FP,
// (3) the same mix with the pack done on the integer pipe (add 0x8000 + PRMT), (4) mix without any pack.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o xu_rate xu_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t f2fp(float lo, float hi) { uint32_t r; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo)); return r; }
__device__ __forceinline__ uint32_t pack_int(float lo, float hi) {
    uint32_t a = __float_as_uint(lo) + 0x8000u, b = __float_as_uint(hi) + 0x8000u, r;
    asm volatile("prmt.b32 %0, %1, %2, 0x7632;" : "=r"(r) : "r"(a), "r"(b));
    return r;
}

template <int MODE>
__global__ void k(uint32_t* out, int iters, float a, float b, long long* cyc) {
    float x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = a * (threadIdx.x + i);
    float s = 0.f;
    uint32_t acc = 0;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; i += 2) {
            if (MODE == 0) { x[i] = ex2(x[i]); x[i + 1] = ex2(x[i + 1]); }
            else if (MODE == 1) { uint32_t r = f2fp(x[i], x[i + 1]); acc ^= r; x[i] = __uint_as_float(r); }
            else {
                float e0 = ex2(fmaf(x[i], a, b)), e1 = ex2(fmaf(x[i + 1], a, b));
                s += e0; s += e1;
                if (MODE == 2) acc ^= f2fp(e0, e1);
                if (MODE == 3) acc ^= pack_int(e0, e1);
                x[i] = e0; x[i + 1] = e1;
            }
        }
    }
    const long long t1 = clock64();
    float r = s;
#pragma unroll
    for (int i = 0; i < 8; ++i) r += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc ^ __float_as_uint(r);
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

template <int MODE>
void run(int threads, const char* what) {
    uint32_t* d; long long* c;
    cudaMalloc(&d, 4 * 148 * threads); cudaMalloc(&c, 8);
    const int iters = 4000;
    k<MODE><<<148, threads>>>(d, iters, 0.5f, -1.f, c);
    k<MODE><<<148, threads>>>(d, iters, 0.5f, -1.f, c);
    cudaDeviceSynchronize();
    long long cy; cudaMemcpy(&cy, c, 8, cudaMemcpyDeviceToHost);
    printf("mode %d %-46s %4d threads/SM : %7.1f clk per (8 elements x warp) per SMSP-warp-slot, %.2f elements/clk/SM\n", MODE, what, threads,
           double(cy) / iters / (threads / 128.0), double(iters) * 8 * threads / cy);
    cudaFree(d); cudaFree(c);
}

int main() {
    for (int t : {128, 256, 512}) {
        run<0>(t, "ex2 only");
        run<1>(t, "F2FP pack only (1 per 2 elements)");
        run<4>(t, "ffma+ex2+fadd, no pack");
        run<2>(t, "ffma+ex2+fadd + F2FP pack");
        run<3>(t, "ffma+ex2+fadd + integer pack (2 IADD + PRMT)");
    }
    return 0;
}
