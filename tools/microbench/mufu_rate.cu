// Microbenchmark: MUFU.EX2 throughput per SM (ex2.approx.ftz.f32), alone and mixed with FFMA / FADD the way the
// attention softmax issues them.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_rate mufu_rate.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(float* out, int iters, float a, float b, long long* cyc) {
    float x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = a * (threadIdx.x + i);
    float s = 0.f;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float y;
            if (MODE == 0) {            // ex2 only
                asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x[i]));
                x[i] = y;
            } else {                    // fma -> ex2 -> add (softmax inner loop)
                float t = fmaf(x[i], a, b);
                asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(t));
                s += y;
                x[i] = y;
            }
        }
    }
    const long long t1 = clock64();
    float r = s;
#pragma unroll
    for (int i = 0; i < 8; ++i) r += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

template <int MODE>
void run(int threads, int ctas_per_sm) {
    float* d; long long* c;
    const int grid = 148 * ctas_per_sm;
    cudaMalloc(&d, sizeof(float) * grid * threads); cudaMalloc(&c, 8);
    const int iters = 4000;
    k<MODE><<<grid, threads>>>(d, iters, 0.5f, -1.f, c);
    k<MODE><<<grid, threads>>>(d, iters, 0.5f, -1.f, c);
    cudaDeviceSynchronize();
    long long cy; cudaMemcpy(&cy, c, 8, cudaMemcpyDeviceToHost);
    const double per_sm = double(iters) * 8 * threads * ctas_per_sm;     // ex2 lanes per SM
    printf("mode %d  %4d threads x %d CTAs/SM : %.2f ex2 / clk / SM\n", MODE, threads, ctas_per_sm, per_sm / cy);
    cudaFree(d); cudaFree(c);
}

int main() {
    for (int t : {128, 256, 512, 1024}) run<0>(t, 1);
    run<0>(128, 2); run<0>(128, 3);
    for (int t : {128, 256, 512, 1024}) run<1>(t, 1);
    run<1>(128, 2); run<1>(128, 3);
    return 0;
}
