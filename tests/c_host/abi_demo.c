/* A host with no Python and no torch: plain C against include/mfb200.h + libmfb200.so + the CUDA runtime.
 * Builds a linear layer (F.linear, S/models/lora.py:445-451) followed by a LayerNorm (S/models/attention.py:313) as a RECORDED
 * PROGRAM (mfb_program_begin / _end), replays it from one call (mfb_program_run) and checks the result against a scalar C
 * evaluation of the same two ops on the same bf16-rounded operands.  Compiled and run by tests/test_gpu_c_host.py.
 *   gcc abi_demo.c -I include -I $CUDA/include -L <libdir> -lmfb200 -L $CUDA/lib64 -lcudart -lm */
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "mfb200.h"

static unsigned short f2bf(float f) {           /* round to nearest even */
    unsigned u;
    memcpy(&u, &f, 4);
    u += 0x7FFFu + ((u >> 16) & 1u);
    return (unsigned short)(u >> 16);
}
static float bf2f(unsigned short h) {
    unsigned u = (unsigned)h << 16;
    float f;
    memcpy(&f, &u, 4);
    return f;
}
static float frand(unsigned* s) {                /* LCG in [-1, 1) */
    *s = *s * 1664525u + 1013904223u;
    return (float)((*s >> 8) & 0xFFFF) / 32768.0f - 1.0f;
}
#define CK(x)                                                                        \
    do {                                                                             \
        int _rc = (x);                                                               \
        if (_rc != 0) {                                                              \
            fprintf(stderr, "%s -> %d: %s\n", #x, _rc, mfb_last_error());            \
            return 1;                                                                \
        }                                                                            \
    } while (0)
#define CU(x)                                                                        \
    do {                                                                             \
        cudaError_t _e = (x);                                                        \
        if (_e != cudaSuccess) {                                                     \
            fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(_e));                 \
            return 1;                                                                \
        }                                                                            \
    } while (0)

int main(void) {
    enum { M = 384, K = 128, N = 64 };
    if (mfb_abi_version() != MFB_ABI_VERSION) return 1;
    CK(mfb_init(0));
    unsigned seed = 12345u;
    unsigned short* hx = malloc(sizeof(short) * M * K);
    unsigned short* hw = malloc(sizeof(short) * N * K);
    float hb[N], hg[N], hbeta[N];
    for (int i = 0; i < M * K; ++i) hx[i] = f2bf(frand(&seed));
    for (int i = 0; i < N * K; ++i) hw[i] = f2bf(frand(&seed) * 0.1f);
    for (int i = 0; i < N; ++i) { hb[i] = frand(&seed) * 0.1f; hg[i] = 1.0f + 0.1f * frand(&seed); hbeta[i] = 0.05f * frand(&seed); }
    void *dx, *dw, *dlin, *dout;
    float *db, *dg, *dbeta;
    CU(cudaMalloc(&dx, sizeof(short) * M * K)); CU(cudaMalloc(&dw, sizeof(short) * N * K));
    CU(cudaMalloc(&dlin, sizeof(short) * M * N)); CU(cudaMalloc(&dout, sizeof(short) * M * N));
    CU(cudaMalloc((void**)&db, sizeof(float) * N)); CU(cudaMalloc((void**)&dg, sizeof(float) * N)); CU(cudaMalloc((void**)&dbeta, sizeof(float) * N));
    CU(cudaMemcpy(dx, hx, sizeof(short) * M * K, cudaMemcpyHostToDevice)); CU(cudaMemcpy(dw, hw, sizeof(short) * N * K, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(db, hb, sizeof(hb), cudaMemcpyHostToDevice)); CU(cudaMemcpy(dg, hg, sizeof(hg), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(dbeta, hbeta, sizeof(hbeta), cudaMemcpyHostToDevice));
    cudaStream_t st;
    CU(cudaStreamCreate(&st));

    /* a linear layer over an [M, K] token matrix is ksize = 1, B = 1, H = 1, W = M, Cin = K (mfb200.h) */
    mfb_conv_desc d;
    memset(&d, 0, sizeof(d));
    d.B = 1; d.H = 1; d.W = M; d.Cin = K; d.Cout = N; d.ksize = 1; d.stride = 1;
    d.x = dx; d.w = dw; d.bias = db; d.out = dlin;
    mfb_plan* plan = NULL;
    CK(mfb_conv_plan_create(&d, &plan));

    mfb_program* prog = NULL;
    CK(mfb_program_begin(&prog));
    CK(mfb_plan_run(plan, st));                                        /* executed and recorded */
    CK(mfb_layernorm(dlin, M, N, 1e-5f, dg, dbeta, dout, st));
    CK(mfb_program_end());
    if (mfb_program_size(prog) != 2) { fprintf(stderr, "program size %d\n", mfb_program_size(prog)); return 1; }
    CU(cudaMemsetAsync(dout, 0xFF, sizeof(short) * M * N, st));        /* poison, then replay from ONE call */
    CK(mfb_program_run(prog, st));
    CU(cudaStreamSynchronize(st));

    unsigned short* hout = malloc(sizeof(short) * M * N);
    CU(cudaMemcpy(hout, dout, sizeof(short) * M * N, cudaMemcpyDeviceToHost));
    double num = 0, den = 0;
    for (int m = 0; m < M; ++m) {
        float row[N];
        double mean = 0, var = 0;
        for (int n = 0; n < N; ++n) {
            float acc = 0;
            for (int k = 0; k < K; ++k) acc += bf2f(hx[m * K + k]) * bf2f(hw[n * K + k]);
            row[n] = bf2f(f2bf(acc + hb[n]));                         /* the GEMM stores bf16 */
            mean += row[n];
        }
        mean /= N;
        for (int n = 0; n < N; ++n) var += (row[n] - mean) * (row[n] - mean);
        const double rstd = 1.0 / sqrt(var / N + 1e-5);
        for (int n = 0; n < N; ++n) {
            const double want = (row[n] - mean) * rstd * hg[n] + hbeta[n];
            const double got = bf2f(hout[m * N + n]);
            num += (got - want) * (got - want);
            den += want * want;
        }
    }
    const double rel = sqrt(num / den);
    printf("c_host: linear %dx%dx%d + LayerNorm as a recorded program, rel-L2 vs scalar C = %.3e\n", M, K, N, rel);
    CK(mfb_program_destroy(prog));
    CK(mfb_plan_destroy(plan));
    if (!(rel < 5e-3)) { fprintf(stderr, "MISMATCH\n"); return 1; }
    printf("C_HOST_OK\n");
    return 0;
}
