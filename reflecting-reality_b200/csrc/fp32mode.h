// fp32 parity mode (BASELINE config 1: fp32, rel-L2 1e-4 against the reference): the conv / linear plan variant that
// stores activations and weights in fp32 and accumulates with CUDA-core FFMA.  Shares the mfb_conv_desc semantics of the
// tcgen05 path exactly (K-segments, shortcut / tap segments, row bias, alpha, residuals, GEGLU, stride 2, up2x phases).
#pragma once
#include "common.h"

namespace mfb {

struct Conv32Params {
    const float* x;            // main input, NHWC [B, Hin, Win, Cin]
    int B, Hin, Win, Cin;
    int ntaps, dh[9], dw[9];   // input pixel of tap t for GEMM row (b, oh, ow): (oh*stride + dh[t], ow*stride + dw[t])
    int stride;
    int n_extra;               // extra 1x1 K-segments, tensors at FULL output resolution [B, Hf, Wf, exC]
    const float* ex[3];
    int exC[3];
    int Ho, Wo;                // GEMM row geometry: M = B*Ho*Wo
    int o_step, o_py, o_px, Hf, Wf;   // row (b, oh, ow) is pixel (o_step*oh + o_py, o_step*ow + o_px) of the [B, Hf, Wf] output
    const float* w;            // [Cout][ktot], K order = taps (row-major) x Cin, then the extras
    int ktot, Cout;
    const float *bias, *rowbias, *alpha, *res1, *res2;
    int rowbias_ld;
    float* out;
    int out_ld;
    int geglu;
};

// phase (up_py, up_px) of an up2x descriptor, or (-1, -1) for everything else; w = this launch's weight block
int conv32_build(const mfb_conv_desc* d, int up_py, int up_px, const void* w, Conv32Params& q);
int conv32_launch(const Conv32Params& q, cudaStream_t st);

}  // namespace mfb
