// tcgen05 weight gradient (wgrad5.cu), called from mfb_conv_wgrad_tc (train.cu).
#pragma once
#include <cuda_runtime.h>

namespace mfb {
bool wgrad5_supported(int Cin, int Cout);
// split-K plan for the dy grid [B, Ho, Wo]: number of slices (= partial tiles [Cout][k*k*Cin] the workspace must hold)
void wgrad5_plan(int B, int Ho, int Wo, int Cin, int Cout, int ksize, int* slices, int* tiles_per_slice);
// part[slice][Cout][k*k*Cin] fp32 partial sums; x [B, Ho*stride, Wo*stride, Cin], dy [B, Ho, Wo, Cout] bf16.  With ONE slice `part`
// may be the gradient tensor itself (accumulate != 0: added to its contents), which saves the partial tile and the reduce pass.
int wgrad5_run(const void* x, const void* dy, int B, int Ho, int Wo, int Cin, int Cout, int ksize, int stride, float* part, int accumulate,
               cudaStream_t st);
}  // namespace mfb
