// First convolution of the matching-feature net: Conv2d(3, 32, 3, stride 2, pad 1) + eval-mode BN (folded) + ReLU
// (networks/psm_submodule.py:42-44, `convbn(3, 32, 3, 2, 1, 1)` + ReLU) straight from the NCHW image stack into the vol4 layout
// the planar tensor-core layers read.  3 input channels do not make a tensor-core operand (K = 27), and the layer is a memory
// problem anyway: 18.4 MB in, 49.2 MB out, 0.66 GFLOP at 5 x 480 x 640.  cuDNN's fp32 implicit-GEMM kernel needs 127 us for it
// and leaves an NCHW tensor that costs another 21 us to turn into vol4; this kernel does both in one pass.
//   thread = one output pixel x all 32 channels (27 image taps in registers, weights broadcast from shared memory);
//   a warp = 32 consecutive output columns, so every chunk store is 512 contiguous bytes.
#include "common.cuh"
#include "conv3d_common.cuh"

namespace estd {

constexpr int kStemCout = 32, kStemTaps = 27;

// SPLIT: write the result pre-split (vol4s: chunk 2g = 8 x fp16 x_hi of channels 8g..8g+7, chunk 2g+1 = x_lo)
template <bool SPLIT>
__global__ void __launch_bounds__(256) stem_conv_kernel(const float* __restrict__ img, const float* __restrict__ weight,
                                                        const float* __restrict__ bias, float* __restrict__ out,
                                                        int N, int H, int W, int Ho, int Wo, int* status) {
    __shared__ float s_w[kStemTaps][kStemCout];            // [tap = (ci*3 + ky)*3 + kx][cout]
    __shared__ float s_b[kStemCout];
    for (int i = threadIdx.x; i < kStemTaps * kStemCout; i += blockDim.x) {
        const int co = i % kStemCout, tap = i / kStemCout;
        s_w[tap][co] = weight[co * kStemTaps + tap];       // weight is [cout][3][3][3] (PyTorch layout)
    }
    if (threadIdx.x < kStemCout) s_b[threadIdx.x] = bias[threadIdx.x];
    __syncthreads();
    const int wo = blockIdx.x * 32 + (threadIdx.x & 31);
    const int ho = blockIdx.y * 8 + (threadIdx.x >> 5);
    const int n = blockIdx.z;
    if (wo >= Wo || ho >= Ho) return;
    float x[kStemTaps];
    const float* base = img + (size_t)n * 3 * H * W;
#pragma unroll
    for (int ci = 0; ci < 3; ++ci)
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            const int hi = 2 * ho - 1 + ky;
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const int wi = 2 * wo - 1 + kx;
                const bool in = hi >= 0 && hi < H && wi >= 0 && wi < W;
                x[(ci * 3 + ky) * 3 + kx] = in ? __ldg(base + ((size_t)ci * H + hi) * W + wi) : 0.0f;      // zero padding
            }
        }
    const size_t plane = (size_t)N * Ho * Wo * 4;          // floats per chunk
    float* dst = out + (((size_t)n * Ho + ho) * Wo + wo) * 4;
    float amax = 0.0f;
#pragma unroll 1
    for (int g = 0; g < kStemCout / 8; ++g) {
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = s_b[8 * g + j];
#pragma unroll
        for (int tap = 0; tap < kStemTaps; ++tap) {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = fmaf(s_w[tap][8 * g + j], x[tap], v[j]);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.0f);
        if (SPLIT) {
            uint4 hi, lo;
            split8(v, hi, lo, amax);
            *reinterpret_cast<uint4*>(dst + (size_t)(2 * g) * plane) = hi;
            *reinterpret_cast<uint4*>(dst + (size_t)(2 * g + 1) * plane) = lo;
        } else {
            st4(dst + (size_t)(2 * g) * plane, make_float4(v[0], v[1], v[2], v[3]));
            st4(dst + (size_t)(2 * g + 1) * plane, make_float4(v[4], v[5], v[6], v[7]));
        }
    }
    if (SPLIT && !(amax <= 65504.0f) && status) atomicOr(status, 1);
}

}  // namespace estd

extern "C" int estd_stem_conv(const float* img_nchw, const float* weight, const float* bias, float* out_vol4, int N, int H, int W,
                              int out_split, int* status, void* stream) {
    ESTD_REQUIRE(img_nchw && weight && bias && out_vol4, "estd_stem_conv: null pointer");
    ESTD_REQUIRE(N > 0 && N <= 65535 && H > 1 && W > 1 && estd::aligned16(out_vol4), "estd_stem_conv: bad arguments");
    const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;   // floor((H + 2 - 3) / 2) + 1
    const dim3 grid((Wo + 31) / 32, (Ho + 7) / 8, N);
    ESTD_REQUIRE(grid.y <= 65535, "estd_stem_conv: image too tall");
    if (out_split) estd::stem_conv_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(img_nchw, weight, bias, out_vol4, N, H, W, Ho, Wo, status);
    else           estd::stem_conv_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(img_nchw, weight, bias, out_vol4, N, H, W, Ho, Wo, status);
    return estd::check_launch("estd_stem_conv");
}
