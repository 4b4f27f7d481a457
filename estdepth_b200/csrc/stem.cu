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


// ---------------------------------------------------------------------------------------------------------------------------
// First convolution of the context encoder: torchvision ResNet `conv1` = Conv2d(3, 64, 7, stride 2, pad 3) + eval-mode BN (folded)
// + ReLU (hybrid_models/resnet_encoder.py:40-51 runs encoder.conv1 / bn1 / relu), NCHW images -> vol4 (optionally pre-split).
// 147 taps x 3 channels is no tensor-core shape either (K = 147 of fp32 image data); cuDNN's implicit-GEMM kernel took 204 us for
// 3 x 480 x 640 and left an NCHW tensor.  Here: block = 5 output rows x 64 output columns; the 15 x 136 input tile of each colour
// plane and all 147 x 64 weights sit in shared memory; thread = 2 adjacent output pixels x 16 channels (32 accumulators): per
// input row 3 aligned 16-byte tile loads feed 7 x 32 FMAs against weights broadcast from shared memory.  (A first version with
// 32 channels per thread -- 133 registers, 10 warps per SM -- and the PyTorch weight layout read with a stride ran at 190 us.)
constexpr int kStem7Cout = 64, kStem7Taps = 147, kStem7Rows = 5, kStem7Cols = 64;
constexpr int kStem7TileH = 2 * kStem7Rows + 5, kStem7TileW = 2 * kStem7Cols + 8;   // 15 x 136 (one spare column left, two right)
constexpr int kStem7Parts = 4, kStem7Cpt = kStem7Cout / kStem7Parts;                  // channel quarters; channels per thread
constexpr int kStem7Pairs = kStem7Rows * (kStem7Cols / 2);                            // 160 pixel pairs per tile
constexpr int kStem7Threads = kStem7Pairs * kStem7Parts;                              // 640: 5 warps per channel quarter
constexpr size_t kStem7Smem = (size_t)(kStem7Taps * kStem7Cout + 3 * kStem7TileH * kStem7TileW + kStem7Cout) * sizeof(float);

template <bool SPLIT>
__global__ void __launch_bounds__(kStem7Threads, 1) stem7_conv_kernel(const float* __restrict__ img, const float* __restrict__ weight,
                                                                      const float* __restrict__ bias, float* __restrict__ out,
                                                                      int N, int H, int W, int Ho, int Wo, int* status) {
    extern __shared__ __align__(16) float s7[];
    float* s_w = s7;                                        // [tap = (ci*7 + ky)*7 + kx][64]
    float* s_x = s7 + kStem7Taps * kStem7Cout;              // [ci][15][136]: column j <-> input column 2*X0 - 4 + j
    float* s_b = s_x + 3 * kStem7TileH * kStem7TileW;
    const int X0 = blockIdx.x * kStem7Cols, Y0 = blockIdx.y * kStem7Rows, n = blockIdx.z;
    for (int i = threadIdx.x; i < kStem7Taps * kStem7Cout / 4; i += kStem7Threads)       // weight is [3][7][7][64]: tap-major
        reinterpret_cast<float4*>(s_w)[i] = ldg4(weight + 4 * i);
    if (threadIdx.x < kStem7Cout) s_b[threadIdx.x] = __ldg(bias + threadIdx.x);
    const float* base = img + (size_t)n * 3 * H * W;
    for (int i = threadIdx.x; i < 3 * kStem7TileH * kStem7TileW; i += kStem7Threads) {
        const int j = i % kStem7TileW, r = (i / kStem7TileW) % kStem7TileH, ci = i / (kStem7TileW * kStem7TileH);
        const int hi = 2 * Y0 - 3 + r, wi = 2 * X0 - 4 + j;
        s_x[i] = (hi >= 0 && hi < H && wi >= 0 && wi < W) ? __ldg(base + ((size_t)ci * H + hi) * W + wi) : 0.0f;   // zero padding
    }
    __syncthreads();
    const int part = threadIdx.x / kStem7Pairs;             // channels 16*part .. 16*part + 15 (uniform within a warp: 5 warps per part)
    const int pair = threadIdx.x % kStem7Pairs;
    const int r = pair / (kStem7Cols / 2), p = pair % (kStem7Cols / 2);
    float a0[kStem7Cpt], a1[kStem7Cpt];                     // output pixels (Y0 + r, X0 + 2p) and (Y0 + r, X0 + 2p + 1)
#pragma unroll
    for (int c = 0; c < kStem7Cpt; ++c) a0[c] = a1[c] = s_b[kStem7Cpt * part + c];
#pragma unroll 1
    for (int ci = 0; ci < 3; ++ci) {
#pragma unroll 1
        for (int ky = 0; ky < 7; ++ky) {
            // input columns 2x-3 .. 2x+3 of output column x = X0 + 2p are tile columns 4p+1 .. 4p+7, those of x+1 are 4p+3 .. 4p+9
            const float4* row = reinterpret_cast<const float4*>(s_x + (ci * kStem7TileH + 2 * r + ky) * kStem7TileW) + p;
            const float4 v0 = row[0], v1 = row[1], v2 = row[2];
            const float v[12] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w, v2.x, v2.y, v2.z, v2.w};
            const float* wrow = s_w + ((ci * 7 + ky) * 7) * kStem7Cout + kStem7Cpt * part;
#pragma unroll
            for (int kx = 0; kx < 7; ++kx) {
                const float x0 = v[1 + kx], x1 = v[3 + kx];
#pragma unroll
                for (int c4 = 0; c4 < kStem7Cpt / 4; ++c4) {
                    const float4 wv = *reinterpret_cast<const float4*>(wrow + kx * kStem7Cout + 4 * c4);     // same address in the whole warp
                    a0[4 * c4 + 0] = fmaf(wv.x, x0, a0[4 * c4 + 0]); a1[4 * c4 + 0] = fmaf(wv.x, x1, a1[4 * c4 + 0]);
                    a0[4 * c4 + 1] = fmaf(wv.y, x0, a0[4 * c4 + 1]); a1[4 * c4 + 1] = fmaf(wv.y, x1, a1[4 * c4 + 1]);
                    a0[4 * c4 + 2] = fmaf(wv.z, x0, a0[4 * c4 + 2]); a1[4 * c4 + 2] = fmaf(wv.z, x1, a1[4 * c4 + 2]);
                    a0[4 * c4 + 3] = fmaf(wv.w, x0, a0[4 * c4 + 3]); a1[4 * c4 + 3] = fmaf(wv.w, x1, a1[4 * c4 + 3]);
                }
            }
        }
    }
    const int ho = Y0 + r, wo = X0 + 2 * p;
    if (ho >= Ho) return;
    const size_t plane = (size_t)N * Ho * Wo * 4;           // floats per chunk
    float amax = 0.0f;
#pragma unroll
    for (int q = 0; q < 2; ++q) {                           // the two pixels
        if (wo + q >= Wo) break;
        float* acc = q ? a1 : a0;
        float* dst = out + (((size_t)n * Ho + ho) * Wo + wo + q) * 4 + (size_t)(kStem7Cpt / 4 * part) * plane;
#pragma unroll
        for (int g = 0; g < kStem7Cpt / 8; ++g) {           // 8 channels = 2 chunks
            float v8[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v8[j] = fmaxf(acc[8 * g + j], 0.0f);
            if (SPLIT) {
                uint4 hi, lo;
                split8(v8, hi, lo, amax);
                *reinterpret_cast<uint4*>(dst + (size_t)(2 * g) * plane) = hi;
                *reinterpret_cast<uint4*>(dst + (size_t)(2 * g + 1) * plane) = lo;
            } else {
                st4(dst + (size_t)(2 * g) * plane, make_float4(v8[0], v8[1], v8[2], v8[3]));
                st4(dst + (size_t)(2 * g + 1) * plane, make_float4(v8[4], v8[5], v8[6], v8[7]));
            }
        }
    }
    if (SPLIT && !(amax <= 65504.0f) && status) atomicOr(status, 1);
}

// MaxPool2d(3, stride 2, pad 1) over vol4 / vol4s maps (torchvision ResNet `maxpool` after the stem): [C/4][N][H][W][4] ->
// [C/4][N][Ho][Wo][4]; thread = one output pixel x 8 channels (a chunk pair, so that both forms are handled alike).
template <bool IN_SPLIT, bool OUT_SPLIT>
__global__ void __launch_bounds__(256) maxpool3x3s2_vol4_kernel(const float* __restrict__ in, float* __restrict__ out, int pairs, int N,
                                                                int H, int W, int Ho, int Wo, int* status) {
    const size_t total = (size_t)pairs * N * Ho * Wo;
    const size_t iplane = (size_t)N * H * W * 4, oplane = (size_t)N * Ho * Wo * 4;
    float amax = 0.0f;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int wo = (int)(i % Wo), ho = (int)((i / Wo) % Ho), n = (int)((i / ((size_t)Wo * Ho)) % N), g = (int)(i / ((size_t)Wo * Ho * N));
        float m[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) m[j] = -INFINITY;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            const int hi = 2 * ho - 1 + ky;
            if (hi < 0 || hi >= H) continue;
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const int wi = 2 * wo - 1 + kx;
                if (wi < 0 || wi >= W) continue;
                const float* src = in + (size_t)(2 * g) * iplane + (((size_t)n * H + hi) * W + wi) * 4;
                const float4 c0 = ldg4(src), c1 = ldg4(src + iplane);
                float v[8];
                if (IN_SPLIT) join8(c0, c1, v);
                else { v[0] = c0.x; v[1] = c0.y; v[2] = c0.z; v[3] = c0.w; v[4] = c1.x; v[5] = c1.y; v[6] = c1.z; v[7] = c1.w; }
#pragma unroll
                for (int j = 0; j < 8; ++j) m[j] = fmaxf(m[j], v[j]);
            }
        }
        float* dst = out + (size_t)(2 * g) * oplane + (((size_t)n * Ho + ho) * Wo + wo) * 4;
        if (OUT_SPLIT) {
            uint4 hi, lo;
            split8(m, hi, lo, amax);
            *reinterpret_cast<uint4*>(dst) = hi;
            *reinterpret_cast<uint4*>(dst + oplane) = lo;
        } else {
            st4(dst, make_float4(m[0], m[1], m[2], m[3]));
            st4(dst + oplane, make_float4(m[4], m[5], m[6], m[7]));
        }
    }
    if (OUT_SPLIT && !(amax <= 65504.0f) && status) atomicOr(status, 1);
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

extern "C" int estd_stem7_conv(const float* img_nchw, const float* weight, const float* bias, float* out_vol4, int N, int H, int W,
                               int out_split, int* status, void* stream) {
    using namespace estd;
    ESTD_REQUIRE(img_nchw && weight && bias && out_vol4, "estd_stem7_conv: null pointer");
    ESTD_REQUIRE(N > 0 && N <= 65535 && H > 6 && W > 6 && aligned16(out_vol4), "estd_stem7_conv: bad arguments");
    const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;   // floor((H + 6 - 7) / 2) + 1
    const dim3 grid((Wo + kStem7Cols - 1) / kStem7Cols, (Ho + kStem7Rows - 1) / kStem7Rows, N);
    ESTD_REQUIRE(grid.y <= 65535, "estd_stem7_conv: image too tall");
    auto kern = out_split ? stem7_conv_kernel<true> : stem7_conv_kernel<false>;
    static DeviceOnce attr_set[2];               // the opt-in shared-memory limit is a PER-DEVICE function attribute
    if (!attr_set[out_split ? 1 : 0].done()) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kStem7Smem);
        if (e != cudaSuccess) return fail(ESTD_ECUDA, "estd_stem7_conv: cannot reserve %zu B of shared memory: %s", kStem7Smem, cudaGetErrorString(e));
        attr_set[out_split ? 1 : 0].set();
    }
    kern<<<grid, kStem7Threads, kStem7Smem, (cudaStream_t)stream>>>(img_nchw, weight, bias, out_vol4, N, H, W, Ho, Wo, status);
    return check_launch("estd_stem7_conv");
}

extern "C" int estd_maxpool3x3s2_vol4(const float* in_vol4, float* out_vol4, int chunks, int N, int H, int W, int in_split, int out_split,
                                      int* status, void* stream) {
    using namespace estd;
    ESTD_REQUIRE(in_vol4 && out_vol4 && aligned16(in_vol4) && aligned16(out_vol4), "estd_maxpool3x3s2_vol4: null or unaligned pointer");
    ESTD_REQUIRE(chunks > 0 && (chunks % 2) == 0 && N > 0 && H > 0 && W > 0, "estd_maxpool3x3s2_vol4: needs an even number of chunks");
    const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;   // floor((H + 2 - 3) / 2) + 1
    const size_t total = (size_t)(chunks / 2) * N * Ho * Wo;
    const size_t want = (total + 255) / 256;
    const int blocks = (int)(want < (size_t)(148 * 16) ? want : (size_t)(148 * 16));
    cudaStream_t st = (cudaStream_t)stream;
    if (in_split && out_split)  maxpool3x3s2_vol4_kernel<true, true><<<blocks, 256, 0, st>>>(in_vol4, out_vol4, chunks / 2, N, H, W, Ho, Wo, status);
    else if (in_split)          maxpool3x3s2_vol4_kernel<true, false><<<blocks, 256, 0, st>>>(in_vol4, out_vol4, chunks / 2, N, H, W, Ho, Wo, status);
    else if (out_split)         maxpool3x3s2_vol4_kernel<false, true><<<blocks, 256, 0, st>>>(in_vol4, out_vol4, chunks / 2, N, H, W, Ho, Wo, status);
    else                        maxpool3x3s2_vol4_kernel<false, false><<<blocks, 256, 0, st>>>(in_vol4, out_vol4, chunks / 2, N, H, W, Ho, Wo, status);
    return check_launch("estd_maxpool3x3s2_vol4");
}
