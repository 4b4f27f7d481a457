// Shared host/device helpers for libestdepth_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdarg>
#include <atomic>

#include "../../include/estdepth_b200.h"

namespace estd {

// ---- error reporting (thread-local message, never throws across the C ABI) ----
char* error_buffer();                       // 512-byte thread-local buffer
int fail(int code, const char* fmt, ...);
extern std::atomic<unsigned long long> g_launches;

inline int check_launch(const char* what) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(ESTD_ECUDA, "%s: %s", what, cudaGetErrorString(e));
    return ESTD_OK;
}

#define ESTD_REQUIRE(cond, ...) do { if (!(cond)) return ::estd::fail(ESTD_EINVAL, __VA_ARGS__); } while (0)

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ---- device helpers ----
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

__device__ __forceinline__ float apply_act(float v, int act) {
    if (act == ESTD_ACT_RELU) return fmaxf(v, 0.0f);
    if (act == ESTD_ACT_TANH) return tanhf(v);
    return v;
}

__device__ __forceinline__ float sigmoidf_acc(float v) { return 1.0f / (1.0f + expf(-v)); }

// ---- truncation-bias compensation of the tensor-core convolutions ----
// tcgen05.mma adds its 16-product sum into the fp32 accumulator with TRUNCATION, not round-to-nearest: every accumulating MMA
// shrinks the accumulator by a fixed relative amount on average.  Measured on B200 against fp64 (profiles/trunc_probe.py): the
// signed error of a layer is a clean multiplicative bias, slope = -(0.272 n + ~1) * 2^-24 for n accumulating MMAs with a
// non-zero addend -- the same constant for the 3x3x3 ring schedule (n = 162), the output-stationary kernel (54) and the
// planar kernel (36 ... 720), for signed and for non-negative inputs.  A systematic shrink of every layer compounds linearly
// through the ~25 3-D and ~45 2-D layers of the net (the random part only grows like a square root), so the epilogues undo
// it: the accumulator that received n such MMAs is multiplied by 1 + n * kTruncBiasPerMma.  n is counted per output
// element: taps that fall outside the tensor are zero-filled by TMA, add exactly zero and do not truncate anything.
constexpr float kTruncBiasPerMma = 0.272f * 5.9604644775390625e-08f;        // 0.272 * 2^-24
// how many of the 3 taps (offsets -dil, 0, +dil) around index x lie inside [0, n)
__device__ __forceinline__ int taps_inside(int x, int n, int dil) { return 3 - (x < dil ? 1 : 0) - (x >= n - dil ? 1 : 0); }

// ATen grid_sampler_unnormalize (GridSampler.h): align_corners=False -> ((c+1)*size-1)/2, True -> (c+1)/2*(size-1)
__device__ __forceinline__ float unnormalize(float c, int size, int align_corners) {
    return align_corners ? (c + 1.0f) * 0.5f * (float)(size - 1) : ((c + 1.0f) * (float)size - 1.0f) * 0.5f;
}

// coords outside [-1,1] are forced to 2 (homo_utils.py:488-491, :193-198): every tap falls out of bounds.
// NaN compares false and passes through, exactly like the reference (quirk Q10).
__device__ __forceinline__ float force_outside(float c) { return (c > 1.0f || c < -1.0f) ? 2.0f : c; }

}  // namespace estd
