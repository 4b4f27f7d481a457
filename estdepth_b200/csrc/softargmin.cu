// K4: 1x1x1 logit head + softmax over depth + expectation (soft-argmin) at quarter resolution, written
// `up` x `up` replicated, plus the small layout conversions at the boundary.
//
// Reference seam: nn.Conv3d(16,1,1,bias=True) of stereo_head{0,1} (hybrid_depth_decoder.py:106,111),
// F.interpolate(scale_factor=4) (:202,259,359,379) and depthlayer (:33-38).  The reference up-samples the logits 16x
// and then runs softmax / sum / max over a [T,64,480,640] tensor (~0.5 GB of traffic per target); nearest up-sampling
// commutes with a per-pixel softmax (quirk Q11), so this kernel reads the 16-channel hidden volume once, writes the
// quarter-resolution logits (needed by the 2-D refinement, :268) and replicates depth / prob / argmax.
#include "common.cuh"

namespace estd {

// Block = 32 pixels (lanes, consecutive w) x kSlices depth slices (warps).  Each warp owns a contiguous range of planes and
// keeps an online (max, sum, weighted sum, argmax); the 8 partial states of a pixel are merged in plane order by warp 0, so
// "first maximum wins" (torch.max) is preserved.  (v1 walked all D planes in one thread: 150 blocks, latency bound, 12 % of HBM.)
constexpr int kSlices = 8;

struct SoftState { float m, s, ws; int best; };

__device__ __forceinline__ void soft_update(SoftState& st, float l, float dv, int d) {
    if (l > st.m) {                                    // strict: first maximum wins, like torch.max
        const float scale = expf(st.m - l);            // exp(-inf) = 0 on the first plane
        st.s = fmaf(st.s, scale, 1.0f);
        st.ws = fmaf(st.ws, scale, dv);
        st.m = l;
        st.best = d;
    } else {
        const float e = expf(l - st.m);
        st.s += e;
        st.ws = fmaf(e, dv, st.ws);
    }
}

__global__ void __launch_bounds__(32 * kSlices) head_softargmin_kernel(const float* __restrict__ hidden, const float* __restrict__ head_w,
                                                              const float* __restrict__ head_b, const float* __restrict__ logits_in,
                                                              const float* __restrict__ depth_values, float* __restrict__ logits_out,
                                                              float* __restrict__ depth_out, float* __restrict__ prob_out,
                                                              int* __restrict__ argmax_out, int D, int H, int W, int up) {
    __shared__ SoftState part[kSlices][32];
    const int HW = H * W;
    const int lane = threadIdx.x & 31, slice = threadIdx.x >> 5;
    const int p = blockIdx.x * 32 + lane;
    const bool valid = p < HW;
    const int pc = valid ? p : HW - 1;
    const size_t vox = (size_t)D * HW;
    const int per = (D + kSlices - 1) / kSlices;
    const int d_begin = slice * per, d_end = min(D, d_begin + per);
    float4 hw4[4];
    float bias = 0.0f;
    if (hidden) {
#pragma unroll
        for (int j = 0; j < 4; ++j) hw4[j] = ldg4(head_w + j * 4);
        bias = __ldg(head_b);
    }
    SoftState st = {-INFINITY, 0.0f, 0.0f, d_begin};
#pragma unroll 4
    for (int d = d_begin; d < d_end; ++d) {
        float l;
        if (hidden) {
            l = 0.0f;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float4 x = ldg4(hidden + (j * vox + (size_t)d * HW + pc) * 4);
                l = fmaf(x.x, hw4[j].x, l); l = fmaf(x.y, hw4[j].y, l);
                l = fmaf(x.z, hw4[j].z, l); l = fmaf(x.w, hw4[j].w, l);
            }
            l += bias;
        } else {
            l = __ldg(logits_in + (size_t)d * HW + pc);
        }
        if (logits_out && valid) logits_out[(size_t)d * HW + p] = l;
        soft_update(st, l, __ldg(depth_values + d), d);
    }
    part[slice][lane] = st;
    __syncthreads();
    if (slice != 0 || !valid) return;
    for (int k = 1; k < kSlices; ++k) {                // merge in plane order
        const SoftState o = part[k][lane];
        if (o.s == 0.0f) continue;                     // empty slice (D < kSlices * per)
        if (o.m > st.m) {
            const float scale = expf(st.m - o.m);
            st.s = fmaf(st.s, scale, o.s);
            st.ws = fmaf(st.ws, scale, o.ws);
            st.m = o.m;
            st.best = o.best;
        } else {
            const float scale = expf(o.m - st.m);
            st.s = fmaf(o.s, scale, st.s);
            st.ws = fmaf(o.ws, scale, st.ws);
        }
    }
    const int h = p / W, w = p - h * W;
    const float depth = __fdiv_rn(st.ws, st.s);
    const float prob = __fdiv_rn(1.0f, st.s);
    const int best = st.best;
    const int WU = W * up;
    for (int r = 0; r < up; ++r) {
        const size_t row = ((size_t)(h * up + r)) * WU + (size_t)w * up;
        if (up == 4) {
            if (depth_out) st4(depth_out + row, make_float4(depth, depth, depth, depth));
            if (prob_out) st4(prob_out + row, make_float4(prob, prob, prob, prob));
            if (argmax_out) *reinterpret_cast<int4*>(argmax_out + row) = make_int4(best, best, best, best);
        } else {
            for (int c = 0; c < up; ++c) {
                if (depth_out) depth_out[row + c] = depth;
                if (prob_out) prob_out[row + c] = prob;
                if (argmax_out) argmax_out[row + c] = best;
            }
        }
    }
}

__global__ void __launch_bounds__(256) vol4_to_ncdhw_kernel(const float* __restrict__ vol4, float* __restrict__ ncdhw,
                                                            size_t vox, size_t total) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t j = i / vox, v = i - j * vox;
        const float4 x = ldg4(vol4 + i * 4);
        float* o = ncdhw + (j * 4) * vox + v;
        o[0] = x.x; o[vox] = x.y; o[2 * vox] = x.z; o[3 * vox] = x.w;
    }
}

__global__ void __launch_bounds__(256) ncdhw_to_vol4_kernel(const float* __restrict__ ncdhw, float* __restrict__ vol4,
                                                            size_t vox, size_t total) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t j = i / vox, v = i - j * vox;
        const float* s = ncdhw + (j * 4) * vox + v;
        st4(vol4 + i * 4, make_float4(__ldg(s), __ldg(s + vox), __ldg(s + 2 * vox), __ldg(s + 3 * vox)));
    }
}

// [N][C][H][W] (torch NCHW) <-> vol4 [C/4][N][H][W][4]: the stack of N feature maps seen as a volume with D = N planes
__global__ void __launch_bounds__(256) nchw_to_vol4_kernel(const float* __restrict__ nchw, float* __restrict__ vol4,
                                                           int C, int N, size_t HW, size_t total) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t p = i % HW;
        const size_t n = (i / HW) % N;
        const size_t j = i / (HW * N);
        const float* s = nchw + (n * C + j * 4) * HW + p;
        st4(vol4 + i * 4, make_float4(__ldg(s), __ldg(s + HW), __ldg(s + 2 * HW), __ldg(s + 3 * HW)));
    }
}

__global__ void __launch_bounds__(256) vol4_to_nchw_kernel(const float* __restrict__ vol4, float* __restrict__ nchw,
                                                           int C, int N, size_t HW, size_t total) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t p = i % HW;
        const size_t n = (i / HW) % N;
        const size_t j = i / (HW * N);
        const float4 x = ldg4(vol4 + i * 4);
        float* o = nchw + (n * C + j * 4) * HW + p;
        o[0] = x.x; o[HW] = x.y; o[2 * HW] = x.z; o[3 * HW] = x.w;
    }
}

// Bilinear resize (ATen upsample_bilinear2d, align_corners=False: src = max(0, (dst + 0.5) * in/out - 0.5), second tap clamped to
// the last row / column) of small NCHW maps [N][C][h][w] straight into vol4 [C/4][N][H][W][4], with an optional per-channel
// offset + ReLU applied to the taps as they are read (the SPP branches of the matching-feature net: 1x1 conv + BN + ReLU on the
// pooled map, F.upsample, torch.cat -- networks/psm_submodule.py:56-76,104-114).  One thread = one output pixel of one chunk.
__global__ void __launch_bounds__(256) upsample_bilinear_vol4_kernel(const float* __restrict__ src, const float* __restrict__ bias,
                                                                     float* __restrict__ vol4, int C, int N, int h, int w,
                                                                     int H, int W, int relu, size_t total) {
    const float rh = (float)h / (float)H, rw = (float)w / (float)W;
    const size_t HW = (size_t)H * W, hw = (size_t)h * w;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t p = i % HW;
        const size_t n = (i / HW) % N;
        const size_t j = i / (HW * N);
        const int y = (int)(p / W), x = (int)(p - (size_t)y * W);
        const float sy = fmaxf(__fadd_rn(__fmul_rn(rh, (float)y + 0.5f), -0.5f), 0.0f);
        const float sx = fmaxf(__fadd_rn(__fmul_rn(rw, (float)x + 0.5f), -0.5f), 0.0f);
        const int y0 = (int)sy, x0 = (int)sx;
        const int yp = (y0 < h - 1) ? w : 0, xp = (x0 < w - 1) ? 1 : 0;
        const float ly1 = sy - (float)y0, ly0 = 1.0f - ly1, lx1 = sx - (float)x0, lx0 = 1.0f - lx1;
        float o[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int c = (int)j * 4 + k;
            const float* s = src + (n * C + c) * hw + (size_t)y0 * w + x0;
            float v00 = __ldg(s), v01 = __ldg(s + xp), v10 = __ldg(s + yp), v11 = __ldg(s + yp + xp);
            if (bias) { const float b = __ldg(bias + c); v00 += b; v01 += b; v10 += b; v11 += b; }
            if (relu) { v00 = fmaxf(v00, 0.0f); v01 = fmaxf(v01, 0.0f); v10 = fmaxf(v10, 0.0f); v11 = fmaxf(v11, 0.0f); }
            o[k] = __fadd_rn(__fmul_rn(ly0, __fadd_rn(__fmul_rn(lx0, v00), __fmul_rn(lx1, v01))),
                             __fmul_rn(ly1, __fadd_rn(__fmul_rn(lx0, v10), __fmul_rn(lx1, v11))));
        }
        st4(vol4 + i * 4, make_float4(o[0], o[1], o[2], o[3]));
    }
}

__global__ void __launch_bounds__(256) scalar_to_vol4_kernel(const float* __restrict__ in, float* __restrict__ out, size_t vox) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < vox; i += (size_t)gridDim.x * blockDim.x)
        st4(out + i * 4, make_float4(__ldg(in + i), 0.0f, 0.0f, 0.0f));
}

static int ew_blocks(size_t total) {
    size_t b = (total + 255) / 256;
    return (int)(b < 148 * 16 ? b : 148 * 16);
}

}  // namespace estd

extern "C" int estd_head_softargmin(const float* hidden_vol4, const float* head_w, const float* head_b, const float* logits_in,
                                    const float* depth_values, float* logits_out, float* depth_out, float* prob_out,
                                    int* argmax_out, int D, int H, int W, int up, void* stream) {
    using namespace estd;
    ESTD_REQUIRE(depth_values && D > 0 && H > 0 && W > 0 && up >= 1 && up <= 8, "estd_head_softargmin: bad arguments");
    ESTD_REQUIRE((hidden_vol4 && head_w && head_b) || logits_in, "estd_head_softargmin: need hidden+head or logits_in");
    if (up == 4) ESTD_REQUIRE((!depth_out || aligned16(depth_out)) && (!prob_out || aligned16(prob_out)) &&
                              (!argmax_out || aligned16(argmax_out)), "estd_head_softargmin: outputs must be 16-byte aligned");
    head_softargmin_kernel<<<(H * W + 31) / 32, 32 * kSlices, 0, (cudaStream_t)stream>>>(
        hidden_vol4, head_w, head_b, hidden_vol4 ? nullptr : logits_in, depth_values, logits_out, depth_out, prob_out,
        argmax_out, D, H, W, up);
    return check_launch("estd_head_softargmin");
}

extern "C" int estd_vol4_to_ncdhw(const float* vol4, float* ncdhw, int C, int D, int H, int W, void* stream) {
    ESTD_REQUIRE(vol4 && ncdhw && C > 0 && (C % 4) == 0 && D > 0 && H > 0 && W > 0, "estd_vol4_to_ncdhw: bad arguments");
    const size_t vox = (size_t)D * H * W, total = vox * (C / 4);
    estd::vol4_to_ncdhw_kernel<<<estd::ew_blocks(total), 256, 0, (cudaStream_t)stream>>>(vol4, ncdhw, vox, total);
    return estd::check_launch("estd_vol4_to_ncdhw");
}

extern "C" int estd_ncdhw_to_vol4(const float* ncdhw, float* vol4, int C, int D, int H, int W, void* stream) {
    ESTD_REQUIRE(vol4 && ncdhw && C > 0 && (C % 4) == 0 && D > 0 && H > 0 && W > 0, "estd_ncdhw_to_vol4: bad arguments");
    const size_t vox = (size_t)D * H * W, total = vox * (C / 4);
    estd::ncdhw_to_vol4_kernel<<<estd::ew_blocks(total), 256, 0, (cudaStream_t)stream>>>(ncdhw, vol4, vox, total);
    return estd::check_launch("estd_ncdhw_to_vol4");
}

extern "C" int estd_nchw_to_vol4(const float* nchw, float* vol4, int N, int C, int H, int W, void* stream) {
    ESTD_REQUIRE(vol4 && nchw && C > 0 && (C % 4) == 0 && N > 0 && H > 0 && W > 0, "estd_nchw_to_vol4: bad arguments");
    const size_t HW = (size_t)H * W, total = HW * N * (C / 4);
    estd::nchw_to_vol4_kernel<<<estd::ew_blocks(total), 256, 0, (cudaStream_t)stream>>>(nchw, vol4, C, N, HW, total);
    return estd::check_launch("estd_nchw_to_vol4");
}

extern "C" int estd_vol4_to_nchw(const float* vol4, float* nchw, int N, int C, int H, int W, void* stream) {
    ESTD_REQUIRE(vol4 && nchw && C > 0 && (C % 4) == 0 && N > 0 && H > 0 && W > 0, "estd_vol4_to_nchw: bad arguments");
    const size_t HW = (size_t)H * W, total = HW * N * (C / 4);
    estd::vol4_to_nchw_kernel<<<estd::ew_blocks(total), 256, 0, (cudaStream_t)stream>>>(vol4, nchw, C, N, HW, total);
    return estd::check_launch("estd_vol4_to_nchw");
}

extern "C" int estd_upsample_bilinear_vol4(const float* src_nchw, const float* bias, float* vol4, int N, int C, int h, int w,
                                           int H, int W, int relu, void* stream) {
    ESTD_REQUIRE(src_nchw && vol4 && C > 0 && (C % 4) == 0 && N > 0 && h > 0 && w > 0 && H > 0 && W > 0 && estd::aligned16(vol4),
                 "estd_upsample_bilinear_vol4: bad arguments");
    const size_t total = (size_t)H * W * N * (C / 4);
    estd::upsample_bilinear_vol4_kernel<<<estd::ew_blocks(total), 256, 0, (cudaStream_t)stream>>>(src_nchw, bias, vol4, C, N, h, w,
                                                                                                  H, W, relu, total);
    return estd::check_launch("estd_upsample_bilinear_vol4");
}

extern "C" int estd_scalar_to_vol4(const float* dhw, float* vol4_1chunk, int D, int H, int W, void* stream) {
    ESTD_REQUIRE(dhw && vol4_1chunk && D > 0 && H > 0 && W > 0, "estd_scalar_to_vol4: bad arguments");
    const size_t vox = (size_t)D * H * W;
    estd::scalar_to_vol4_kernel<<<estd::ew_blocks(vox), 256, 0, (cudaStream_t)stream>>>(dhw, vol4_1chunk, vox);
    return estd::check_launch("estd_scalar_to_vol4");
}
