// K1: fused plane-sweep homography warp + bilinear sample + folded pre0  ->  x0 (cost-volume input).
//
// Reference seam: utils/homo_utils.py:458-504 (homo_warping) + hybrid_models/model_hybrid.py:76,90-94
// (ref_volume repeat, cat, pre0 = 1x1x1 conv 64->32 + BN).  Because bilinear sampling with zero padding
// is linear and maps 0 -> 0, pre0 folds into two 32x32 matvecs at 2-D resolution (estd_premix):
//     x0[c,d,h,w] = (s.W_ref) ref[:,h,w] + b   +   warp_d( (s.W_src) src )[c,h,w]
// so ref_volume, the warped volume and their concat (157+157+315 MB at 480x640/D=64) never exist; the kernel
// reads two L2-resident 2.5 MB maps and streams x0 (157 MB) out once: it is HBM-store bound.
//
// Data layout: maps are map4 [C/4][H][W][4], x0 is vol4 [C/4][D][H][W][4].  One thread owns one voxel (d,h,w)
// for all channel chunks: the homography is evaluated once, every global access is a 16-byte vector, and for a
// fixed chunk the 32 lanes of a warp (32 consecutive w) store 512 contiguous bytes.
#include "common.cuh"

namespace estd {

constexpr int kPremixMaxC = 64;

__global__ void __launch_bounds__(256) premix_kernel(const float* __restrict__ fea, const float* __restrict__ weight,
                                                     const float* __restrict__ bias, float* __restrict__ out,
                                                     int cin, int cout, int HW) {
    extern __shared__ float s_w[];                      // [cout][cin] then [cout] bias
    float* s_b = s_w + cout * cin;
    for (int i = threadIdx.x; i < cout * cin; i += blockDim.x) s_w[i] = weight[i];
    for (int i = threadIdx.x; i < cout; i += blockDim.x) s_b[i] = bias ? bias[i] : 0.0f;
    __syncthreads();
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= HW) return;
    float x[kPremixMaxC];
#pragma unroll
    for (int ci = 0; ci < kPremixMaxC; ++ci)
        if (ci < cin) x[ci] = __ldg(fea + (size_t)ci * HW + p);
    for (int c4 = 0; c4 < cout; c4 += 4) {
        float acc[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float* wr = s_w + (c4 + k) * cin;
            float a = 0.0f;
#pragma unroll
            for (int ci = 0; ci < kPremixMaxC; ++ci)
                if (ci < cin) a = fmaf(wr[ci], x[ci], a);
            acc[k] = a + s_b[c4 + k];
        }
        st4(out + ((size_t)(c4 >> 2) * HW + p) * 4, make_float4(acc[0], acc[1], acc[2], acc[3]));
    }
}

template <int ALIGN>
__global__ void __launch_bounds__(256) warp_cost_kernel(const float* __restrict__ ref_mix, const float* __restrict__ src_mix,
                                                        const float* __restrict__ homo12,
                                                        const float* __restrict__ depth_values,
                                                        float* __restrict__ x0, int chunks, int D, int H, int W) {
    const int HW = H * W;
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    const int d = blockIdx.y;
    if (p >= HW) return;
    const int h = p / W;
    const int w = p - h * W;

    // --- homography (utils/homo_utils.py:469-485), fp32, same operation order as the reference ---
    const float r00 = __ldg(homo12 + 0), r01 = __ldg(homo12 + 1), r02 = __ldg(homo12 + 2);
    const float r10 = __ldg(homo12 + 3), r11 = __ldg(homo12 + 4), r12 = __ldg(homo12 + 5);
    const float r20 = __ldg(homo12 + 6), r21 = __ldg(homo12 + 7), r22 = __ldg(homo12 + 8);
    const float t0 = __ldg(homo12 + 9), t1 = __ldg(homo12 + 10), t2 = __ldg(homo12 + 11);
    const float depth = __ldg(depth_values + d);
    const float fx = (float)w, fy = (float)h;
    const float qx = fmaf(r02, 1.0f, fmaf(r01, fy, r00 * fx));
    const float qy = fmaf(r12, 1.0f, fmaf(r11, fy, r10 * fx));
    const float qz = fmaf(r22, 1.0f, fmaf(r21, fy, r20 * fx));
    const float px3 = __fadd_rn(__fmul_rn(qx, depth), t0);
    const float py3 = __fadd_rn(__fmul_rn(qy, depth), t1);
    const float pz3 = __fadd_rn(__fmul_rn(qz, depth), t2);
    const float zden = __fadd_rn(pz3, 1e-8f);
    const float px = __fdiv_rn(px3, zden);
    const float py = __fdiv_rn(py3, zden);
    float xn = __fadd_rn(__fdiv_rn(px, (float)(W - 1) * 0.5f), -1.0f);
    float yn = __fadd_rn(__fdiv_rn(py, (float)(H - 1) * 0.5f), -1.0f);
    xn = force_outside(xn);
    yn = force_outside(yn);
    const float ix = unnormalize(xn, W, ALIGN);
    const float iy = unnormalize(yn, H, ALIGN);
    const float x0f = floorf(ix), y0f = floorf(iy);
    // ATen GridSampler bilinear weights: nw=(x1-ix)(y1-iy), ne=(ix-x0)(y1-iy), sw=(x1-ix)(iy-y0), se=(ix-x0)(iy-y0)
    const float wx1 = ix - x0f, wx0 = (x0f + 1.0f) - ix;
    const float wy1 = iy - y0f, wy0 = (y0f + 1.0f) - iy;
    // NaN / huge coordinates: the float->int conversion saturates and the bounds test rejects the tap
    const int xi = (int)x0f, yi = (int)y0f;
    const bool vx0 = (xi >= 0) && (xi < W), vx1 = (xi + 1 >= 0) && (xi + 1 < W);
    const bool vy0 = (yi >= 0) && (yi < H), vy1 = (yi + 1 >= 0) && (yi + 1 < H);
    const float w_nw = (vx0 && vy0) ? wx0 * wy0 : 0.0f;
    const float w_ne = (vx1 && vy0) ? wx1 * wy0 : 0.0f;
    const float w_sw = (vx0 && vy1) ? wx0 * wy1 : 0.0f;
    const float w_se = (vx1 && vy1) ? wx1 * wy1 : 0.0f;
    const bool any = (vx0 || vx1) && (vy0 || vy1) && (ix == ix) && (iy == iy);
    // clamped tap addresses (weight is already 0 for out-of-bounds taps)
    const int cx0 = min(max(xi, 0), W - 1), cx1 = min(max(xi + 1, 0), W - 1);
    const int cy0 = min(max(yi, 0), H - 1), cy1 = min(max(yi + 1, 0), H - 1);
    const int o_nw = (cy0 * W + cx0) * 4, o_ne = (cy0 * W + cx1) * 4;
    const int o_sw = (cy1 * W + cx0) * 4, o_se = (cy1 * W + cx1) * 4;

    const size_t plane = (size_t)HW * 4;
    const float* refp = ref_mix + (size_t)p * 4;
    float* outp = x0 + ((size_t)d * HW + p) * 4;
    const size_t out_chunk = (size_t)D * HW * 4;
    if (!any) {                                         // whole sample out of range: x0 = ref part (exact zeros added)
#pragma unroll 4
        for (int j = 0; j < chunks; ++j) st4(outp + j * out_chunk, ldg4(refp + j * plane));
        return;
    }
#pragma unroll 2
    for (int j = 0; j < chunks; ++j) {
        const float* s = src_mix + j * plane;
        const float4 a = ldg4(s + o_nw), b = ldg4(s + o_ne), c = ldg4(s + o_sw), e = ldg4(s + o_se);
        const float4 r = ldg4(refp + j * plane);
        float4 o;
        o.x = r.x + fmaf(e.x, w_se, fmaf(c.x, w_sw, fmaf(b.x, w_ne, a.x * w_nw)));
        o.y = r.y + fmaf(e.y, w_se, fmaf(c.y, w_sw, fmaf(b.y, w_ne, a.y * w_nw)));
        o.z = r.z + fmaf(e.z, w_se, fmaf(c.z, w_sw, fmaf(b.z, w_ne, a.z * w_nw)));
        o.w = r.w + fmaf(e.w, w_se, fmaf(c.w, w_sw, fmaf(b.w, w_ne, a.w * w_nw)));
        st4(outp + j * out_chunk, o);
    }
}

}  // namespace estd

extern "C" int estd_premix(const float* fea_chw, const float* weight, const float* bias, float* out_map4,
                           int cin, int cout, int H, int W, void* stream) {
    ESTD_REQUIRE(fea_chw && weight && out_map4, "estd_premix: null pointer");
    ESTD_REQUIRE(cin > 0 && cin <= estd::kPremixMaxC && cout > 0 && cout <= 64 && (cout % 4) == 0,
                 "estd_premix: cin=%d cout=%d unsupported (cin<=64, cout<=64, cout%%4==0)", cin, cout);
    ESTD_REQUIRE(H > 0 && W > 0 && estd::aligned16(out_map4), "estd_premix: bad shape/alignment");
    const int HW = H * W;
    const size_t smem = (size_t)(cout * cin + cout) * sizeof(float);
    estd::premix_kernel<<<(HW + 255) / 256, 256, smem, (cudaStream_t)stream>>>(fea_chw, weight, bias, out_map4,
                                                                               cin, cout, HW);
    return estd::check_launch("estd_premix");
}

extern "C" int estd_warp_cost(const float* ref_mix_map4, const float* src_mix_map4, const float* homo12,
                              const float* depth_values, float* x0_vol4, int C, int D, int H, int W,
                              int align_corners, void* stream) {
    ESTD_REQUIRE(ref_mix_map4 && src_mix_map4 && homo12 && depth_values && x0_vol4, "estd_warp_cost: null pointer");
    ESTD_REQUIRE(C > 0 && (C % 4) == 0 && D > 0 && D <= 65535 && H > 1 && W > 1,
                 "estd_warp_cost: unsupported shape C=%d D=%d H=%d W=%d", C, D, H, W);
    ESTD_REQUIRE(estd::aligned16(ref_mix_map4) && estd::aligned16(src_mix_map4) && estd::aligned16(x0_vol4),
                 "estd_warp_cost: tensors must be 16-byte aligned");
    const int HW = H * W;
    dim3 grid((HW + 255) / 256, D);
    if (align_corners)
        estd::warp_cost_kernel<1><<<grid, 256, 0, (cudaStream_t)stream>>>(ref_mix_map4, src_mix_map4, homo12,
                                                                          depth_values, x0_vol4, C / 4, D, H, W);
    else
        estd::warp_cost_kernel<0><<<grid, 256, 0, (cudaStream_t)stream>>>(ref_mix_map4, src_mix_map4, homo12,
                                                                          depth_values, x0_vol4, C / 4, D, H, W);
    return estd::check_launch("estd_warp_cost");
}
