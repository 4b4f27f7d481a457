// K3 + K5: the EST ("epipolar spatio-temporal transformer") block outside its two convolutions.
//
// K3  estd_est_attend : for each target voxel, warp N source (key,value) volumes into the target frustum
//     (trilinear gather, zeros padding), correlate the 16-channel keys, softmax over the N sources and return
//     h = mean_n(a_n * value_n).  Replaces 2N warp_volume calls (utils/homo_utils.py:240-279 with helpers
//     :40-62, :26-37, :107-134, :170-205), the stack/repeat to [B,16,D,H,W,N] and the attention of
//     transformer/epipolar_transformer.py:62-73 (quirk Q6: MEAN, i.e. the softmax-weighted sum divided by N).
//     Key and value of a source share one sampling grid, so they are gathered together; nothing but h is written.
// K5  estd_gn_finalize / estd_gru_reset / estd_gru_blend : GroupNorm(1,16) statistics (deterministic two-stage
//     reduction of the partial sums produced by the conv epilogues) and the ConvGRU gate arithmetic of
//     transformer/epipolar_transformer.py:31-54,80-83.
#include "common.cuh"

namespace estd {

struct AttendSources {
    const float* keys[ESTD_MAX_SOURCES];
    const float* values[ESTD_MAX_SOURCES];
};

struct Taps3 {
    int off[8];          // voxel offsets (in float4 units within one chunk) of the 8 taps, clamped in-bounds
    float wgt[8];        // trilinear weights, 0 for out-of-bounds taps
    bool any;
};

// Coordinates of warp_volume (Appendix A.2 of SURVEY.md), fp32, reference operation order.
template <int ALIGN>
__device__ __forceinline__ Taps3 volume_taps(const float* __restrict__ m30, float fx, float fy, float depth,
                                             float depth_min, float depth_interval, int D, int H, int W) {
    const float* Kinv = m30;
    const float* Minv = m30 + 9;
    const float* K = m30 + 21;
    // pixel2cam (homo_utils.py:51-54): K^-1 (x,y,1) * depth
    const float rx = fmaf(__ldg(Kinv + 2), 1.0f, fmaf(__ldg(Kinv + 1), fy, __ldg(Kinv + 0) * fx));
    const float ry = fmaf(__ldg(Kinv + 5), 1.0f, fmaf(__ldg(Kinv + 4), fy, __ldg(Kinv + 3) * fx));
    const float rz = fmaf(__ldg(Kinv + 8), 1.0f, fmaf(__ldg(Kinv + 7), fy, __ldg(Kinv + 6) * fx));
    const float cx = __fmul_rn(rx, depth), cy = __fmul_rn(ry, depth), cz = __fmul_rn(rz, depth);
    // cam2cam (:26-37): inverse(rel_pose) [cam;1]
    const float sx = fmaf(__ldg(Minv + 3), 1.0f, fmaf(__ldg(Minv + 2), cz, fmaf(__ldg(Minv + 1), cy, __ldg(Minv + 0) * cx)));
    const float sy = fmaf(__ldg(Minv + 7), 1.0f, fmaf(__ldg(Minv + 6), cz, fmaf(__ldg(Minv + 5), cy, __ldg(Minv + 4) * cx)));
    const float sz = fmaf(__ldg(Minv + 11), 1.0f, fmaf(__ldg(Minv + 10), cz, fmaf(__ldg(Minv + 9), cy, __ldg(Minv + 8) * cx)));
    // cam2pixel_depth (:116-122)
    const float u = fmaf(__ldg(K + 2), sz, fmaf(__ldg(K + 1), sy, __ldg(K + 0) * sx));
    const float v = fmaf(__ldg(K + 5), sz, fmaf(__ldg(K + 4), sy, __ldg(K + 3) * sx));
    const float z = fmaf(__ldg(K + 8), sz, fmaf(__ldg(K + 7), sy, __ldg(K + 6) * sx));
    const float zden = __fadd_rn(z, 1e-10f);
    const float px = __fdiv_rn(u, zden), py = __fdiv_rn(v, zden);
    // normalize_pixel_coords_volume (:183-198)
    float xn = __fadd_rn(__fdiv_rn(__fmul_rn(2.0f, px), (float)(W - 1)), -1.0f);
    float yn = __fadd_rn(__fdiv_rn(__fmul_rn(2.0f, py), (float)(H - 1)), -1.0f);
    float zn = __fadd_rn(__fdiv_rn(__fmul_rn(2.0f, __fdiv_rn(__fadd_rn(z, -depth_min), depth_interval)), (float)(D - 1)), -1.0f);
    xn = force_outside(xn); yn = force_outside(yn); zn = force_outside(zn);
    const float ix = unnormalize(xn, W, ALIGN), iy = unnormalize(yn, H, ALIGN), iz = unnormalize(zn, D, ALIGN);
    const float x0 = floorf(ix), y0 = floorf(iy), z0 = floorf(iz);
    const float wx[2] = {(x0 + 1.0f) - ix, ix - x0};
    const float wy[2] = {(y0 + 1.0f) - iy, iy - y0};
    const float wz[2] = {(z0 + 1.0f) - iz, iz - z0};
    const int xi = (int)x0, yi = (int)y0, zi = (int)z0;
    Taps3 t;
    const bool finite = (ix == ix) && (iy == iy) && (iz == iz);
    bool any = false;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int dx = k & 1, dy = (k >> 1) & 1, dz = k >> 2;
        const int X = xi + dx, Y = yi + dy, Z = zi + dz;
        const bool ok = finite && X >= 0 && X < W && Y >= 0 && Y < H && Z >= 0 && Z < D;
        any |= ok;
        const int Xc = min(max(X, 0), W - 1), Yc = min(max(Y, 0), H - 1), Zc = min(max(Z, 0), D - 1);
        t.off[k] = (Zc * H + Yc) * W + Xc;
        t.wgt[k] = ok ? wx[dx] * wy[dy] * wz[dz] : 0.0f;
    }
    t.any = any;
    return t;
}

// Block = 32 voxels (lanes, consecutive w) x 4 channel chunks (warps).  Warp n first evaluates the sampling taps of
// source n for the block's 32 voxels and parks them in shared memory; then every warp gathers ITS 4-channel chunk of each
// source's key (partial correlations are summed across the 4 warps through shared memory, in fixed order) and of each
// source's value.  Per load instruction a warp touches 32 consecutive voxels of one chunk = ~512 contiguous bytes.
// (v1 ran one thread per voxel over all 16+16 channels: 94 registers, 31 % occupancy, 30 % of the HBM roofline.)
template <int N, int ALIGN>
__global__ void __launch_bounds__(128) est_attend_kernel(const float* __restrict__ key_t, const AttendSources src,
                                                         const float* __restrict__ warp30,
                                                         const float* __restrict__ depth_values, float depth_min,
                                                         float depth_interval, float* __restrict__ h_out, int D, int H, int W) {
    __shared__ int s_off[N][8][32];
    __shared__ float s_wgt[N][8][32];
    __shared__ float s_corr[N][4][32];
    const int HW = H * W;
    const int lane = threadIdx.x & 31, j = threadIdx.x >> 5;        // j = channel chunk of this warp
    const int p_raw = blockIdx.x * 32 + lane;
    const bool valid = p_raw < HW;
    const int p = valid ? p_raw : HW - 1;
    const int d = blockIdx.y;
    const int h = p / W, w = p - h * W;
    const size_t vox = (size_t)D * HW;
    const size_t me = (size_t)d * HW + p;
    const float depth = __ldg(depth_values + d);

#pragma unroll
    for (int n = j; n < N; n += 4) {                                 // taps of source n for these 32 voxels
        const Taps3 t = volume_taps<ALIGN>(warp30 + n * 30, (float)w, (float)h, depth, depth_min, depth_interval, D, H, W);
#pragma unroll
        for (int k = 0; k < 8; ++k) { s_off[n][k][lane] = t.off[k]; s_wgt[n][k][lane] = t.wgt[k]; }
    }
    const float4 kt = ldg4(key_t + (j * vox + me) * 4);
    __syncthreads();

#pragma unroll
    for (int n = 0; n < N; ++n) {                                    // partial correlation over this warp's 4 channels
        const float* kp = src.keys[n] + (size_t)j * vox * 4;
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float wk = s_wgt[n][k][lane];
            const float4 q = ldg4(kp + (size_t)s_off[n][k][lane] * 4);
            a.x = fmaf(q.x, wk, a.x); a.y = fmaf(q.y, wk, a.y); a.z = fmaf(q.z, wk, a.z); a.w = fmaf(q.w, wk, a.w);
        }
        s_corr[n][j][lane] = fmaf(kt.w, a.w, fmaf(kt.z, a.z, fmaf(kt.y, a.y, kt.x * a.x)));
    }
    __syncthreads();

    float corr[N];
    float m = -INFINITY;
#pragma unroll
    for (int n = 0; n < N; ++n) {                                    // same order in all 4 warps -> identical weights
        corr[n] = ((s_corr[n][0][lane] + s_corr[n][1][lane]) + s_corr[n][2][lane]) + s_corr[n][3][lane];
        m = fmaxf(m, corr[n]);
    }
    float den = 0.0f;
#pragma unroll
    for (int n = 0; n < N; ++n) { corr[n] = expf(corr[n] - m); den += corr[n]; }   // softmax over sources (:69)

    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int n = 0; n < N; ++n) {
        const float a_n = __fdiv_rn(corr[n], den);
        const float* vp = src.values[n] + (size_t)j * vox * 4;
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float wk = s_wgt[n][k][lane];
            const float4 q = ldg4(vp + (size_t)s_off[n][k][lane] * 4);
            a.x = fmaf(q.x, wk, a.x); a.y = fmaf(q.y, wk, a.y); a.z = fmaf(q.z, wk, a.z); a.w = fmaf(q.w, wk, a.w);
        }
        acc.x = fmaf(a.x, a_n, acc.x); acc.y = fmaf(a.y, a_n, acc.y);
        acc.z = fmaf(a.z, a_n, acc.z); acc.w = fmaf(a.w, a_n, acc.w);
    }
    const float inv_n = 1.0f / (float)N;          // torch.mean over the source axis (quirk Q6)
    if (valid) st4(h_out + (j * vox + me) * 4, make_float4(acc.x * inv_n, acc.y * inv_n, acc.z * inv_n, acc.w * inv_n));
}

template <int N>
static int launch_attend(const float* key_t, const AttendSources& src, const float* warp30, const float* depth_values,
                         float depth_min, float depth_interval, float* h_out, int D, int H, int W, int align, cudaStream_t s) {
    dim3 grid((H * W + 31) / 32, D);
    if (align) est_attend_kernel<N, 1><<<grid, 128, 0, s>>>(key_t, src, warp30, depth_values, depth_min, depth_interval, h_out, D, H, W);
    else       est_attend_kernel<N, 0><<<grid, 128, 0, s>>>(key_t, src, warp30, depth_values, depth_min, depth_interval, h_out, D, H, W);
    return check_launch("estd_est_attend");
}

// ---------------------------------------------------------------- GroupNorm statistics + GRU glue
__global__ void gn_finalize_kernel(const double* __restrict__ partials, int n_rows, int n_groups, double count,
                                   float eps, float* __restrict__ stats) {
    const int g = threadIdx.x;
    if (g >= n_groups) return;
    double s = 0.0, q = 0.0;
    for (int r = 0; r < n_rows; ++r) {           // fixed order -> bit-reproducible
        s += partials[((size_t)r * 2 + g) * 2 + 0];
        q += partials[((size_t)r * 2 + g) * 2 + 1];
    }
    const double mean = s / count;
    double var = q / count - mean * mean;        // biased variance, as nn.GroupNorm
    if (var < 0.0) var = 0.0;
    stats[g * 2 + 0] = (float)mean;
    stats[g * 2 + 1] = (float)(1.0 / sqrt(var + (double)eps));
}

__device__ __forceinline__ float4 gn4(float4 x, float mean, float rstd, const float* __restrict__ gamma,
                                      const float* __restrict__ beta, int c) {
    float4 y;
    y.x = fmaf((x.x - mean) * rstd, __ldg(gamma + c + 0), __ldg(beta + c + 0));
    y.y = fmaf((x.y - mean) * rstd, __ldg(gamma + c + 1), __ldg(beta + c + 1));
    y.z = fmaf((x.z - mean) * rstd, __ldg(gamma + c + 2), __ldg(beta + c + 2));
    y.w = fmaf((x.w - mean) * rstd, __ldg(gamma + c + 3), __ldg(beta + c + 3));
    return y;
}

// rh = sigmoid(GN_r(f[0:16])) * h
__global__ void __launch_bounds__(256) gru_reset_kernel(const float* __restrict__ f, const float* __restrict__ h,
                                                        const float* __restrict__ stats, const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, float* __restrict__ rh, size_t vox) {
    const size_t total = 4 * vox;
    const float mean = __ldg(stats + 0), rstd = __ldg(stats + 1);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(i / vox) * 4;
        const float4 r = gn4(ldg4(f + i * 4), mean, rstd, gamma, beta, c);
        const float4 hv = ldg4(h + i * 4);
        st4(rh + i * 4, make_float4(sigmoidf_acc(r.x) * hv.x, sigmoidf_acc(r.y) * hv.y, sigmoidf_acc(r.z) * hv.z,
                                    sigmoidf_acc(r.w) * hv.w));
    }
}

// out = u*h + (1-u)*tanh(GN_o(o)),  u = sigmoid(GN_u(f[16:32]))
__global__ void __launch_bounds__(256) gru_blend_kernel(const float* __restrict__ f, const float* __restrict__ h,
                                                        const float* __restrict__ o, const float* __restrict__ stats_f,
                                                        const float* __restrict__ stats_o, const float* __restrict__ gamma_u,
                                                        const float* __restrict__ beta_u, const float* __restrict__ gamma_o,
                                                        const float* __restrict__ beta_o, float* __restrict__ out, size_t vox) {
    const size_t total = 4 * vox;
    const float mean_u = __ldg(stats_f + 2), rstd_u = __ldg(stats_f + 3);
    const float mean_o = __ldg(stats_o + 0), rstd_o = __ldg(stats_o + 1);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(i / vox) * 4;
        const float4 un = gn4(ldg4(f + (i + total) * 4), mean_u, rstd_u, gamma_u, beta_u, c);   // chunks 4..7 of f
        const float4 on = gn4(ldg4(o + i * 4), mean_o, rstd_o, gamma_o, beta_o, c);
        const float4 hv = ldg4(h + i * 4);
        float4 y;
        { const float u = sigmoidf_acc(un.x); y.x = u * hv.x + (1.0f - u) * tanhf(on.x); }
        { const float u = sigmoidf_acc(un.y); y.y = u * hv.y + (1.0f - u) * tanhf(on.y); }
        { const float u = sigmoidf_acc(un.z); y.z = u * hv.z + (1.0f - u) * tanhf(on.z); }
        { const float u = sigmoidf_acc(un.w); y.w = u * hv.w + (1.0f - u) * tanhf(on.w); }
        st4(out + i * 4, y);
    }
}

static int ew_grid(size_t total) {
    size_t b = (total + 255) / 256;
    return (int)(b < 148 * 16 ? b : 148 * 16);
}

}  // namespace estd

extern "C" int estd_est_attend(const float* key_t, int n_src, const float* const* src_keys, const float* const* src_values,
                               const float* warp30, const float* depth_values, float depth_min, float depth_interval,
                               float* h_out, int D, int H, int W, int align_corners, void* stream) {
    using namespace estd;
    ESTD_REQUIRE(key_t && src_keys && src_values && warp30 && depth_values && h_out, "estd_est_attend: null pointer");
    ESTD_REQUIRE(n_src >= 1 && n_src <= ESTD_MAX_SOURCES, "estd_est_attend: n_src=%d out of range 1..%d", n_src, ESTD_MAX_SOURCES);
    ESTD_REQUIRE(D > 1 && D <= 65535 && H > 1 && W > 1, "estd_est_attend: bad volume %dx%dx%d", D, H, W);
    AttendSources src;
    for (int n = 0; n < ESTD_MAX_SOURCES; ++n) {
        src.keys[n] = n < n_src ? src_keys[n] : nullptr;
        src.values[n] = n < n_src ? src_values[n] : nullptr;
        if (n < n_src) ESTD_REQUIRE(src.keys[n] && src.values[n] && aligned16(src.keys[n]) && aligned16(src.values[n]),
                                    "estd_est_attend: source %d pointer null or unaligned", n);
    }
    cudaStream_t s = (cudaStream_t)stream;
#define ESTD_ATT(N) case N: return launch_attend<N>(key_t, src, warp30, depth_values, depth_min, depth_interval, h_out, D, H, W, align_corners, s)
    switch (n_src) {
        ESTD_ATT(1); ESTD_ATT(2); ESTD_ATT(3); ESTD_ATT(4); ESTD_ATT(5); ESTD_ATT(6); ESTD_ATT(7); ESTD_ATT(8);
    }
#undef ESTD_ATT
    return fail(ESTD_EINVAL, "estd_est_attend: unreachable");
}

extern "C" int estd_gn_finalize(const double* partials, int n_rows, int n_groups, double count_per_group, float eps,
                                float* stats, void* stream) {
    ESTD_REQUIRE(partials && stats && n_rows > 0 && n_groups >= 1 && n_groups <= 2 && count_per_group > 0,
                 "estd_gn_finalize: bad arguments");
    estd::gn_finalize_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(partials, n_rows, n_groups, count_per_group, eps, stats);
    return estd::check_launch("estd_gn_finalize");
}

extern "C" int estd_gru_reset(const float* f_vol4, const float* h_vol4, const float* stats, const float* gamma,
                              const float* beta, float* rh_vol4, int D, int H, int W, void* stream) {
    ESTD_REQUIRE(f_vol4 && h_vol4 && stats && gamma && beta && rh_vol4 && D > 0 && H > 0 && W > 0, "estd_gru_reset: bad arguments");
    const size_t vox = (size_t)D * H * W;
    estd::gru_reset_kernel<<<estd::ew_grid(4 * vox), 256, 0, (cudaStream_t)stream>>>(f_vol4, h_vol4, stats, gamma, beta, rh_vol4, vox);
    return estd::check_launch("estd_gru_reset");
}

extern "C" int estd_gru_blend(const float* f_vol4, const float* h_vol4, const float* o_vol4, const float* stats_f,
                              const float* stats_o, const float* gamma_u, const float* beta_u, const float* gamma_o,
                              const float* beta_o, float* out_vol4, int D, int H, int W, void* stream) {
    ESTD_REQUIRE(f_vol4 && h_vol4 && o_vol4 && stats_f && stats_o && gamma_u && beta_u && gamma_o && beta_o && out_vol4 &&
                 D > 0 && H > 0 && W > 0, "estd_gru_blend: bad arguments");
    const size_t vox = (size_t)D * H * W;
    estd::gru_blend_kernel<<<estd::ew_grid(4 * vox), 256, 0, (cudaStream_t)stream>>>(f_vol4, h_vol4, o_vol4, stats_f, stats_o,
                                                                                    gamma_u, beta_u, gamma_o, beta_o, out_vol4, vox);
    return estd::check_launch("estd_gru_blend");
}
